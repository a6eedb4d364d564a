"""The Fortran side of the boundary (fortran/) against the C ABI (include/cales_b200.h).  No Fortran compiler exists in this
image, so the check is structural: the bind(C) interface module is regenerated from the header and compared, and every
call a hand-written wrapper makes names an existing export with the right number of arguments."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_fortran_iface as gen  # noqa: E402


def test_interface_module_is_generated_from_the_header():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_iface.py"), "--check"]).returncode == 0, \
        "fortran/cales_b200_c.f90 is stale: run python tools/gen_fortran_iface.py"
    protos = gen.prototypes()
    txt = open(os.path.join(ROOT, "fortran", "cales_b200_c.f90")).read()
    assert len(protos) == txt.count("bind(C, name='") == len(__import__('cales_b200.lib', fromlist=['SIGNATURES']).SIGNATURES)
    for ret, name, params in protos:
        m = re.search(r"function %s\((.*?)\) &\n\s+bind\(C, name='%s'\)" % (name, name), txt, re.S)
        assert m, name
        args = [a for a in re.sub(r"[&\s]", "", m.group(1)).split(",") if a]
        assert args == [p[1] for p in params], name
    assert max(len(l) for l in txt.splitlines()) <= 132               # free-form line limit


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_wrappers_call_existing_exports_with_matching_arity():
    arity = {name: len(params) for _, name, params in gen.prototypes()}
    src = open(os.path.join(ROOT, "fortran", "cales_b200_iface.f90")).read()
    src = "\n".join(l.split("!")[0] if not l.lstrip().startswith("!$acc") else "" for l in src.splitlines())
    src = re.sub(r"&\s*\n\s*", "", src)                               # join continuation lines
    called = set()
    for m in re.finditer(r"\b(cales_\w+)\(", src):
        name = m.group(1)
        if name in ("cales_b200_start", "cales_b200_stop", "cales_bound", "cales_b200_c", "cales_b200_iface"):
            continue
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = split_args(src[m.end():i - 1])
        assert name in arity, "wrapper calls unknown export %s" % name
        assert len(args) == arity[name], (name, len(args), arity[name])
        called.add(name)
    # every procedure main.f90 calls on the hot path has a wrapper
    for need in ("cales_init", "cales_get_decomp", "cales_initsolver", "cales_fftend", "cales_rk", "cales_bulk_forcing", "cales_bulk_mean",
                 "cales_bounduvw", "cales_boundp", "cales_cmpt_rhs_b", "cales_updt_rhs_b", "cales_fillps", "cales_solver",
                 "cales_solver_gaussel_z", "cales_correc", "cales_updatep", "cales_cmpt_sgs", "cales_chkdt", "cales_chkdiv",
                 "cales_finalize"):
        assert need in called, need
    assert max(len(l) for l in open(os.path.join(ROOT, "fortran", "cales_b200_iface.f90")).read().splitlines()) <= 132
