"""Multi-GPU parity: the pencil-decomposed run (halo exchange, NCCL all-to-all transposes, all-reduced plane
averages) against the oracle's serial emulation of the same decomposition.  Needs >= 2 GPUs on the box; on a
single-GPU box these tests are skipped (the world_size-2 host logic is covered on CPU by test_decomp_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(case, prow, pcol, nsteps, port, arith="-", impdiff=None, env=None):
    n = prow * pcol
    from conftest import need_gpu
    need_gpu()
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs, box has %d" % (n, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), case, str(prow), str(pcol), str(nsteps), arith] + ([impdiff] if impdiff else [])
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not (r.returncode == 0 and lines and json.loads(lines[-1])["ok"]):
        err = [l for l in r.stderr.splitlines() if l.strip() and "OMP_NUM_THREADS" not in l and "****" not in l]
        pytest.fail("rc=%d\n%s\n%s" % (r.returncode, "\n".join(lines[-1:])[:1500], "\n".join(err[-25:])[:3000]), pytrace=False)


@pytest.mark.parametrize("case,prow,pcol", [("channel_dsmag", 1, 2), ("channel_dsmag", 2, 1), ("tgv_smag", 1, 2), ("channel_wm_smag", 2, 1),
                                            ("duct_smag", 1, 2), ("cavity_smag", 2, 1)])
def test_two_gpus(case, prow, pcol):
    run(case, prow, pcol, 5, 29511)


@pytest.mark.parametrize("case,prow,pcol", [("channel_dsmag", 2, 2), ("tgv_smag", 1, 4), ("channel_wm_dsmag", 4, 1)])
def test_four_gpus(case, prow, pcol):
    run(case, prow, pcol, 5, 29512)


@pytest.mark.parametrize("case,prow,pcol", [("channel_dsmag", 1, 8), ("channel_dsmag", 2, 4), ("channel_wm_smag", 4, 2), ("tgv_smag", 1, 8)])
def test_eight_gpus(case, prow, pcol):
    run(case, prow, pcol, 5, 29513)


@pytest.mark.parametrize("case,prow,pcol,env", [("tgv_smag", 1, 2, {"CALES_ZDIST": "0", "CALES_SOLVER_PIPE": "1"}),
                                                ("channel_dsmag", 1, 2, {"CALES_ZDIST": "0", "CALES_SOLVER_PIPE": "1", "CALES_SOLVER_CHUNKS": "3"}),
                                                ("channel_smag", 1, 2, {"CALES_ZDIST": "0", "CALES_SOLVER_PIPE": "0"}), ("tgv_smag", 1, 2, {"CALES_ZDIST": "0"}),
                                                ("duct_smag", 1, 2, {"CALES_HALO_FUSED": "0"}),
                                                ("channel_dsmag", 1, 2, {"CALES_NO_P2P": "1"}), ("tgv_smag", 1, 2, {"CALES_B200_ARITH": "strict"}),
                                                ("tgv_smag", 1, 2, {"CALES_B200_ARITH": "strict", "CALES_ZDIST": "1"}),
                                                ("channel_wm_smag", 1, 2, {"CALES_B200_ARITH": "strict", "CALES_ZDIST": "1"})])
def test_two_gpus_exchange_variants(case, prow, pcol, env):
    """every exchange mechanism on the same cases: the distributed z solve is the product build's default wherever z is
    decomposed (all other tests of this file); here the transposing paths it replaced (copy-engine pipeline, kernel-fused
    transposes), the separate-kernel halo exchange, NCCL send/recv without peer memory, the strict arithmetic build (bit-exact
    Thomas through the transposes) and the strict build with the distributed z solve forced on"""
    run(case, prow, pcol, 5, 29514, env=env)


@pytest.mark.parametrize("case,prow,pcol,mode", [("channel_smag", 1, 2, "1d"), ("channel_wm_smag", 2, 1, "1d"), ("tgv_smag", 1, 2, "1d"), ("channel_smag", 1, 2, "3d")])
def test_two_gpus_implicit_diffusion(case, prow, pcol, mode):
    """Crank-Nicolson across ranks: _IMPDIFF_1D with z decomposed runs solver_gaussel_z through the pencil transposes
    (src/solver.f90:199-231); _IMPDIFF runs three Helmholtz solves per substep through the distributed solver"""
    run(case, prow, pcol, 5, 29515, impdiff=mode)
