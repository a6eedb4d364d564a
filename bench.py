#!/usr/bin/env python
"""bench.py -- RK3 step throughput of the CaLES hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (libcales_b200.so through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]    reference arm: the C/OpenMP restatement of the
                                                                reference (oracle/c), all host cores

Default workload (the one the driver records): BASELINE config 2, tri-periodic decaying Taylor-Green turbulence with the
static Smagorinsky model, 256^3 per GPU (weak scaling: the z extent grows with N, dims = 1 x N).  A step = one RK3 time
step (3 substeps: momentum + SGS + pressure correction).  Prints ONE JSON line on rank 0.

Other BASELINE configurations, for the tables in DESIGN.md / profiles/ (strong scaling: the grid is fixed, N ranks share it):
  --workload channel1   config 1: periodic channel Re_tau~180, 64^3, dynamic Smagorinsky
  --workload channel3   config 3: wall-modelled channel Re_tau~2000, 512x256x192
  --workload duct4 | cavity4   config 4: square duct / lid-driven cavity 512x256x256
  --workload channel5   config 5: channel LES 1024x512x512
  --sgs smag|dsmag  --dims P Q  --grid NX NY NZ  --arith fma|strict override their defaults.

With WORLD_SIZE > 1 the run first checks parity on the live process group (tests/parity_mgpu.py: short runs against the
oracle's emulation of the same decomposition + the device transposes on a global-index payload) and exits non-zero
when it fails; the verdict is part of the JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "rk3_step_throughput"
UNIT = "Mcell-updates/s"
NVLINK_PEAK = 900.0   # GB/s per direction per GPU (NVLink 5)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json: {"kernel": bytes, "_source": ...}); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks/throttle reasons during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def bind_near_gpu(index):
    """Restrict the calling thread to the CPUs NVML reports as local to GPU `index` (its NUMA node), so that the pinned
    host buffers of the end-to-end leg are allocated next to the PCIe root the GPU hangs off (first touch follows the
    allocating thread).  Returns (previous mask, record for the JSON line) or (None, reason); never raises."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * wi + b for wi, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cur = os.sched_getaffinity(0)
        new = near & cur
        if not new:
            return None, "NVML's CPU set for the GPU does not intersect this process's CPUs"
        if new == cur:
            return None, "already local (%d CPUs)" % len(cur)
        os.sched_setaffinity(0, new)
        return cur, "host thread bound to the %d of %d CPUs local to the GPU while the pinned buffers are allocated and the copies driven" % (len(new), len(cur))
    except Exception as e:                                   # no NVML, no permission, ...: run unbound
        return None, "unbound (%s)" % type(e).__name__


# ---- workloads --------------------------------------------------------------------------------------------------------
def workload(args, world):
    """-> (deck constructor name, kwargs, global grid, dims, label, scaling)"""
    w = args.workload
    dims = tuple(args.dims) if args.dims else None
    if w == "tgv256":
        nloc = tuple(args.grid) if args.grid else (256, 256, 256)
        dims = dims or (1, world)
        ng = (nloc[0], nloc[1] * dims[0], nloc[2] * dims[1])
        return ("deck_tgv", dict(ng=ng, sgstype=args.sgs or "smag"), ng, dims,
                "BASELINE config 2: tri-periodic decaying turbulence (TGV init), static Smagorinsky, %dx%dx%d per GPU, explicit diffusion" % nloc, "weak")
    if w == "channel1":
        ng = tuple(args.grid) if args.grid else (64, 64, 64)
        return ("deck_channel", dict(ng=ng, sgstype=args.sgs or "dsmag"), ng, dims or (1, world),
                "BASELINE config 1: periodic channel Re_tau~180, %dx%dx%d, %s" % (ng + (args.sgs or "dsmag",)), "strong")
    if w in ("channel3", "channel5"):
        ng = tuple(args.grid) if args.grid else ((512, 256, 192) if w == "channel3" else (1024, 512, 512))
        sgs = args.sgs or "smag"
        wm = not args.no_wall_model
        if dims is None:          # smag + z walls: at most two ranks across z (sanity.f90:98-111)
            dims = (1, world) if (sgs == "dsmag" or world <= 2) else (world // 2, 2)
        return ("deck_channel", dict(ng=ng, sgstype=sgs, wall_model=wm, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), ng, dims,
                "BASELINE config %s: %schannel Re_tau~2000, %dx%dx%d, %s%s" % ("3" if w == "channel3" else "5", "wall-modelled " if wm else "",
                                                                              ng[0], ng[1], ng[2], sgs, " (log-law wall model)" if wm else ""), "strong")
    if w in ("duct4", "cavity4"):
        ng = tuple(args.grid) if args.grid else (512, 256, 256)
        if dims is None:
            dims = (2, world // 2) if world >= 4 else (1, world)
        name = "deck_duct" if w == "duct4" else "deck_cavity"
        return (name, dict(ng=ng, sgstype=args.sgs or "smag"), ng, dims,
                "BASELINE config 4: %s, %dx%dx%d, %s" % ("square duct" if w == "duct4" else "lid-driven cavity", ng[0], ng[1], ng[2], args.sgs or "smag"), "strong")
    raise SystemExit("unknown workload " + w)


# ---- reference arm -------------------------------------------------------------------------------------------------------
def cpu_port_rate(ng, steps, warmup, threads=None, deck=None):
    """Mcell-updates/s of the CPU restatement of the reference (oracle/c: C + OpenMP, the same loops as the Fortran) on the
    bench workload (default deck: TGV smag on grid `ng`), on `threads` OpenMP threads (default: every core this process may
    run on -- set explicitly, because torchrun exports OMP_NUM_THREADS=1 to its workers).  Returns (rate, seconds/step, threads)."""
    import oracle.param as op
    from oracle.cport import CSim
    threads = threads or len(os.sched_getaffinity(0))
    s = CSim(deck if deck is not None else op.deck_tgv(ng=ng), threads=threads)
    for _ in range(warmup):
        s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    dt = (time.perf_counter() - t0) / steps
    th = s.threads()
    s.close()
    return float(np.prod(ng)) / dt / 1e6, dt, th


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for this path on the host cores.  The Fortran/MPI/FFTW build
    cannot be produced in this image (no gfortran/MPI/FFTW), so this is the C/OpenMP port in oracle/c ("kind": "port").
    Workload = our arm's: 256^3 TGV smag per GPU, i.e. the grid 256 x 256 x 256N at --gpus N (one process, all host
    threads; rank 0 only under torchrun).  Every timed step is one full RK3 step of that grid; when K steps would take
    longer than ~150 s the number of timed steps is reduced and said so in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.param as op
    from oracle.cport import CSim
    n = max(1, args.gpus)
    name, kw, ng, dims, label, scaling = workload(args, n)       # our arm's workload at N GPUs: the same global grid
    kw = dict(kw); kw["ng"] = tuple(ng)
    deck = getattr(op, name)(**kw)
    if CSim.kind(deck) is None:
        print(json.dumps({"impl": "reference", "unavailable": "the C/OpenMP port (oracle/c) covers static-Smagorinsky tri-periodic and plane-channel decks; "
                                                              "not %s" % label}))
        return
    threads = len(os.sched_getaffinity(0))
    est = 0.3 * (float(np.prod(ng)) / 256 ** 3) * 16.0 / max(threads, 1)   # s/step guess: 0.3 s per 256^3 on 16 threads
    steps = max(1, min(args.steps, int(150.0 / max(est, 1e-3))))
    val, dt, th = cpu_port_rate(ng, steps, max(1, min(args.warmup, 2)), threads, deck=deck)
    sample = "%s, grid %dx%dx%d (the full grid of our arm at %d GPU%s), %d timed RK3 steps, C/OpenMP port of the reference loops, %d threads" % (
        (label,) + tuple(ng) + (n, "" if n == 1 else "s", steps, th))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label,
                       "grid": list(ng), "timed_steps": steps,
                       "note": "restated CPU path (the Fortran/MPI/FFTW reference cannot be built in this image); one process, all host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- our arm -----------------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=10):
    import torch
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def parity_check(args, dims, rank, world, local, lib, L):
    """Short runs on THIS process group against the oracle's emulation of the same decomposition (rank 0 runs the oracle),
    and the device transposes on a global-index payload.  Returns the record that goes into the JSON line."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_mgpu as pm
    nz = max(32, 8 * dims[1]); ny = max(24, 8 * dims[0])
    cases = [("deck_tgv", dict(ng=(32, max(32, ny), nz)), 5),
             ("deck_channel", dict(ng=(32, ny, nz), sgstype="dsmag"), 5),
             ("deck_channel", dict(ng=(32, ny, nz), sgstype="dsmag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), 3)]
    recs = []
    for name, kw, nsteps in cases:
        uid = pm.nccl_uid(lib, L, rank)
        recs.append(pm.case_vs_oracle(name, kw, dims, nsteps, rank, world, local, uid, arith=args.arith))
    uid = pm.nccl_uid(lib, L, rank)
    tr = pm.transpose_round_trip((32, max(24, 4 * dims[0]), max(24, 4 * dims[1])), dims, rank, world, local, uid, arith=args.arith)
    ok = all(r["ok"] for r in recs) and tr["ok"]
    return {"ok": bool(ok), "dims": list(dims), "against": "numpy oracle emulating the same decomposition (tests/parity_mgpu.py), tolerance 1e-10",
            "cases": [{k: r.get(k) for k in ("case", "ng", "steps", "solver_exchange", "errs", "divmax", "ok")} for r in recs], "transposes": tr}


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from cales_b200 import lib as L
    import cales_b200.deck as pd
    from cales_b200.deck import rkcoeff
    from cales_b200.driver import Simulation
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    lib = L.load(args.arith)
    name, kw, ng, dims, label, scaling = workload(args, world)
    if dims[0] * dims[1] != world:
        raise SystemExit("bench.py: dims %s do not match %d ranks" % (dims, world))
    uid = None
    parity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity_mgpu as pm
        if not args.no_parity_check:
            parity = parity_check(args, dims, rank, world, local, lib, L)
            if not parity["ok"] and os.environ.get("CALES_ZDIST") is None:
                # safety net: the transposing solver paths are the fallback of the distributed z solve; say so in the line
                os.environ["CALES_ZDIST"] = "0"
                first = parity
                parity = parity_check(args, dims, rank, world, local, lib, L)
                parity["zdist_disabled_after_failed_check"] = {k: first.get(k) for k in ("cases", "transposes")}
            if not parity["ok"]:
                if rank == 0:
                    print(json.dumps({"metric": METRIC, "n_gpus": world, "parity_check": parity, "error": "multi-GPU parity check failed; nothing timed"}))
                dist.destroy_process_group()
                sys.exit(1)
        uid = pm.nccl_uid(lib, L, rank)
    deck = getattr(pd, name)(dims=dims, **kw)
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local, arith=args.arith)

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return t.item()
    sim.init_flow(mean_allreduce=allsum)
    sim.start()
    ncell_loc = float(np.prod(sim.n)); ncell = float(np.prod(ng))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    for _ in range(args.warmup):
        sim.step()
    # ---- timed region: K steps, device events, max over ranks --------------------------------------------
    sampler = ClockSampler(local); sampler.start()
    lc = getattr(sim.lib, "cales_launch_count", None)
    barrier()
    n0 = lc(sim.ctx) if lc else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sim.step()
    e1.record()
    barrier()
    n1 = lc(sim.ctx) if lc else 0
    sampler.stop_flag = True
    t_step = maxranks(e0.elapsed_time(e1) * 1e-3) / args.steps
    # ---- sanity of the state the timed steps produced: finite fields, divergence at round-off ----------------
    divtot, divmax = sim.chkdiv()
    finite = maxranks(0.0 if all(bool(torch.isfinite(sim.fields[c]).all().item()) for c in "uvwp") else 1.0) == 0.0
    umax = maxranks(float(sim.fields["u"].abs().max().item()))
    sanity = {"finite": bool(finite), "divmax": divmax, "umax": umax, "ok": bool(finite and divmax < 1e-9 * max(1.0, umax) * float(max(deck.dli)))}
    # ---- Poisson solve alone (max over ranks) ---------------------------------------------------------------
    barrier()
    t_poi = maxranks(time_kernel(lambda: sim.solver(sim.poi, "pp"), 10))
    # ---- per-phase table of one substep (each phase timed alone: ranks barriered, device synchronised) ----------
    dtrk = (rkcoeff[1][0] + rkcoeff[1][1]) * sim.dt
    d = deck

    def phase(fn, iters=5):
        fn(); barrier()
        tot = 0.
        for _ in range(iters):
            barrier()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return maxranks(tot / iters)
    phases = {}
    if not args.no_phases:
        for nm, fn in (("rk", lambda: sim.rk(1, want_f=False)), ("bounduvw", lambda: sim.bounduvw(True, False)), ("fillps", lambda: sim.fillps(1. / dtrk)),
                       ("solver", lambda: sim.solver(sim.poi, "pp")), ("boundp", lambda: sim.boundp(d.cbcpre, sim.bcp, "pp")),
                       ("correc", lambda: sim.correc(0.0)), ("updatep", lambda: sim.updatep()), ("cmpt_sgs", lambda: sim.cmpt_sgs())):
            phases[nm] = phase(fn)
    # ---- NVLink: the y -> z exchange of the solver alone (every rank stores (P-1)/P of its pencil into its peers) --------
    nvlink = None
    if world > 1 and dims[1] > 1:
        ysz = [int(x) for x in sim.n_y_fft]
        nmax = int(max(np.prod(sim.n_x_fft), np.prod(sim.n_y_fft), np.prod(sim.n_z)) * 1.1) + 1024
        src, dst = C.c_void_p(), C.c_void_p()
        sim.chk(sim.lib.cales_peer_alloc(sim.ctx, b"bench_src", nmax * 8, C.byref(src)))
        sim.chk(sim.lib.cales_peer_alloc(sim.ctx, b"bench_dst", nmax * 8, C.byref(dst)))
        t_tr = phase(lambda: sim.chk(sim.lib.cales_transpose(sim.ctx, 1, src, dst)), 10)
        sent = 8.0 * float(np.prod(ysz)) * (dims[1] - 1) / dims[1]
        exch = sim.lib.cales_solver_exchange(sim.ctx).decode()
        nvlink = {"what": "y->z pencil transpose alone (one kernel storing each sub-box into its owner's Z-pencil over NVLink + flag barrier), max over ranks",
                  "bytes_sent_per_gpu": sent, "ms": t_tr, "gbs_per_direction": sent / t_tr / 1e6, "peak": NVLINK_PEAK,
                  "frac_of_900": sent / t_tr / 1e6 / NVLINK_PEAK, "solver_exchange": exch, "poisson_ms": t_poi * 1e3}
        if exch.startswith("distributed z solve"):
            # the pressure solve no longer transposes y<->z: each rank pushes the first and last plane of its block to its peers
            nvlink["bytes_per_solve_per_gpu"] = 2 * 8.0 * ysz[0] * ysz[1] * (dims[1] - 1)
            nvlink["bytes_per_solve_per_gpu_transposing"] = 2 * sent
        else:
            nvlink["exchanges_per_solve"] = 2
            nvlink["bytes_per_solve_per_gpu"] = 2 * sent
    # ---- e2e: host buffers; every step uploads its inputs (u,v,w,p) from pinned host memory and downloads its
    # results (u,v,w,p) to pinned host memory.  The copies are pipelined the way a production host would drive
    # them: the upload of step s+1 (copy-in stream) and the download of step s-1 (copy-out stream) overlap the
    # compute of step s on the library's stream, through double-buffered device staging; PCIe is full duplex.
    e2e = None
    if not args.no_e2e:
        torch.ones(1 << 22).sum().item()      # host thread pool (shared OpenMP runtime) comes up with the process's full CPU mask first
        old_mask, numa_note = bind_near_gpu(local)
        names = ("u", "v", "w", "p")
        hin = {nm: torch.empty(sim.ncell, dtype=torch.float64).pin_memory() for nm in names}
        hout = {nm: torch.empty(sim.ncell, dtype=torch.float64).pin_memory() for nm in names}
        for nm in names:
            hin[nm].copy_(sim.fields[nm])
        nbytes = sum(hin[nm].numel() * 8 for nm in names)
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        stage_in = [{nm: torch.empty_like(sim.fields[nm]) for nm in names} for _ in range(2)]
        stage_out = [{nm: torch.empty_like(sim.fields[nm]) for nm in names} for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]       # upload into stage_in[b] finished
        ev_in_free = [torch.cuda.Event() for _ in range(2)]  # stage_in[b] consumed by the compute stream
        ev_out = [torch.cuda.Event() for _ in range(2)]      # stage_out[b] filled by the compute stream
        ev_out_free = [torch.cuda.Event() for _ in range(2)] # download of stage_out[b] finished

        def upload(sidx):
            b = sidx % 2
            with torch.cuda.stream(s_in):
                if sidx >= 2:
                    s_in.wait_event(ev_in_free[b])
                for nm in names:
                    stage_in[b][nm].copy_(hin[nm], non_blocking=True)
                ev_in[b].record(s_in)

        def e2e_run(k):
            upload(0)
            for sidx in range(k):
                b = sidx % 2
                if sidx + 1 < k:
                    upload(sidx + 1)
                main.wait_event(ev_in[b])
                for nm in names:
                    sim.fields[nm].copy_(stage_in[b][nm], non_blocking=True)
                ev_in_free[b].record(main)
                sim.step()
                if sidx >= 2:
                    main.wait_event(ev_out_free[b])
                for nm in names:
                    stage_out[b][nm].copy_(sim.fields[nm], non_blocking=True)
                ev_out[b].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_out[b])
                    for nm in names:
                        hout[nm].copy_(stage_out[b][nm], non_blocking=True)
                    ev_out_free[b].record(s_out)
            main.wait_stream(s_out)                          # the timed region ends when the last result is on the host

        e2e_run(2); barrier()
        ke = max(4, min(args.steps, 10))
        e0.record()
        e2e_run(ke)
        e1.record()
        barrier()
        t_e2e = maxranks(e0.elapsed_time(e1) * 1e-3) / ke
        del stage_in, stage_out
        if old_mask is not None:
            os.sched_setaffinity(0, old_mask)
        e2e = {"value": ncell / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": t_e2e * 1e3,
               "what": "every step: pinned host u,v,w,p -> device, one RK3 step through the C ABI, u,v,w,p -> pinned host; uploads/downloads "
                       "pipelined on copy streams (double-buffered staging), timed until the last result is on the host; bytes are per rank",
               "numa": numa_note}
    # ---- roofline of the kernels (live CUDA-event timing on the launch stream) -------------------------
    n = sim.n
    scr = [torch.zeros(int(ncell_loc), dtype=torch.float64, device="cuda") for _ in range(3)]
    D = sim.d
    kern = {}
    kern["mom_xyz_ad"] = (56, time_kernel(lambda: sim.chk(sim.lib.cales_mom_xyz_ad(
        sim.ctx, L._ia(n), d.dli[0], d.dli[1], D["dzci"].data_ptr(), D["dzfi"].data_ptr(), d.visc, sim.ptr("u"), sim.ptr("v"), sim.ptr("w"),
        sim.ptr("visct"), scr[0].data_ptr(), scr[1].data_ptr(), scr[2].data_ptr(), None, None, None))))
    if not deck.impdiff:
        hs = [torch.zeros(sim.ncell, dtype=torch.float64, device="cuda") for _ in range(3)]
        kern["mom_rk_fused"] = (112, time_kernel(lambda: sim.chk(sim.lib.cales_rk_fused(
            sim.ctx, L._da(rkcoeff[1]), L._ia(n), L._da(d.dli), D["dzci"].data_ptr(), D["dzfi"].data_ptr(), D["gvr_c"].data_ptr(), D["gvr_f"].data_ptr(),
            d.visc, sim.dt, sim.ptr("p"), L._ia(np.zeros(3, dtype=np.int32)), L._da(d.velf), L._da(d.bforce), sim.ptr("visct"),
            sim.ptr("u"), sim.ptr("v"), sim.ptr("w"), hs[0].data_ptr(), hs[1].data_ptr(), hs[2].data_ptr()))))
    if world == 1:
        wk = torch.zeros(int(ncell_loc), dtype=torch.float64, device="cuda")
        nn = L._ia(n)
        bcx = (deck.cbcpre[0, 0] + deck.cbcpre[1, 0]).encode(); bcy = (deck.cbcpre[0, 1] + deck.cbcpre[1, 1]).encode()
        for dname, dir_, bc in (("x", 0, bcx), ("y", 1, bcy)):
            for bw in (0, 1):
                kern["fft_%s_%s" % (dname, "bwd" if bw else "fwd")] = (16, time_kernel(
                    lambda: sim.chk(sim.lib.cales_fft_lines(sim.ctx, nn, dir_, bc, b"c", bw, wk.data_ptr()))))
        zper = deck.cbcpre[0, 2] == "P"
        kern["gaussel_periodic" if zper else "gaussel"] = (16, time_kernel(lambda: sim.chk(sim.lib.cales_gaussel(
            sim.ctx, int(n[0]), int(n[1]), int(n[2]), 1 if zper else 0, sim.poi["a"].data_ptr(), sim.poi["b"].data_ptr(), sim.poi["c"].data_ptr(),
            sim.poi["lam"].data_ptr(), wk.data_ptr()))))
    kern["fillps"] = (32, time_kernel(lambda: sim.fillps(1.0)))
    kern["correc"] = (56, time_kernel(lambda: sim.correc(0.0)))
    kern["updatep"] = (24, time_kernel(lambda: sim.updatep()))
    sgs_bytes = 32 if deck.sgstype.strip() == "smag" else 64
    kern["cmpt_sgs_" + deck.sgstype.strip()] = (sgs_bytes, time_kernel(lambda: sim.cmpt_sgs()))
    kern["rk(mom+update+forcing)"] = (160, time_kernel(lambda: sim.rk(1, want_f=False)))
    peak, peak_src = peaks()
    kinfo = {k: {"alg_bytes_per_cell": b, "ms": tt * 1e3, "achieved_gbs": b * ncell_loc / tt / 1e9, "frac": b * ncell_loc / tt / 1e9 / peak}
             for k, (b, tt) in kern.items()}
    single = [k for k in kinfo if not k.startswith("rk(") and not k.startswith("cmpt_sgs_dsmag")]
    dom = max(single, key=lambda k: kinfo[k]["ms"])          # the dominant single kernel of the step (every one runs 3x per step)
    ok = sanity["ok"]
    if rank == 0:
        # CPU baseline (oracle port) on a bounded sample, N=1 only
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle.param as op
            from oracle.cport import CSim
            okw = dict(kw); okw["ng"] = tuple(ng)
            odeck = getattr(op, name)(**okw)
            if CSim.kind(odeck) is not None:
                nst = 5 if ncell <= 2.0e7 else 2
                val, dtc, th = cpu_port_rate(ng, nst, 1, deck=odeck)
                cpu = {"value": val, "unit": UNIT, "cores": th, "kind": "port",
                       "sample": "the same %dx%dx%d workload, %d RK3 steps, C/OpenMP port of the reference loops (oracle/c), "
                                 "%d threads (%.2f s/step)" % (tuple(ng) + (nst, th, dtc))}
        line = {"metric": METRIC, "value": ncell / t_step / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": label, "grid": list(ng), "dims": list(dims),
                           "l2": "working set (6 fields + 6 RK arrays + scratch, %.0f MB per field per GPU) exceeds the 126 MB L2; no flush needed" % (ncell_loc * 8 / 1e6)
                                 if ncell_loc * 8 * 12 > 2 * 126e6 else "working set fits the 126 MB L2 (launch-bound case); steps run back to back on live data",
                           "arith": "%s (%s)" % (sim.lib.arith, "fp64 contraction on, as the reference's GPU build; tolerance parity 1e-12/1e-10, tests/" if sim.lib.arith == "fma"
                                                 else "-fmad=false, bit-identical stencils")},
                "poisson_ms": t_poi * 1e3,
                "gpu_launches": int(n1 - n0) if lc else None,
                "sanity": sanity,
                "roofline": {"kernel": dom, "bound": "hbm", "achieved": kinfo[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": kinfo[dom]["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
                             "alg_bytes_per_launch": kinfo[dom]["alg_bytes_per_cell"] * ncell_loc,
                             "step_frac_of_fused_bound": 330.0 * 3 * ncell_loc / t_step / 1e9 / peak if deck.sgstype.strip() == "smag" else None},
                "kernels": kinfo,
                "phases_ms": phases,
                "clocks": sampler.summary()}
        if e2e:
            line["e2e"] = e2e
        if nvlink:
            line["nvlink"] = nvlink
        if parity:
            line["parity_check"] = parity
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tgv256", choices=["tgv256", "channel1", "channel3", "duct4", "cavity4", "channel5"])
    ap.add_argument("--sgs", default=None, choices=["smag", "dsmag", "none"])
    ap.add_argument("--dims", type=int, nargs=2, default=None)
    ap.add_argument("--grid", type=int, nargs=3, default=None, help="tgv256: per-GPU grid; others: global grid")
    ap.add_argument("--arith", default=None, choices=["fma", "strict"], help="library variant (default: the product build, fma)")
    ap.add_argument("--no-wall-model", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-phases", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
