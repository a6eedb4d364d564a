#!/usr/bin/env python
"""bench.py -- RK3 step throughput of the CaLES hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (libcales_b200.so through the C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]    reference arm: the C/OpenMP restatement of the
                                                                reference (oracle/c), all host cores

A step = one RK3 time step (3 substeps: momentum + SGS + pressure correction) of BASELINE config 2,
tri-periodic decaying Taylor-Green turbulence with the static Smagorinsky model, 256^3 per GPU
(weak scaling: the z extent grows with N).  Prints ONE JSON line on rank 0.
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


def cpu_port_rate(ng, steps, warmup):
    """Mcell-updates/s of the CPU restatement of the reference (oracle/c: C + OpenMP, all host threads, the same loops
    as the Fortran) on the bench workload.  Returns (rate, seconds/step, threads)."""
    import oracle.param as op
    from oracle.cport import CSim
    s = CSim(op.deck_tgv(ng=ng))
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
    cannot be produced in this image (no gfortran/MPI/FFTW), so this is the C/OpenMP port in oracle/c ("kind": "port"),
    on the SAME workload as our arm (256^3 TGV, static Smagorinsky), every timed step one full RK3 step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ng = (256, 256, 256)
    steps = max(1, args.steps)
    val, dt, th = cpu_port_rate(ng, steps, max(1, min(args.warmup, 2)))
    sample = "TGV smag %dx%dx%d (the full bench grid), %d RK3 steps, C/OpenMP port of the reference loops, %d threads" % (ng + (steps, th))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: tri-periodic decaying turbulence (TGV init), static Smagorinsky, "
                                   "%dx%dx%d, explicit diffusion" % ng,
                       "grid": list(ng), "note": "restated CPU path (the Fortran/MPI/FFTW reference cannot be built in this image); one rank, "
                                                 "all host threads, independent of --gpus"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def time_kernel(fn, iters=10):
    import torch
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cales_b200 import lib as L
    from cales_b200.deck import deck_tgv
    from cales_b200.driver import Simulation
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lib = L.load()
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(128)
            L.check(None, lib.cales_get_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    nloc = (256, 256, 256)
    ng = (nloc[0], nloc[1], nloc[2] * world)              # weak scaling: z slabs, dims = (1, N)
    deck = deck_tgv(ng=ng, dims=(1, world))
    sim = Simulation(deck, rank=rank, nranks=world, uid=uid, device=local)
    sim.init_flow()
    sim.start()
    ncell_loc = float(np.prod(sim.n)); ncell = float(np.prod(ng))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
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
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_step = t.item() / args.steps
    # ---- Poisson solve alone -------------------------------------------------------------------------------
    t_poi = time_kernel(lambda: sim.solver(sim.poi, "pp"), 10)
    # ---- e2e: host buffers; every step uploads its inputs (u,v,w,p) from pinned host memory and downloads its
    # results (u,v,w,p) to pinned host memory.  The copies are pipelined the way a production host would drive
    # them: the upload of step s+1 (copy-in stream) and the download of step s-1 (copy-out stream) overlap the
    # compute of step s on the library's stream, through double-buffered device staging; PCIe is full duplex.
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
    te = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = te.item() / ke
    del stage_in, stage_out
    # ---- roofline of the dominant kernels (live CUDA-event timing on the launch stream) -------------------------
    import ctypes as C
    n = sim.n; d = deck
    scr = [torch.zeros(int(ncell_loc), dtype=torch.float64, device="cuda") for _ in range(3)]
    D = sim.d
    kern = {}
    kern["mom_xyz_ad"] = (56, time_kernel(lambda: sim.chk(sim.lib.cales_mom_xyz_ad(
        sim.ctx, L._ia(n), d.dli[0], d.dli[1], D["dzci"].data_ptr(), D["dzfi"].data_ptr(), d.visc, sim.ptr("u"), sim.ptr("v"), sim.ptr("w"),
        sim.ptr("visct"), scr[0].data_ptr(), scr[1].data_ptr(), scr[2].data_ptr(), None, None, None))))
    wk = torch.zeros(int(ncell_loc), dtype=torch.float64, device="cuda")
    nn = L._ia(n)
    kern["fft_x_fwd"] = (16, time_kernel(lambda: sim.chk(sim.lib.cales_fft_lines(sim.ctx, nn, 0, b"PP", b"c", 0, wk.data_ptr()))))
    kern["fft_y_fwd"] = (16, time_kernel(lambda: sim.chk(sim.lib.cales_fft_lines(sim.ctx, nn, 1, b"PP", b"c", 0, wk.data_ptr()))))
    if world == 1:
        kern["gaussel_periodic"] = (16, time_kernel(lambda: sim.chk(sim.lib.cales_gaussel(
            sim.ctx, int(n[0]), int(n[1]), int(n[2]), 1, sim.poi["a"].data_ptr(), sim.poi["b"].data_ptr(), sim.poi["c"].data_ptr(),
            sim.poi["lam"].data_ptr(), wk.data_ptr()))))
    kern["fillps"] = (32, time_kernel(lambda: sim.fillps(1.0)))
    kern["correc"] = (56, time_kernel(lambda: sim.correc(0.0)))
    kern["cmpt_sgs_smag"] = (32, time_kernel(lambda: sim.cmpt_sgs()))
    peak, peak_src = peaks()
    kinfo = {k: {"alg_bytes_per_cell": b, "ms": tt * 1e3, "achieved_gbs": b * ncell_loc / tt / 1e9, "frac": b * ncell_loc / tt / 1e9 / peak}
             for k, (b, tt) in kern.items()}
    dom = "mom_xyz_ad"
    line = None
    if rank == 0:
        # CPU baseline (oracle port) on a bounded sample, N=1 only
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            val, dtc, th = cpu_port_rate(nloc, 5, 1)
            cpu = {"value": val, "unit": UNIT, "cores": th, "kind": "port",
                   "sample": "the same 256x256x256 workload, 5 RK3 steps, C/OpenMP port of the reference loops (oracle/c), "
                             "%d threads (%.2f s/step)" % (th, dtc)}
        line = {"metric": METRIC, "value": ncell / t_step / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "BASELINE config 2: tri-periodic decaying turbulence (TGV init), static Smagorinsky, "
                                       "%dx%dx%d per GPU, explicit diffusion" % nloc,
                           "grid": list(ng), "dims": [1, world], "l2": "working set (6 fields x 134 MB + scratch) exceeds the 126 MB L2; no flush needed",
                           "parity_mode": "-fmad=false"},
                "poisson_ms": t_poi * 1e3,
                "e2e": {"value": ncell / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                        "ms_per_step": t_e2e * 1e3, "what": "every step: pinned host u,v,w,p -> device, one RK3 step through the C ABI, u,v,w,p -> pinned host; uploads/downloads pipelined on copy streams (double-buffered staging), timed until the last result is on the host"},
                "gpu_launches": int(n1 - n0) if lc else None,
                "roofline": {"kernel": dom, "bound": "hbm", "achieved": kinfo[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": kinfo[dom]["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
                             "alg_bytes_per_launch": kinfo[dom]["alg_bytes_per_cell"] * ncell_loc},
                "kernels": kinfo,
                "clocks": sampler.summary()}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
