"""Worker for the multi-GPU parity tests (launched by torch.distributed.run, one rank per GPU).
Runs the product on dims(1) x dims(2) ranks, gathers the fields on rank 0 and compares them with the
oracle's serial emulation of the same decomposition, then the device transposes on a global-index payload.
Usage: mgpu_worker.py CASE P_ROW P_COL NSTEPS [ARITH] [IMPDIFF]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    case, prow, pcol, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    arith = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] != "-" else None
    impdiff = sys.argv[6] if len(sys.argv) > 6 else None
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cales_b200 import lib as L
    import parity_mgpu as pm
    lib = L.load(arith)
    uid = pm.nccl_uid(lib, L, rank)
    name, kw = pm.CASES[case]
    rec = pm.case_vs_oracle(name, kw, (prow, pcol), nsteps, rank, world, local, uid, arith=arith, impdiff=impdiff)
    uid2 = pm.nccl_uid(lib, L, rank)
    tr = pm.transpose_round_trip((24, 20, 18), (prow, pcol), rank, world, local, uid2, arith=arith)
    ok = rec["ok"] and tr["ok"]
    if rank == 0:
        rec["transposes"] = tr
        rec["ok"] = ok
        print(json.dumps(rec))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
