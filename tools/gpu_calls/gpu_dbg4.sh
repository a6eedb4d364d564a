#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_worker.py"
f() { grep -v "OMP_NUM\|\*\*\*\*\|^$" | tail -${1:-12} | cut -c1-700; }
echo "=== tgv 1x4 default"; timeout 120 $W tgv_smag 1 4 3 2>&1 | f 14
echo "=== tgv 1x4 halo unfused"; CALES_HALO_FUSED=0 timeout 120 $W tgv_smag 1 4 3 2>&1 | f 6
echo "=== tgv 1x4 pipe"; CALES_SOLVER_PIPE=1 timeout 120 $W tgv_smag 1 4 3 2>&1 | f 6
echo "=== tgv 1x4 pipe + halo unfused"; CALES_SOLVER_PIPE=1 CALES_HALO_FUSED=0 timeout 120 $W tgv_smag 1 4 3 2>&1 | f 6
echo "=== wm smag 2x2"; timeout 120 $W channel_wm_smag 2 2 5 2>&1 | f 6
echo "=== wm smag 4x1"; timeout 120 $W channel_wm_smag 4 1 5 2>&1 | f 6
