#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${1:-8}
run() { env TAG="$1" $2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 tools/nccl_allreduce_bench.py 2>gpurun_out/nccl_err.log | grep "|" ; }
run "default" "X=1"
NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29634 tools/nccl_allreduce_bench.py 2>&1 | grep -i "nvls\|algo\|channels\|Connected" | sort | uniq -c | sort -rn | head -12
run "NVLS" "NCCL_ALGO=NVLS"
run "Ring" "NCCL_ALGO=Ring"
run "Tree" "NCCL_ALGO=Tree"
run "minch32" "NCCL_MIN_NCHANNELS=32"
run "ctas32" "NCCL_MIN_CTAS=32"
