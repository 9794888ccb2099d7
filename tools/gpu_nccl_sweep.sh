#!/bin/bash
# (run under `gpurun --gpus N`) NCCL settings sweep for the projector gradient message: all-reduce / reduce-scatter of
# 218 MB fp32 and 109 MB bf16 (tools/nccl_allreduce_bench.py); every setting is its own launch (NCCL reads its
# environment when the process group is created).  profiles/r02u_nccl_sweep_8gpu.txt
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${1:-8}
run() { env TAG="$1" $2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 tools/nccl_allreduce_bench.py 2>gpurun_out/nccl_err.log | grep "|" ; }
run "default" "X=1"
run "Ring" "NCCL_ALGO=Ring"
run "minch32" "NCCL_MIN_NCHANNELS=32"
run "ctas32" "NCCL_MIN_CTAS=32"
