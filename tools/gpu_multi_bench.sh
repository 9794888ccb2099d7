#!/bin/bash
# Multi-GPU measurement call (run under `gpurun --gpus N`): headline bench, text-only training step (strong + weak),
# mixed multitask packing bench, each as one torchrun launch over the N GPUs of the box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-r01}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29601 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline" > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
timeout 300 bash -c "$(declare -f run); N=$N; run 29602 tools/bench_train.py --steps 30" > gpurun_out/train_strong_n${N}_$TAG.json 2> gpurun_out/train_strong_n${N}_$TAG.err
timeout 300 bash -c "$(declare -f run); N=$N; run 29603 tools/bench_train.py --steps 30 --weak" > gpurun_out/train_weak_n${N}_$TAG.json 2> gpurun_out/train_weak_n${N}_$TAG.err
timeout 300 bash -c "$(declare -f run); N=$N; run 29604 tools/bench_mixed.py --steps 10" > gpurun_out/mixed_n${N}_$TAG.json 2> gpurun_out/mixed_n${N}_$TAG.err
for f in bench_n${N}_$TAG train_strong_n${N}_$TAG train_weak_n${N}_$TAG mixed_n${N}_$TAG; do echo "== $f"; tail -c 1500 gpurun_out/$f.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/$f.err | tail -4; done
