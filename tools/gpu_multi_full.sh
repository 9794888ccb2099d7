#!/bin/bash
# multi-GPU validation (run under `gpurun --gpus N`, N >= 2): whole GPU test suite incl. the NCCL tests, then the headline
# bench with its packed-path / training-step sections at 1 and N GPUs of the same box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r02c}
: > gpurun_out/rc_$TAG.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -v -m gpu --timeout 500 > gpurun_out/t_multi_$TAG.log 2>&1; echo "multi rc=$?" >> gpurun_out/rc_$TAG.txt
tail -8 gpurun_out/t_multi_$TAG.log
if [ "${SKIP_SUITE:-0}" != 1 ]; then
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_multi.py > gpurun_out/t_suite_$TAG.log 2>&1; echo "suite rc=$?" >> gpurun_out/rc_$TAG.txt
tail -12 gpurun_out/t_suite_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench1 rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "benchN rc=$?" >> gpurun_out/rc_$TAG.txt
grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$TAG.err | tail -8
# host limit of the end-to-end path: pinned H2D + D2H on all N GPUs at once (one process per GPU, CPU slices as in bench.py)
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i timeout 120 python tools/micro/pcie_duplex.py > gpurun_out/pcie_${TAG}_gpu$i.json 2>/dev/null &
done
wait
python - $N $TAG <<'PY'
import json, sys
N, TAG = int(sys.argv[1]), sys.argv[2]
tot = {"h2d_gbs": 0.0, "d2h_gbs": 0.0, "duplex_total_gbs": 0.0}
for i in range(N):
    try:
        d = json.load(open("gpurun_out/pcie_%s_gpu%d.json" % (TAG, i)))["66MB"]
        for k in tot:
            tot[k] += d[k]
    except Exception as e:
        print("pcie gpu %d failed: %s" % (i, e))
print("PCIe, %d GPUs at once (sum over GPUs; the phases of the processes are not aligned, so this is a lower bound of the contention):" % N, json.dumps(tot))
json.dump(tot, open("gpurun_out/pcie_%s_sum.json" % TAG, "w"))
PY
python - $N $TAG <<'PY'
import json, sys
N, TAG = sys.argv[1], sys.argv[2]
for n in ("1", N):
    try:
        d = json.load(open("gpurun_out/bench_n%s_%s.json" % (n, TAG)))
    except Exception as e:
        print("bench n=%s FAILED %s" % (n, e)); continue
    print("n=%s value %.3e ms/step %.3f e2e %.3e sustained %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d.get("sustained"))))
    print("   comm", json.dumps(d.get("comm"))[:2500])
PY
cat gpurun_out/rc_$TAG.txt
