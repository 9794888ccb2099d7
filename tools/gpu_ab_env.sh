#!/bin/bash
# A/B of one environment switch on the headline bench, interleaved A B A B (same box, same call):
#   VAR=NAME [VALS="0 1"] bash tools/gpu_ab_env.sh tag   → gpurun_out/ab_<tag>_<i>.json + a summary line per run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-ab}
i=0; for rep in 1 2; do for v in ${VALS:-0 1}; do i=$((i+1))
  env $VAR=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-comm --no-fp32-leg --sustained-seconds 0 ${BENCH_FLAGS:-} \
      > gpurun_out/ab_${TAG}_${i}.json 2> gpurun_out/ab_${TAG}_${i}.err || tail -3 gpurun_out/ab_${TAG}_${i}.err
  python - gpurun_out/ab_${TAG}_${i}.json "$VAR=$v" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); k = d["kernels"]
print(sys.argv[2], "%.4f ms/step (1-stream %.4f) e2e %.3f | " % (d["ms_per_step"], d["single_stream"]["ms_per_step"], d["e2e"]["ms_per_step"]) +
      " ".join("%s %.4f" % (n[:12], v["ms"]) for n, v in k.items()))
PY
done; done
