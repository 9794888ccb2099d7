#!/bin/bash
# quick 1-GPU loop: (optional) test subset, headline bench without the slow sections, optional extra command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-q}
: > gpurun_out/rc_$TAG.txt
if [ -n "${TESTS:-}" ]; then
timeout 1500 python -m pytest $TESTS -q -m gpu -x --timeout 600 > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_$TAG.txt
tail -6 gpurun_out/t_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-comm --no-fp32-leg ${BENCH_FLAGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/rc_$TAG.txt
tail -3 gpurun_out/bench_$TAG.err
python - $TAG <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
k = d["kernels"]
print("%.3f ms/step value %.3e | " % (d["ms_per_step"], d["value"]) + " ".join("%s %.3f" % (n[:14], v["ms"]) for n, v in k.items()) + " | e2e %.3f" % d["e2e"]["ms_per_step"])
print("sustained", d["sustained"]); print("launches", d["gpu_launches"], "clocks", d["clocks"])
PY
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA" > gpurun_out/extra_$TAG.log 2>&1; echo "extra rc=$?" >> gpurun_out/rc_$TAG.txt; tail -40 gpurun_out/extra_$TAG.log; fi
cat gpurun_out/rc_$TAG.txt
