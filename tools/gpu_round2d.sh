#!/bin/bash
# 1 GPU: whole GPU test suite, smoke, headline bench, reference arm at the real batch size (timed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02d}
: > gpurun_out/rc_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_$TAG.txt
tail -25 gpurun_out/t_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/rc_$TAG.txt
tail -5 gpurun_out/bench_$TAG.err
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err ) 2> gpurun_out/ref_time_$TAG.txt; echo "ref rc=$?" >> gpurun_out/rc_$TAG.txt
cat gpurun_out/ref_time_$TAG.txt; cat gpurun_out/bench_ref_$TAG.json | cut -c1-900
python - $TAG <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
k = d["kernels"]
print("%.3f ms/step value %.3e | " % (d["ms_per_step"], d["value"]) + " ".join("%s %.3f" % (n[:14], v["ms"]) for n, v in k.items()) + " | e2e %.3f" % d["e2e"]["ms_per_step"])
print("roofline", d["roofline"]); print("sustained", d["sustained"]); print("cpu", d.get("cpu_baseline")); print("launches", d["gpu_launches"], "amb", d["config"]["frames_refined_in_fp32_last_step"])
PY
cat gpurun_out/rc_$TAG.txt
