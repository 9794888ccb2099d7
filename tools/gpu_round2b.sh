#!/bin/bash
# round 2, call B (1 GPU): whole GPU test suite, smoke, headline bench (+ A/B of the GEMM-1 variants)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/rc_r2b.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2b.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc_r2b.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/t_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_r2b.txt
tail -15 gpurun_out/t_r2b.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?" >> gpurun_out/rc_r2b.txt
tail -5 gpurun_out/bench_r2b.err
for flag in "--no-pair-gemm" "--streamk" "--host-bf16"; do
    name="gpurun_out/bench_r2b$(echo "$flag" | tr -d ' ' | tr -- '-' '_').json"
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-comm --sustained-seconds 0 $flag > "$name" 2> /dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2b*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "FAILED", e); continue
    k = d["kernels"]
    print("%-44s %.3f ms/step | " % (f[11:], d["ms_per_step"]) + " ".join("%s %.3f" % (n[:14], v["ms"]) for n, v in k.items()) + " | e2e %.3f" % d["e2e"]["ms_per_step"])
    if "sustained" in d and d["sustained"]:
        print("   sustained", d["sustained"], "whole", d["whole_step"])
    if "comm" in d:
        print("   comm", json.dumps(d["comm"])[:1500])
PY
cat gpurun_out/rc_r2b.txt
