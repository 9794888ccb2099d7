#!/bin/bash
# One GPU call for the paths that are written but not yet validated on a B200 (tests/test_gpu_experimental.py):
# every test runs under its own timeout so that a hang in a new tcgen05 kernel cannot hold the box, and the pair-mode
# GEMM is timed against the default kernel on the projector shapes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TASU_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu -x --timeout 120 > gpurun_out/t_experimental.log 2>&1
echo "experimental rc=$?" > gpurun_out/rc_experimental.txt
tail -25 gpurun_out/t_experimental.log
if grep -q "rc=0" gpurun_out/rc_experimental.txt; then
    timeout 300 python tools/bench_pair_gemm.py > gpurun_out/pair_gemm.md 2>&1
    cat gpurun_out/pair_gemm.md
fi
cat gpurun_out/rc_experimental.txt
# A/B of the headline bench with the experimental GEMM paths (only after their tests passed)
if grep -q "rc=0" gpurun_out/rc_experimental.txt; then
    for flag in "" "--epi-prefetch 3" "--streamk" "--streamk --epi-prefetch 3" "--pair-gemm 1" "--pair-gemm 2" "--pair-gemm 3" "--pair-gemm 4" "--pair-gemm 7" "--streamk --pair-gemm 6"; do
        timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $flag > "gpurun_out/bench_exp${flag// /_}.json" 2> /dev/null
        python - "gpurun_out/bench_exp${flag// /_}.json" "$flag" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("bench %-22s %.3f ms/step  gemm1 %.3f ms  e2e %.3f ms" % (sys.argv[2] or "(default)", d["ms_per_step"], d["kernels"]["projector_gemm1"]["ms"], d["e2e"]["ms_per_step"]))
PY
    done
fi
