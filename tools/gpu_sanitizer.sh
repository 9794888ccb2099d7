#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (small cases; every run under its own timeout)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_${1:-r02}.log; : > $OUT
run() { echo "### compute-sanitizer --tool $1 python -m pytest $2 -m gpu -k \"$3\"" >> $OUT
        timeout ${4:-420} compute-sanitizer --tool $1 python -m pytest $2 -q -m gpu -x --timeout 400 -k "$3" 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -6 >> $OUT; echo "rc=$?" >> $OUT; }
run memcheck tests/test_gpu_kernels.py "exact_decisions or merge_golden"
run memcheck tests/test_gpu_gemm_variants.py "(streamk_gemm and 300-512-2048) or (pair_gemm_matches and 2-) or device_side_row_count"
run memcheck tests/test_gpu_fullsize.py "fp32x3 or edge_cases or cached_weight"
run memcheck tests/test_gpu_train.py "splice_backward"
run racecheck tests/test_gpu_kernels.py "exact_decisions_on_adversarial or merge_golden"
run racecheck tests/test_gpu_fullsize.py "cached_weight or edge_cases"
cat $OUT
