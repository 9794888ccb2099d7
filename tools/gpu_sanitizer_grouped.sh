#!/bin/bash
# compute-sanitizer over the grouped kept-frame layout (group plan, gather, kGrouped GEMM epilogue, LN finish, perm splice)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_${1:-grouped}.log; : > $OUT
run() { echo "### compute-sanitizer --tool $1 python -m pytest $2 -m gpu -k \"$3\"" >> $OUT
        timeout ${4:-300} compute-sanitizer --tool $1 python -m pytest $2 -q -m gpu -x --timeout 280 -k "$3" 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -6 >> $OUT; echo "rc=$?" >> $OUT; }
run memcheck tests/test_gpu_grouped.py "6-120-0 or 9-300-7 or 3-40-2"
run memcheck tests/test_gpu_grouped.py "capacity_overflow"
run racecheck tests/test_gpu_grouped.py "6-120-0 or 3-40-2"
cat $OUT
