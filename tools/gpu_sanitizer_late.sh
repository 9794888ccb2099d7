#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (fused splice, fused cross-attention, host pipeline); small cases
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_${1:-r02late}.log; : > $OUT
run() { echo "### compute-sanitizer --tool $1 python -m pytest $2 -m gpu -k \"$3\"" >> $OUT
        timeout ${4:-420} compute-sanitizer --tool $1 python -m pytest $2 -q -m gpu -x --timeout 400 -k "$3" 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -6 >> $OUT; echo "rc=$?" >> $OUT; }
run memcheck tests/test_gpu_kernels.py "merge_golden or merge_random"
run memcheck tests/test_gpu_train.py "splice_backward"
run memcheck tests/test_gpu_gemm.py "attn_softmax_pv and (129-515 or 50-77 or 128-256)"
run memcheck tests/test_gpu_gemm.py "cross_attention_projector and 25055"
run memcheck tests/test_gpu_gemm_variants.py "host_pipeline"
run racecheck tests/test_gpu_gemm.py "attn_softmax_pv and (129-515 or 50-77)"
run racecheck tests/test_gpu_kernels.py "merge_golden"
cat $OUT
