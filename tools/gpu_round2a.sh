#!/bin/bash
# round 2, call A: (1) the three experimental groups whose round-1 tests tripped over the pad-column check, run to the end;
# (2) the exact-decision tests; (3) A/B of the headline bench with exact decisions on (the new default).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
free -g | head -2; nproc
export TASU_EXPERIMENTAL=1
: > gpurun_out/rc_r2a.txt
for group in pair; do
    timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 120 --timeout-method=thread -k "$group" \
        > gpurun_out/t2a_$group.log 2>&1
    echo "$group rc=$?" >> gpurun_out/rc_r2a.txt
    tail -4 gpurun_out/t2a_$group.log
    if ! timeout 60 python -c "import torch; torch.zeros(1, device='cuda').sum().item()" > /dev/null 2>&1; then
        echo "GPU unresponsive after group $group" >> gpurun_out/rc_r2a.txt; cat gpurun_out/rc_r2a.txt; exit 1
    fi
done
unset TASU_EXPERIMENTAL
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --timeout 600 -k "exact or refined or fullsize_vs_oracle or midsize or overflow" \
    > gpurun_out/t2a_exact.log 2>&1
echo "exact rc=$?" >> gpurun_out/rc_r2a.txt
tail -30 gpurun_out/t2a_exact.log
cat gpurun_out/rc_r2a.txt
for flag in "" "--no-exact-decisions" "--wide-epi" "--streamk" "--pair-gemm 1" "--pair-gemm 2" "--pair-gemm 4" "--wide-epi --streamk --epi-prefetch 2"; do
    name="gpurun_out/bench2a$(echo "$flag" | tr -d ' ' | tr -- '-' '_').json"
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $flag > "$name" 2> gpurun_out/bench2a.err || tail -5 gpurun_out/bench2a.err
    python - "$name" "$flag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("bench %-40s FAILED (%s)" % (sys.argv[2], e)); sys.exit(0)
k = d["kernels"]
print("bench %-40s %.3f ms/step | stats %.3f refine %.3f softmax-gemm %.3f pool %.3f gemm1 %.3f gemm2 %.3f splice %.3f | e2e %.3f ms | n_out %d" % (
    sys.argv[2] or "(default)", d["ms_per_step"], k["ctc_head_stats"]["ms"], k.get("refine_ambiguous", {}).get("ms", 0), k["ctc_softmax_gemm"]["ms"],
    k["pool_tail"]["ms"], k["projector_gemm1"]["ms"], k["projector_gemm2"]["ms"], k["splice_scatter"]["ms"], d["e2e"]["ms_per_step"],
    d["config"]["compressed_rows_per_step"]))
PY
done
