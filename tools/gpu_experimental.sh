#!/bin/bash
# One GPU call for the paths that are written but not yet validated on a B200 (tests/test_gpu_experimental.py,
# DESIGN.md §9).  Every feature is tested on its own (a hang or failure in one new kernel must not hide the others),
# every pytest run has a per-test timeout, and the headline bench is A/B'd only with the features whose tests passed.
#   bash tools/gpu_experimental.sh            (under gpurun: give the call ~40 minutes; each group is capped at 15)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export TASU_EXPERIMENTAL=1
: > gpurun_out/rc_experimental.txt
declare -A OK
for group in bf16_encoder prefetch wide_ctc widegemm streamk pair; do
    timeout 420 python -m pytest tests/test_gpu_experimental.py -q -m gpu -x --timeout 120 --timeout-method=thread -k "$group" \
        > gpurun_out/t_experimental_$group.log 2>&1
    rc=$?
    echo "$group rc=$rc" >> gpurun_out/rc_experimental.txt
    OK[$group]=$rc
    tail -6 gpurun_out/t_experimental_$group.log
    # a wedged GPU would fail everything after it: stop early and say so
    if ! timeout 60 python -c "import torch; torch.zeros(1, device='cuda').sum().item()" > /dev/null 2>&1; then
        echo "GPU unresponsive after group $group" >> gpurun_out/rc_experimental.txt
        cat gpurun_out/rc_experimental.txt
        exit 1
    fi
done
cat gpurun_out/rc_experimental.txt

if [ "${OK[pair]}" = 0 ] && [ "${OK[streamk]}" = 0 ]; then
    timeout 300 python tools/bench_pair_gemm.py > gpurun_out/pair_gemm.md 2>&1
    cat gpurun_out/pair_gemm.md
fi

# optional (SANITIZE=1): compute-sanitizer memcheck + racecheck over one small case of every group that passed
if [ "${SANITIZE:-0}" = 1 ]; then
    declare -A SMALL=( [prefetch]="prefetch_gemm and 130-260-72" [wide_ctc]="wide_ctc_head_stats and 2-130-4-300-64-7"
                       [widegemm]="widegemm_epilogue and 130-260-72" [streamk]="streamk_gemm and 300-512-2048 and not device"
                       [pair]="pair_gemm_matches and 300-260-1100" )
    for group in prefetch wide_ctc widegemm streamk pair; do
        [ "${OK[$group]}" = 0 ] || continue
        for tool in memcheck racecheck; do
            timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_experimental.py -q -m gpu \
                -k "${SMALL[$group]}" > gpurun_out/sanitize_${group}_$tool.log 2>&1
            echo "$group $tool rc=$?" >> gpurun_out/rc_experimental.txt
            grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_${group}_$tool.log | tail -2
        done
    done
fi

# A/B of the headline bench: default, then every validated feature alone, then all validated features together
FLAGS=("")
ALL=""
[ "${OK[bf16_encoder]}" = 0 ] && FLAGS+=("--host-bf16")
[ "${OK[prefetch]}" = 0 ] && FLAGS+=("--epi-prefetch 1" "--epi-prefetch 2" "--epi-prefetch 3") && ALL="$ALL --epi-prefetch 3"
[ "${OK[wide_ctc]}" = 0 ] && FLAGS+=("--stats-wide") && ALL="${ALL/--epi-prefetch 3/--epi-prefetch 1} --stats-wide"
[ "${OK[widegemm]}" = 0 ] && FLAGS+=("--wide-epi") && ALL="${ALL/--epi-prefetch 1/} --wide-epi"
[ "${OK[streamk]}" = 0 ] && FLAGS+=("--streamk") && ALL="$ALL --streamk"
[ "${OK[pair]}" = 0 ] && FLAGS+=("--pair-gemm 1" "--pair-gemm 2" "--pair-gemm 4" "--pair-gemm 7")
[ -n "$ALL" ] && FLAGS+=("$ALL")
for flag in "${FLAGS[@]}"; do
    name="gpurun_out/bench_exp$(echo "$flag" | tr -d ' ' | tr -- '-' '_').json"
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $flag > "$name" 2> /dev/null
    python - "$name" "$flag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("bench %-40s FAILED (%s)" % (sys.argv[2], e)); sys.exit(0)
k = d["kernels"]
print("bench %-40s %.3f ms/step | stats %.3f  softmax-gemm %.3f  gemm1 %.3f | e2e %.3f ms" % (
    sys.argv[2] or "(default)", d["ms_per_step"], k["ctc_head_stats"]["ms"], k["ctc_softmax_gemm"]["ms"],
    k["projector_gemm1"]["ms"], d["e2e"]["ms_per_step"]))
PY
done
