#!/bin/bash
# One GPU call: smoke, bench (own arm + reference arm), ncu launch list and full captures of the top kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" > gpurun_out/rc_$TAG.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/rc_$TAG.txt
kill $SMI
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench_ref rc=$?" >> gpurun_out/rc_$TAG.txt
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/rc_$TAG.txt
# full capture of the hot kernels (one launch each, after warm-up launches)
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'frame_stats_kernel|meanpool_kernel|gemm_bf16_tn_kernel|splice_scatter_kernel|ctc_stats_kernel|pool_tail_kernel|gather_kept_rows_kernel' -s 21 -c 7 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/rc_$TAG.txt
cat gpurun_out/rc_$TAG.txt; tail -3 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
