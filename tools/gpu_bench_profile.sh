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
    -k regex:'frame_stats_kernel|meanpool_kernel|gemm_bf16_tn_kernel|splice_fused_kernel|ctc_stats_kernel|pool_tail_kernel|gather_kept_rows_kernel' -s 24 -c 8 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/rc_$TAG.txt
cat gpurun_out/rc_$TAG.txt; tail -3 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err

# ---- configs[2] text-only training step: bench line, in-situ kernel table, ncu launch list, full capture of its kernels
timeout 300 python tools/bench_train.py --steps 50 --kernel-table > gpurun_out/train_$TAG.json 2> gpurun_out/train_table_$TAG.md; echo "train rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 300 python tools/bench_train.py --steps 30 --dense > gpurun_out/train_dense_$TAG.json 2> /dev/null
timeout 300 python tools/bench_mixed.py --steps 10 > gpurun_out/mixed_$TAG.json 2> gpurun_out/mixed_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$TAG.csv \
    python tools/bench_train.py --steps 2 --warmup 1 --no-prefetch > gpurun_out/ncu_train_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'tokrow_|gather_rows_kernel|splice_fused_kernel|transpose_cast_kernel|reduce_partials' -s 16 -c 10 \
    -o gpurun_out/prof_train_$TAG -f python tools/bench_train.py --steps 2 --warmup 1 --no-prefetch > gpurun_out/ncu_full_train_$TAG.log 2>&1; echo "ncu_full_train rc=$?" >> gpurun_out/rc_$TAG.txt
cat gpurun_out/train_$TAG.json gpurun_out/train_dense_$TAG.json gpurun_out/mixed_$TAG.json; tail -3 gpurun_out/rc_$TAG.txt
