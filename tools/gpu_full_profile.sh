#!/bin/bash
# 1 GPU: whole GPU test suite, smoke, headline bench with clocks, ncu launch list of the bench
# command, ncu --set full capture of the hot kernels.  Outputs under gpurun_out/*_$TAG.*
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02e}
: > gpurun_out/rc_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc_$TAG.txt
if [ "${SKIP_SUITE:-0}" != 1 ]; then
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_$TAG.txt
tail -25 gpurun_out/t_$TAG.log
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/rc_$TAG.txt
kill $SMI
tail -5 gpurun_out/bench_$TAG.err
if [ "${SKIP_REF:-0}" != 1 ]; then
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err ) 2> gpurun_out/ref_time_$TAG.txt; echo "ref rc=$?" >> gpurun_out/rc_$TAG.txt
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-comm --no-fp32-leg --sustained-seconds 0 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_bf16_tn_kernel|gemm_streamk_kernel|splice_fused_kernel|ctc_stats_kernel|pool_tail_kernel|gather_kept_rows_kernel|gather_grouped_kernel|group_plan_kernel|group_ln_finish_kernel|collapse|cast_rows' -s 36 -c 14 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-comm --no-fp32-leg --sustained-seconds 0 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/rc_$TAG.txt
fi
if [ "${SKIP_OPS:-0}" != 1 ]; then
timeout 300 python tools/bench_attn.py > gpurun_out/attn_$TAG.txt 2>&1; echo "attn rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 600 python tools/bench_ops.py > gpurun_out/ops_$TAG.md 2> gpurun_out/ops_$TAG.err; echo "ops rc=$?" >> gpurun_out/rc_$TAG.txt
timeout 120 python tools/micro/pcie_duplex.py > gpurun_out/pcie_$TAG.json 2>/dev/null; echo "pcie rc=$?" >> gpurun_out/rc_$TAG.txt
fi
python - $TAG <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
k = d["kernels"]
print("%.3f ms/step value %.3e | " % (d["ms_per_step"], d["value"]) + " ".join("%s %.3f" % (n[:14], v["ms"]) for n, v in k.items()) + " | e2e %.3f" % d["e2e"]["ms_per_step"])
print("roofline", d["roofline"]); print("sustained", d["sustained"]); print("fp32", d.get("fp32_leg")); print("launches", d["gpu_launches"], "amb", d["config"]["frames_refined_in_fp32_last_step"])
PY
cat gpurun_out/rc_$TAG.txt
