#!/usr/bin/env python
"""Summarise gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into profiles/ (tracked).

    python tools/summarize_ncu.py r01b
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# ---- launch list: per-kernel count / total / share (cold-cache, serialised: compare SHARES)
rows = []
path = os.path.join(go, f"launches_{tag}.csv")
if os.path.isfile(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list — {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n"
                "(cold-cache, serialised launches: compare shares, not absolutes)\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |\n")

# ---- full capture: key raw metrics per captured kernel
rep = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hdr, units = rd[0], rd[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out_dir, f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full — {tag}\n\n`ncu --set full --clock-control none --import-source on -k regex:… python bench.py --steps 2 --warmup 1`\n\n")
        for r in rd[2:]:
            f.write("## " + r[hdr.index("Kernel Name")][:110] + "\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w, i in idx[1:]:
                f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            f.write("\n")

    # dram traffic per launch of every captured kernel → profiles/traffic.json (read by bench.py's roofline.traffic)
    stage_of = [("ctc_stats_kernel", "ctc_head_stats"), ("gemm_bf16_tn_kernel<1, 6", "ctc_softmax_gemm"),
                ("gemm_bf16_tn_kernel<1, 4", "projector_gemm1"), ("gemm_streamk_kernel<1, 4", "projector_gemm1"),
                ("gemm_bf16_tn_kernel<1, 1", "projector_gemm2"),
                ("gemm_bf16_tn_kernel<0, 1", "ctc_lo_gemm"), ("pool_tail_kernel", "pool_tail"),
                ("splice_fused_kernel", "splice_scatter"), ("frame_stats_kernel", "frame_stats"),
                ("meanpool_kernel", "softmax_meanpool"), ("gather_kept_rows_kernel", "gather_kept_rows"),
                ("gather_grouped_kernel", "gather_kept_rows")]
    tpath = os.path.join(out_dir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    for r in rd[2:]:
        for pat, stage in stage_of:
            if pat in r[ik]:
                traffic[stage] = {"bytes_per_launch": float(r[ir]) * scale.get(units[ir], 1) + float(r[iw]) * scale.get(units[iw], 1),
                                  "source": f"ncu --set full, {tag} (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1)

b = os.path.join(go, f"bench_{tag}.json")
if os.path.isfile(b):
    for src, dst in ((b, f"{tag}_bench.json"), (os.path.join(go, f"bench_ref_{tag}.json"), f"{tag}_bench_reference.json")):
        if os.path.isfile(src):
            with open(src) as f:
                txt = f.read().strip()
            if txt:
                with open(os.path.join(out_dir, dst), "w") as f:
                    json.dump(json.loads(txt.splitlines()[-1]), f, indent=1)
c = os.path.join(go, f"clocks_{tag}.csv")
if os.path.isfile(c):
    with open(c) as f:
        lines = f.read().strip().splitlines()
    sm = sorted(int(l.split(",")[1].strip().split()[0]) for l in lines[1:] if l.count(",") >= 8)
    active = sorted({l.split(",")[4].strip() for l in lines[1:] if l.count(",") >= 8})
    with open(os.path.join(out_dir, f"{tag}_clocks.txt"), "w") as f:
        f.write(f"samples={len(sm)} sm_mhz median={sm[len(sm) // 2] if sm else None} min={sm[0] if sm else None} max={sm[-1] if sm else None} "
                f"reasons_active={active}\n")

# ---- configs[2] training step / configs[3] mixed batch evidence (tools/gpu_bench_profile.sh)
def _copy_json(src_name, dst_name):
    src = os.path.join(go, src_name)
    if os.path.isfile(src):
        txt = open(src).read().strip()
        if txt:
            with open(os.path.join(out_dir, dst_name), "w") as f:
                json.dump(json.loads(txt.splitlines()[-1]), f, indent=1)


for a, b_ in ((f"train_{tag}.json", f"{tag}_train_bench.json"), (f"train_dense_{tag}.json", f"{tag}_train_dense_bench.json"),
              (f"mixed_{tag}.json", f"{tag}_mixed_bench.json")):
    _copy_json(a, b_)
for n in (2, 4, 8):
    for a, b_ in ((f"bench_n{n}_{tag}.json", f"{tag}_bench_n{n}.json"), (f"train_strong_n{n}_{tag}.json", f"{tag}_train_strong_n{n}.json"),
                  (f"train_weak_n{n}_{tag}.json", f"{tag}_train_weak_n{n}.json"), (f"mixed_n{n}_{tag}.json", f"{tag}_mixed_n{n}.json")):
        _copy_json(a, b_)
tt = os.path.join(go, f"train_table_{tag}.md")
if os.path.isfile(tt):
    lines = [l for l in open(tt).read().splitlines() if l.startswith("|") or l.startswith("device-busy")]
    with open(os.path.join(out_dir, f"{tag}_train_kernels.md"), "w") as f:
        f.write(f"# text-only training step, in-situ kernel times — {tag}\n\n`python tools/bench_train.py --steps 50 --kernel-table` "
                "(CUPTI via torch.profiler over 5 steps, kernels running back to back as in the timed loop)\n\n")
        f.write("\n".join(lines) + "\n")
tl = os.path.join(go, f"launches_train_{tag}.csv")
if os.path.isfile(tl):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), tl, "3"], capture_output=True, text=True).stdout
    with open(os.path.join(out_dir, f"{tag}_train_launches.md"), "w") as f:
        f.write(f"# ncu launch list, text-only training step — {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none "
                "python tools/bench_train.py --steps 2 --warmup 1 --no-prefetch` (cold-cache, serialised: compare shares)\n\n" + out)
rep2 = os.path.join(go, f"prof_train_{tag}.ncu-rep")
if os.path.isfile(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hdr, units = rd[0], rd[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out_dir, f"{tag}_train_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full, text-only training step kernels — {tag}\n\n")
        for r in rd[2:]:
            f.write("## " + r[hdr.index("Kernel Name")][:110] + "\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w, i in idx:
                f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            f.write("\n")
print("wrote", sorted(n for n in os.listdir(out_dir) if n.startswith(tag)))
