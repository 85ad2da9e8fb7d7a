"""Turn gpurun_out/*.ncu-rep (ncu --set full) and launches_*.csv (gpu__time_duration) into the tracked summaries under
profiles/.   python tools/summarize_ncu.py <round-tag> [rep] [launches.csv]"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
launches = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp insts"),
]

raw_csv = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_raw.csv")       # written on the GPU box when the report is too big to travel
if os.path.exists(rep) or os.path.exists(raw_csv):
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        raw = open(raw_csv).read()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none — {os.path.basename(rep)} (tools/run_kernels.py, B200)\n\n")
        f.write("Per-launch values (cold-cache, serialised replay: compare shares, not absolutes).\n\n")
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            f.write(f"## `{name[:150]}`\n\n| metric | value |\n|---|---|\n")
            for key, label in WANT:
                if key in col:
                    f.write(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |\n")
            f.write("\n")
    print("wrote", f"{tag}_ncu_full_summary.md")
    # dram traffic per launch of the GEMMs, keyed by shape: tools/run_kernels.py launches them in this order
    import json
    order = ["32760x1536x1536", "32760x8960x1536", "32760x1536x8960", "32760x1536x1536_gate", "32760x8960x1536_w4a8",
             "32760x4608x1536"]
    gemm_rows = [r for r in rows[2:] if "gemm_i8_kernel" in r[col["Kernel Name"]]]
    def to_bytes(r, key):
        v = float(r[col[key]].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[key]], 1)
    traffic = {"source": f"profiles/{tag}_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
               "gemm": {}, "gemm_duration_us": {}}
    for name, r in zip(order, gemm_rows[:len(order)]):
        traffic["gemm"][name] = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    attn_rows = [r for r in rows[2:] if "attn_i8_kernel" in r[col["Kernel Name"]]]
    if attn_rows:
        traffic["attn_i8_H12_L32760"] = to_bytes(attn_rows[0], "dram__bytes_read.sum") + to_bytes(attn_rows[0], "dram__bytes_write.sum")
    for key, pat in (("attn_bf16_H12_L32760", "attn_bf16_kp_kernel"), ("ln_mod_quant_32760x1536", "ln_mod_quant_kernel"),
                     ("rmsnorm_rope_32760x1536", "rmsnorm_rope_kernel"), ("had_quant_rows_32760x1536", "had_")):
        rr = [r for r in rows[2:] if pat in r[col["Kernel Name"]]]
        if rr:
            traffic[key] = to_bytes(rr[0], "dram__bytes_read.sum") + to_bytes(rr[0], "dram__bytes_write.sum")
    json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
    print("wrote traffic.json", traffic["gemm"])

if os.path.exists(launches):
    lines = [l for l in open(launches) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row.get("Metric Unit", "ns")
        t *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
        a = agg[row["Kernel Name"]]
        a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_by_kernel.md"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none — {os.path.basename(launches)}\n\n")
        f.write("Command: `python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-variants` (2 of 30 "
                "blocks so the listing stays short; per-block shares are those of the full step).\n\n")
        f.write(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time.\n\n")
        f.write("| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k[:140]}` |\n")
    print("wrote", f"{tag}_launches_by_kernel.md")
