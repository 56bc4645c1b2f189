"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the committed summaries under profiles/.
Usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list: every launch + per-kernel share of the step
lines = [l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg, tot = collections.OrderedDict(), 0.0
with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
    f.write("id,kernel,grid,block,duration_us\n")
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        name = r["Kernel Name"].replace("(int)", "").split("(")[0].replace("void ", "").replace("vitk::", "")
        f.write(f'{r["ID"]},{name},"{r["Grid Size"]}","{r["Block Size"]}",{v:.2f}\n')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
with open(os.path.join(P, f"{tag}_step_breakdown.md"), "w") as f:
    f.write(f"# {tag}: one training step (ViT-B/16@384, batch 16, fwd+bwd+AdamW) — ncu launch list\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/profile_step.py`\n"
            "(per-launch times are cold-cache and serialised: compare shares, not absolutes)\n\n"
            f"total {tot / 1e3:.3f} ms over {len(rows)} launches\n\n| kernel | launches | total µs | share | avg µs |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")

# ---- full captures: key metrics per profiled launch
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(os.path.join(P, f"{tag}_ncu_full_summary.csv"), "w") as out:
    w = csv.writer(out)
    w.writerow(["report", "id", "kernel"] + KEYS)
    for rep in sorted(x for x in os.listdir(G) if x.startswith("prof_") and x.endswith(".ncu-rep")):
        raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        if len(r) < 3:
            continue
        hdr, units = r[0], r[1]
        for row in r[2:]:
            name = row[hdr.index("Kernel Name")].replace("(int)", "").split("(")[0].replace("void ", "").replace("vitk::", "")
            vals = []
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals.append(f"{row[i]} {units[i]}".strip())
                else:
                    vals.append("")
            w.writerow([rep, row[hdr.index("ID")], name] + vals)
# ---- DRAM traffic of the dominant kernel (tcgen05 GEMM launches), per launch and per shape, for bench.py's roofline.traffic.
# Shape = (grid, template instantiation): the M = 16 CLS-row launches of the top layer (small grids) are listed but excluded
# from the headline average, which covers the dense launches of one middle layer (forward + backward).
import json
rows = list(csv.DictReader(open(os.path.join(P, f"{tag}_ncu_full_summary.csv"))))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per, tot_b, n = [], 0.0, 0
for r in rows:
    if "gemm" not in r["kernel"]:
        continue
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = r[k].split()
        b += float(v) * UNIT[u]
    grid = r["launch__grid_size"].split()[0]
    dense = float(grid) >= 100
    per.append({"report": r["report"], "id": r["id"], "kernel": r["kernel"], "grid": grid, "dram_bytes": b, "duration": r["gpu__time_duration.sum"],
                "tensor_pipe_pct": r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"].split()[0], "dense": dense})
    if dense:
        tot_b += b
        n += 1
if n:
    json.dump({"kernel": "gemm2_bf16_kernel", "launches_profiled": n, "dram_bytes_per_launch": tot_b / n, "per_shape": per,
               "source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum; the 12 "
                         "dense GEMM launches of one middle encoder layer, forward + backward, inside a training step)"},
              open(os.path.join(P, f"{tag}_roofline_traffic.json"), "w"), indent=1)
print("wrote", [x for x in os.listdir(P) if x.startswith(tag)])
