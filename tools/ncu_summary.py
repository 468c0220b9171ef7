"""Summarise an ncu report (run HERE, no GPU needed):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep TAG BATCH
Writes profiles/<TAG>_<KERNEL>_ncu.txt (KERNEL = optional 4th argument, default mpc_kernel) (key metrics + stall samples per CUDA source line) and updates
profiles/ncu_summary.json (DRAM bytes per hopper of mpc_kernel, read by bench.py for roofline.traffic)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
kname = sys.argv[4] if len(sys.argv) > 4 else "mpc_kernel"      # output file suffix; only "mpc_kernel" updates ncu_summary.json
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
lines = [f"# ncu --set full --clock-control none --import-source on -k regex:{kname}  ({rep}); batch {batch} hoppers per launch"]
vals = {}
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        vals[k] = [r[i] for r in data]
        lines.append(f"{k} [{units[i]}]: " + ", ".join(vals[k]))


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


ur = units[hdr.index("dram__bytes_read.sum")]
uw = units[hdr.index("dram__bytes_write.sum")]
per_launch = [to_bytes(a, ur) + to_bytes(b, uw) for a, b in zip(vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"])]
dram_per_hopper = sum(per_launch) / len(per_launch) / batch

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, agg, h2, stalls, tot_s = None, {}, None, {}, 0
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        h2 = r
        continue
    if h2 is None or len(r) < 8:
        continue
    if r[0] != "" and r[2] == "-":
        try:
            s_, ins = int(r[h2.index("# Samples")]), int(r[h2.index("Instructions Executed")])
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:100]])
        a[0] += s_
        a[1] += ins
        for i, h in enumerate(h2):
            if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit():
                stalls[h] = stalls.get(h, 0) + int(r[i])
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
ts = sum(stalls.values()) or 1
lines.append("")
lines.append("# warp stall reasons (share of samples): " + ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
lines.append("# stall samples per CUDA source line (top 25)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    lines.append(f"{k[0]}:{k[1]:4d} samples {100 * a[0] / tot:5.1f}% inst {100 * a[1] / toti:5.1f}%  {a[2]}")
out = os.path.join(ROOT, "profiles", f"{tag}_{kname}_ncu.txt")
open(out, "w").write("\n".join(lines) + "\n")
js = os.path.join(ROOT, "profiles", "ncu_summary.json") if kname == "mpc_kernel" else os.devnull
json.dump({"source": os.path.basename(out), "capture_batch": batch, "mpc_kernel_dram_bytes_per_hopper": dram_per_hopper,
           "mpc_kernel_dram_bytes_per_launch_at_capture": sum(per_launch) / len(per_launch)}, open(js, "w"), indent=1)
print(open(out).read())
