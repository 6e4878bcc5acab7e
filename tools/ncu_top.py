"""Summarise an .ncu-rep: per kernel headline metrics + top stall source lines.  Usage: ncu_top.py rep [kernel-regex] [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; nl = int(sys.argv[3]) if len(sys.argv) > 3 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
names = []
for r in rows[2:]:
    n = r[idx["Kernel Name"]]; names.append(n)
    print("==", n.split("::")[-1][:40], " ".join(f"{k.split('.')[0].replace('smsp__','').replace('sm__','')[:28]}={r[idx[k]]}" for k in keys if k in idx))
    vals = sorted(((float(r[idx[h]].replace(',', '')), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:5]
    print("   stalls:", [(f"{v:.2f}", n) for v, n in vals])
if kre:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r)
    hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:nl]:
        print(f"{100*int(r[idx['# Samples']])/tot:5.1f}% exec {r[idx['Instructions Executed']]:>9s} {r[idx['Source']].strip()[:90]}")
