"""Turn gpurun_out/*.ncu-rep + launch list into committed summaries under profiles/.
usage: make_profile_summary.py <tag> <full.ncu-rep> [launches.csv]"""
import csv, io, json, os, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"), ("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "red_sectors"),
        ("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum", "atom_sectors"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts")]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
note = sys.argv[4] if len(sys.argv) > 4 else "cfg3 view: 1M Gaussians, 1920x1080, SH deg 3, fused activations; `ncu --set full --clock-control none --import-source on -k regex:k_ ... python tools/profile_one.py cfg3 3`"
lines = [f"# ncu --set full summary ({tag})\n",
         f"source: `{os.path.basename(rep)}` (gpurun_out/, not committed); {note}\n",
         "red_sectors / atom_sectors: L1->L2 write sectors of global reductions / returning atomics (the `lts__t_sectors_op_*` counters are not exposed by this ncu build)\n",
         "| kernel | " + " | ".join(c[1] for c in cols) + " | top stalls |", "|---|" + "---|" * (len(cols) + 1)]
traffic = {}
issue = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("::")[-1].split("(")[0]
    vals = []
    for k, _ in cols:
        if k in idx:
            v = r[idx[k]].replace(",", "")
            try:
                f = float(v); vals.append(f"{f:.3g}" + (" " + units[idx[k]] if units[idx[k]] not in ("%", "", "inst", "register/thread", "sector") else ""))
            except ValueError:
                vals.append(v)
        else:
            vals.append("-")
    st = sorted(((float(r[idx[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:3]
    lines.append(f"| {name} | " + " | ".join(vals) + " | " + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")
    def mb(k):
        v = float(r[idx[k]].replace(",", "")); u = units[idx[k]]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
    base = name.split("<")[0]  # template instantiations of one kernel count together
    traffic[base] = traffic.get(base, 0.0) + mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    ik = "smsp__issue_active.avg.pct_of_peak_sustained_active"
    if ik in idx and base not in issue:
        issue[base] = float(r[idx[ik]].replace(",", ""))
open(os.path.join(out, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
grp = {"preprocess_fwd": traffic.get("k_preprocess_fwd"), "render_fwd": traffic.get("k_render_fwd"), "render_bwd": traffic.get("k_render_bwd"),
       "preprocess_bwd": traffic.get("k_preprocess_bwd"),
       "binning": sum(v for k, v in traffic.items() if k.startswith(("k_tile_scan", "k_scatter", "k_tile_sort"))) or None}
# merge into the committed file (bench.py reads it) only when the capture holds all five stages of a view;
# a partial capture (e.g. tools/profile_aux.py) must not drop the stages it did not see
tf = os.path.join(out, "ncu_traffic.json")
if all(grp.values()):
    old = json.load(open(tf)) if os.path.exists(tf) else {}
    old.update(grp)
    old["_issue_active_pct"] = {k[2:]: round(v, 1) for k, v in issue.items() if k in ("k_preprocess_fwd", "k_render_fwd", "k_render_bwd", "k_preprocess_bwd", "k_tile_sort", "k_scatter")}
    old["_note"] = f"bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, cfg3 view), from {os.path.basename(rep)}"
    json.dump(old, open(tf, "w"), indent=1)
if launches:
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
    tot = {}; n = {}
    for r in rows:
        k = r[4].split("(")[0].split("::")[-1][:60]; t = float(r[-1].replace(",", ""))
        u = r[-2]
        t *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        tot[k] = tot.get(k, 0) + t; n[k] = n.get(k, 0) + 1
    s = sum(tot.values())
    L = [f"# launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches of steady-state steps\n",
         "cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        L.append(f"| {k} | {n[k]} | {tot[k]:.1f} | {100*tot[k]/s:.1f}% |")
    open(os.path.join(out, f"{tag}_launches.md"), "w").write("\n".join(L) + "\n")
print(open(os.path.join(out, f"{tag}_ncu_summary.md")).read()[:3000])
