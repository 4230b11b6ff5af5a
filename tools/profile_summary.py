#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep) into the committed summaries under profiles/.
usage: profile_summary.py <report.ncu-rep> <name>    -> profiles/<name>.md, updates profiles/traffic.json"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, name = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = [f"# ncu summary: {name}", "", f"source report: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on; cold-cache, serialised launches)", ""]
traffic = None
for r in data:
    out.append("| metric | value | unit |"); out.append("|---|---|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} | {units[i]} |")
    try:
        rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        f = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        traffic = rd * f[ur] + wr * f[uw]
        out.append(f"| dram bytes read+write per launch | {traffic:.0f} | byte |")
    except Exception:
        pass
    out.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = None
for r in rows:
    if r and r[0] == "Kernel Name": k = {"hdr": None, "rows": []}; continue
    if k is None: continue
    if k["hdr"] is None: k["hdr"] = r; continue
    k["rows"].append(r)
if k and k["hdr"]:
    ix = {n: i for i, n in enumerate(k["hdr"])}
    stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ix["# Samples"]]) for r in k["rows"]) or 1
    agg = sorted(((sum(int(r[ix[s]]) for r in k["rows"]), s) for s in stalls), reverse=True)
    out.append("## warp stall samples (all warps)"); out.append("")
    out.append("| reason | samples | share |"); out.append("|---|---|---|")
    for v, s in agg[:10]: out.append(f"| {s} | {v} | {100.0 * v / tot:.1f} % |")
    out.append("")
    mix = {}
    for r in k["rows"]:
        op = r[ix["Source"]].strip().split()
        if not op: continue
        o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        o = o.split(".")[0]
        mix[o] = mix.get(o, 0) + int(r[ix["Instructions Executed"]])
    tot_i = sum(mix.values()) or 1
    out.append("## executed warp-instructions by opcode (top 16)"); out.append("")
    out.append("| opcode | warp-instructions | share |"); out.append("|---|---|---|")
    for o, v in sorted(mix.items(), key=lambda t: -t[1])[:16]: out.append(f"| {o} | {v} | {100.0 * v / tot_i:.1f} % |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", name + ".md"), "w").write("\n".join(out) + "\n")
if key and traffic:
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d[key] = traffic
    json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
print("wrote profiles/" + name + ".md", "traffic", traffic)
