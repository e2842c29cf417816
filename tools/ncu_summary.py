"""Text summary of an .ncu-rep for profiles/: per captured kernel the metrics the bench's roofline rests on (duration, DRAM bytes, pipe and
memory utilisation, occupancy limits, top stall reasons), and - for the kernels named with --lines - the source lines carrying the most stall
samples (needs -lineinfo + --import-source on). usage: python tools/ncu_summary.py report.ncu-rep [--lines k_fast,k_ba_solve] [--top 18]"""
import csv, subprocess, sys

rep = sys.argv[1]
lines = []; top = 18
for i, a in enumerate(sys.argv):
    if a == "--lines": lines = sys.argv[i + 1].split(",")
    if a == "--top": top = int(sys.argv[i + 1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
W = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
     ("launch__shared_mem_per_block_static", "smem static"), ("launch__shared_mem_per_block_dynamic", "smem dynamic"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__occupancy_limit_registers", "occ. limit regs (blocks)"),
     ("launch__occupancy_limit_shared_mem", "occ. limit smem (blocks)"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"), ("smsp__inst_executed.sum", "warp instructions"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("# %s" % rep.split("/")[-1])
names = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    names.append(name)
    print("\n## %s   (launch id %s)" % (name, r[hdr.index("ID")]))
    for m, lab in W:
        if m in hdr:
            i = hdr.index(m); print("  %-28s %s %s" % (lab, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)[:5]
    print("  top stalls (warps per issue)  " + ", ".join("%s %.2f" % (n, v) for v, n in st))
if lines:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rs = list(csv.reader(txt.splitlines()))
    cur = None; agg = {}; done = set()
    for r in rs:
        if r and r[0] == "Function Name":
            cur = r[1].split("(")[0].split("::")[-1]
            if cur in done: cur = None      # first launch of each kernel only
            else: done.add(cur); agg[cur] = []
            continue
        if r and r[0] == "Line No": H = r; continue
        if cur and len(r) > 8 and r[0].isdigit() and r[2] == "-":
            try:
                iS = H.index("Warp Stall Sampling (All Samples)"); iI = H.index("Instructions Executed")
                agg[cur].append((int(r[iS] or 0), int(r[iI] or 0), int(r[0]), r[1][:130]))
            except (ValueError, IndexError):
                pass
    for k in lines:
        if k not in agg: continue
        tot = sum(a[0] for a in agg[k]) or 1; ti = sum(a[1] for a in agg[k]) or 1
        print("\n## %s: source lines by stall samples (%d samples, %d warp instructions)" % (k, tot, ti))
        for a in sorted(agg[k], reverse=True)[:top]:
            print("  %5.1f%% smp  %5.1f%% inst  L%-5d %s" % (100.0 * a[0] / tot, 100.0 * a[1] / ti, a[2], a[3]))
