"""Per-source-line instruction / stall-sample table from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel_index] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
k = -1; out = []
for r in rows:
    if r and r[0] == "Function Name":
        k += 1; name = r[1]; continue
    if k == kidx and len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        out.append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, int(r[0]), r[1][:120]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("kernel", kidx, "warp-instructions", tot, "samples", ts)
for o in sorted(out, reverse=True)[:top]:
    print("%10d %5.1f%%  smp %5.1f%%  L%-4d %s" % (o[0], 100.0 * o[0] / tot, 100.0 * o[1] / ts, o[2], o[3]))
