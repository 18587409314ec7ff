"""Per-source-line instruction counts from an ncu report: python scripts/ncu_lines.py <rep> <kernel regex> [top]"""
import csv, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, lines = "", []
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        fname = os.path.basename(r[1]); continue
    if r and r[0].strip().isdigit() and len(r) > 8:
        try:
            lines.append((int(r[7]), fname, int(r[0]), r[1][:105], int(r[4]) if r[4].isdigit() else 0))
        except ValueError:
            pass
tot = sum(l[0] for l in lines)
stall = sum(l[4] for l in lines)
print("total warp instructions", tot, "samples", stall)
for n, f, ln, src, st in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%-4d %s" % (100.0 * n / tot, 100.0 * st / max(stall, 1), f[:10], ln, src))
