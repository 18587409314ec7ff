"""Aggregate ncu per-line instruction counts into named line ranges of a source file.
usage: python scripts/ncu_ranges.py <rep> <kernel regex> <file substring> name:lo-hi [name:lo-hi ...]"""
import csv, os, subprocess, sys
rep, kern, fsub = sys.argv[1], sys.argv[2], sys.argv[3]
ranges = []
for a in sys.argv[4:]:
    name, r = a.split(":"); lo, hi = r.split("-"); ranges.append((name, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
fname, tot, agg, stall, stot = "", 0, {}, {}, 0
for r in csv.reader(out.splitlines()):
    if r and r[0] in ("File Name", "File Path"):
        fname = os.path.basename(r[1]); continue
    if r and r[0].strip().isdigit() and len(r) > 8:
        try:
            n, ln = int(r[7]), int(r[0]); st = int(r[4]) if r[4].isdigit() else 0
        except ValueError:
            continue
        tot += n; stot += st
        key = "other:" + fname
        if fsub in fname:
            for name, lo, hi in ranges:
                if lo <= ln <= hi:
                    key = name; break
        agg[key] = agg.get(key, 0) + n; stall[key] = stall.get(key, 0) + st
print("total", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%-28s %6.1f%% inst %6.1f%% stall-samples" % (k, 100.0 * v / tot, 100.0 * stall[k] / max(stot, 1)))
