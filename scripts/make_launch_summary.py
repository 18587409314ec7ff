"""profiles/<name>_launches.txt from the ncu launch list (gpurun_out/<name>_launches.csv): per kernel count, total and share.

    python scripts/make_launch_summary.py r02
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "r02"
path = os.path.join(ROOT, "gpurun_out", f"{name}_launches.csv")
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
seq = []
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].split("::")[-1]
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
    seq.append((k, v))
tot = sum(v[1] for v in agg.values())
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ : every launch of this library's kernels during",
       f"# `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (setup, warm-up, 2 resident steps, the e2e legs, single-pair latency loop);",
       f"# CUDA sources {bench.source_stamp()}.  Durations under ncu are serialised and cold-cache: shares, not speeds.",
       "%-28s %7s %14s %12s %7s" % ("kernel", "n", "total_us", "avg_us", "share")]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-28s %7d %14.1f %12.2f %7.3f" % (k[:28], v[0], v[1], v[1] / v[0], v[1] / tot))
# the batch-512 launches of the level kernel (the two timed resident steps are among them)
big = [v for k, v in seq if k.startswith("k_level_persistent") and v > 5000]
if big:
    out.append("# k_level_persistent launches longer than 5 ms (512-pair batches): n = %d, mean %.2f ms, min %.2f, max %.2f" %
               (len(big), sum(big) / len(big) / 1e3, min(big) / 1e3, max(big) / 1e3))
open(os.path.join(ROOT, "profiles", f"{name}_launches.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
