"""Instruction mix per kernel from the built library: python scripts/sass_mix.py [kernel substring] [top]"""
import collections, re, subprocess, sys
lib = "dvo_slam_b200/libdvo_b200.so"
want = sys.argv[1] if len(sys.argv) > 1 else ""
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); mix[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(2)] += 1
for fn, c in mix.items():
    if want in fn:
        short = re.sub(r"_ZN8dvo_b200\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+", "", fn)[:40]
        print(short, "total", sum(c.values()), " ".join(f"{k}:{v}" for k, v in c.most_common(top)))
