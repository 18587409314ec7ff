"""Inner loops of a kernel in the built library with their instruction mix:
python scripts/sass_loops.py [kernel substring] [min instructions]"""
import collections, re, subprocess, sys
lib = "dvo_slam_b200/libdvo_b200.so"
want = sys.argv[1] if len(sys.argv) > 1 else "k_level_persistent"
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); funcs[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn:
        funcs[fn].append((int(m.group(1), 16), m.group(2)))
for fn, ins in funcs.items():
    if want not in fn:
        continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(fn[:90], "instructions", len(ins))
    loops = []
    for i, (a, t) in enumerate(ins):
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr and i - addr[tgt] + 1 >= minlen:
                    loops.append((addr[tgt], i))
    for lo, hi in loops:
        if any(l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops) and hi - lo > 1500:
            continue   # outer loops
        c = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in ins[lo:hi + 1])
        print(f"  loop {ins[lo][0]:#x}..{ins[hi][0]:#x} n={hi - lo + 1}", " ".join(f"{k}:{v}" for k, v in c.most_common(24)))
