"""Executed warp-instructions of the level kernel by source function, from an `ncu --set full --import-source on` report:

    python scripts/instruction_accounting.py gpurun_out/r02_full.ncu-rep r02

Writes profiles/<name>_instruction_accounting.txt.  Every SASS instruction is attributed (through -lineinfo) to a source line and
the line to the function that lexically contains it (inlined code counts for the function it was written in)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CSRC = os.path.join(ROOT, "dvo_slam_b200", "csrc")


def function_map(path):
    """line number -> name of the enclosing function (crude: a definition starts at a line that opens with a CUDA / C++
    function qualifier and carries `name(`; it extends to the next such line)."""
    starts = []
    lines = open(path).read().split("\n")
    for i, ln in enumerate(lines, 1):
        if re.match(r"^(template\s*<|__device__|__global__|__host__|static |inline |int |void |DVO_HD)", ln) and not ln.startswith("template"):
            m = re.search(r"([A-Za-z_]\w*)\s*\(", ln.split("//")[0])
            if m and m.group(1) not in ("__launch_bounds__", "__align__", "if", "for", "while"):
                starts.append((i, m.group(1)))
        elif ln.startswith("k_level_persistent("):
            starts.append((i, "k_level_persistent"))
    out = {}
    for k, (ln, name) in enumerate(starts):
        end = starts[k + 1][0] if k + 1 < len(starts) else len(lines) + 1
        for j in range(ln, end):
            out[j] = name
    return out


def main():
    rep, name = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:k_level_persistent", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    maps = {}
    fname = ""
    acc = collections.defaultdict(lambda: [0, 0])
    for r in rows:
        if r and r[0] in ("File Name", "File Path"):
            fname = os.path.basename(r[1])
            continue
        if r and r[0].strip().isdigit() and len(r) > 8:
            try:
                n, st, ln = int(r[7]), int(r[4]) if r[4].isdigit() else 0, int(r[0])
            except ValueError:
                continue
            if fname not in maps:
                p = os.path.join(CSRC, fname)
                maps[fname] = function_map(p) if os.path.exists(p) else {}
            fn = maps[fname].get(ln, "(other)")
            a = acc[(fname, fn)]
            a[0] += n
            a[1] += st
    tot = sum(a[0] for a in acc.values())
    samples = sum(a[1] for a in acc.values())
    path = os.path.join(ROOT, "profiles", f"{name}_instruction_accounting.txt")
    with open(path, "w") as f:
        f.write(f"# k_level_persistent, one 512-pair step (batch 512, 640x480x5): executed warp-instructions and stall samples by source function,\n"
                f"# ncu --set full --import-source on, CUDA sources {bench.source_stamp()}; {tot / 1e9:.2f} G warp-instructions (incl. predicated-off issue), {samples} samples.\n"
                f"# Polling loops (mbar_wait*, squad_wait) are instructions issued while waiting, not work.\n")
        f.write("%-14s %-28s %8s %8s\n" % ("file", "function", "inst %", "stall %"))
        for (fn_file, fn), (n, st) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            if n / tot < 0.002 and st / max(samples, 1) < 0.002:
                continue
            f.write("%-14s %-28s %8.2f %8.2f\n" % (fn_file, fn, 100.0 * n / tot, 100.0 * st / max(samples, 1)))
    print(open(path).read())


if __name__ == "__main__":
    main()
