"""profiles/<name>_traffic.json + profiles/<name>_level_kernels.txt from an `ncu --set full` report of the level kernel.

    python scripts/make_traffic_json.py gpurun_out/r02_full.ncu-rep r02

The report (scripts/gpu_profile.sh) holds the k_level_persistent launches of two consecutive steps of the bench
workload (batch 512, 640x480, 5 levels): the first step is the warm-up, the second half of the launches is one step.
The JSON carries the sha256 stamp of the CUDA sources (bench.source_stamp) so that bench.py only quotes the traffic
of the build it is running.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep, name = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    n = len(body)
    assert n == 1 or (n % 2 == 0 and n > 0), f"{n} launches in the report: expected one step, or a warm-up step and a step"
    step = body if n == 1 else body[n // 2:]
    col = {k: hdr.index(k) for k in KEEP if k in hdr}

    def to_bytes(r, k):
        v, u = fnum(r[col[k]]), units[col[k]]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    def to_ms(r):
        v, u = fnum(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]]
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]

    per = [{"grid": int(fnum(r[col["launch__grid_size"]])), "ms_under_ncu": to_ms(r),
            "dram_bytes": to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")} for r in step]
    total = sum(p["dram_bytes"] for p in per)
    stamp = bench.source_stamp()
    tj = {"source_stamp": stamp, "dram_bytes_per_step": total, "launches_per_step": len(step), "per_launch": per,
          "note": f"dram__bytes_read.sum + dram__bytes_write.sum of the {len(step)} k_level_persistent launches of one step of the bench "
                  f"workload (batch 512, 640x480x5), ncu --set full --clock-control none, profiles/{name}_level_kernels.txt; "
                  f"traffic = this / {len(step)} launches; captured from CUDA sources {stamp}"}
    with open(os.path.join(ROOT, "profiles", f"{name}_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    with open(os.path.join(ROOT, "profiles", f"{name}_level_kernels.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, k_level_persistent launches of one step (second of two), sources {stamp}\n")
        f.write("# (durations under ncu are serialised cold-cache replays: use bench.py's CUDA-event times for speed)\n")
        for i, r in enumerate(step):
            f.write(f"== launch {i}: {r[hdr.index('Kernel Name')][:70]}\n")
            for k in KEEP:
                if k in col:
                    f.write("   %-84s %16s %s\n" % (k, r[col[k]], units[col[k]]))
        f.write("== step totals: dram bytes %.3f GB over %d launches\n" % (total / 1e9, len(step)))
    print(json.dumps(tj)[:400])


if __name__ == "__main__":
    main()
