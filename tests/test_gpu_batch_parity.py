"""Parity at the BENCHMARKED configuration (BASELINE.json configs[2]): the first 512 seeded 640x480 pairs of bench.py,
one dvo_b200_match_batch call (the same plan_level outcome as the bench), against the oracle's FAITHFUL mode -- the mode
that tests/test_reference_pin.py pins bit for bit to the reference's own object code -- run on all host threads.

Asserted for 100 % of the pairs: pose within the stated SE(3) tolerance, identical selected-pixel counts.

Control flow (per-level TerminationCriterion and iteration counts, dense_tracking.cpp:276-284, 312-322, 357-363) is decided
by accept tests `Error < LastError` that near convergence compare numbers equal to ~1e-7: it is sensitive to ANY change of
rounding.  The oracle's own two arithmetic variants (FAITHFUL = the reference's SSE numerics, MIRROR = IEEE operations in
another order) agree on every termination for only ~70 % of pairs and on every iteration count (+-1) for ~50 %
(scripts/oracle_controlflow.py).  The meaningful statement is therefore relative: the CUDA path is no farther from the
reference's numerics than that independent IEEE restatement is.  The test measures both distances on the same 512 pairs and
asserts GPU-vs-FAITHFUL agreement >= MIRROR-vs-FAITHFUL agreement - 5 percentage points, plus absolute floors; the
distributions of |dt|, |dr|, Information and LogLikelihood error are printed and written to gpurun_out/."""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import POSE_TOL_R, POSE_TOL_T, pose_delta

pytestmark = pytest.mark.gpu

B = 512
LEVELS, FIRST, LAST = 5, 4, 0


def _pct(v, q):
    return float(np.percentile(np.asarray(v, dtype=np.float64), q))


def test_batch_512_against_reference_numerics(engine, oracle):
    import torch
    from dvo_slam_b200 import synth
    from dvo_slam_b200.engine import Config

    dev = torch.device("cuda", 0)
    scfg = synth.SceneConfig()
    K = synth.FR1_INTRINSICS
    H, W = scfg.height, scfg.width
    Ir = np.empty((B, H, W), np.float32); Zr = np.empty((B, H, W), np.float32)
    Ic = np.empty((B, H, W), np.float32); Zc = np.empty((B, H, W), np.float32)
    for i in range(B):                      # the bench's seeds: rank 0 uses seeds 0 .. B-1
        p = synth.make_pair(i, scfg, device=dev)
        Ir[i] = p["I_ref"].cpu().numpy(); Zr[i] = p["Z_ref"].cpu().numpy()
        Ic[i] = p["I_cur"].cpu().numpy(); Zc[i] = p["Z_cur"].cpu().numpy()
    cfg = Config(first_level=FIRST, last_level=LAST, max_iterations_per_level=50, precision=1e-4)
    ocfg = oracle.config(first_level=FIRST, last_level=LAST, max_iterations_per_level=50, precision=1e-4)
    res = engine.match_batch(engine.pyramid_batch(Ir, Zr, K, LEVELS), engine.pyramid_batch(Ic, Zc, K, LEVELS), cfg)

    def cpu(i):
        oref, ocur = oracle.Pyramid(Ir[i], Zr[i], K, LEVELS), oracle.Pyramid(Ic[i], Zc[i], K, LEVELS)
        return (oracle.match(oref, ocur, ocfg, oracle.mode("faithful")), oracle.match(oref, ocur, ocfg, oracle.mode("mirror")))

    with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
        cpu_res = list(ex.map(cpu, range(B)))

    def flow(levels):
        return ([l["termination"] for l in levels], [l["num_iterations"] for l in levels])

    dts, drs, info_err, ll_err = [], [], [], []
    agree = {"gpu": {"term": 0, "it1": 0, "both": 0}, "mirror": {"term": 0, "it1": 0, "both": 0}}
    for i in range(B):
        fa, mi = cpu_res[i]
        r = res[i]
        dt, dr = pose_delta(fa["T"], r.transformation)
        assert dt < POSE_TOL_T and dr < POSE_TOL_R, (i, dt, dr)
        assert [l["valid_pixels"] for l in r.levels] == [l["valid_pixels"] for l in fa["levels"]]
        assert not r.is_nan()
        dts.append(dt); drs.append(dr)
        info_err.append(np.linalg.norm(r.information - fa["information"]) / np.linalg.norm(fa["information"]))
        ll_err.append(abs(r.log_likelihood - fa["log_likelihood"]) / abs(fa["log_likelihood"]))
        ft, fi = flow(fa["levels"])
        for name, (t, it) in (("gpu", flow(r.levels)), ("mirror", flow(mi["levels"]))):
            same_t = t == ft
            within1 = all(abs(a - b) <= 1 for a, b in zip(it, fi))
            agree[name]["term"] += same_t; agree[name]["it1"] += within1; agree[name]["both"] += same_t and within1
    rate = {k: {m: v / B for m, v in d.items()} for k, d in agree.items()}
    summary = {
        "pairs": B,
        "pose_dt_m": {"median": _pct(dts, 50), "p95": _pct(dts, 95), "p99": _pct(dts, 99), "max": max(dts)},
        "pose_dr_rad": {"median": _pct(drs, 50), "p95": _pct(drs, 95), "p99": _pct(drs, 99), "max": max(drs)},
        "information_rel_frobenius": {"median": _pct(info_err, 50), "p95": _pct(info_err, 95), "max": max(info_err)},
        "log_likelihood_rel": {"median": _pct(ll_err, 50), "p95": _pct(ll_err, 95), "max": max(ll_err)},
        "control_flow_vs_faithful": rate,
    }
    print("\nbatch-512 parity vs FAITHFUL:", json.dumps(summary))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "batch512_parity.json"), "w") as f:
        json.dump(summary, f, indent=1)
    # the CUDA path is as close to the reference's numerics as an independent IEEE restatement of the same algorithm
    for m in ("term", "it1", "both"):
        assert rate["gpu"][m] >= rate["mirror"][m] - 0.05, (m, rate)
    assert rate["gpu"]["term"] >= 0.55 and rate["gpu"]["it1"] >= 0.35, rate
    # where the control flow coincides the numbers coincide: medians far inside the tolerance
    assert summary["pose_dt_m"]["median"] < 5e-4 and summary["pose_dr_rad"]["median"] < 1e-4
    assert summary["pose_dt_m"]["p99"] < POSE_TOL_T and summary["pose_dr_rad"]["p99"] < POSE_TOL_R
    assert summary["log_likelihood_rel"]["p95"] < 2e-2
