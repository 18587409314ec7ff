"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference ships no golden vectors and cannot be built here, so these fixtures are
produced by the oracle (oracle/liboracle.so) on seeded synthetic inputs: they pin the oracle against
regressions and give the GPU tests fixed inputs; they are NOT reference outputs.  The FAITHFUL-mode numbers in
them are, however, what the reference's own object code produces on the same inputs, bit for bit
(tests/test_reference_pin.py runs both on these fixtures).
Inputs are stored as uint8 intensity + uint16 raw depth (TUM style, 1/5000 m), i.e. exactly what
benchmark_slam.cpp:46-93 would load from disk.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dvo_slam_b200 import synth  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LEVELS = 3


def main():
    cfg_s = synth.SceneConfig(width=160, height=120, intrinsics=tuple(v / 4 for v in synth.FR1_INTRINSICS))
    for seed in (11, 12, 13):
        p = synth.make_pair(seed, cfg_s)
        raw = {}
        for k in ("ref", "cur"):
            I = p[f"I_{k}"].numpy()
            Z = p[f"Z_{k}"].numpy()
            raw[f"grey_{k}"] = I.astype(np.uint8)
            zr = np.where(np.isnan(Z), 0, np.round(Z * 5000.0)).astype(np.uint16)
            raw[f"depth_{k}"] = zr
        K = cfg_s.intrinsics
        Ir, Ic = raw["grey_ref"].astype(np.float32), raw["grey_cur"].astype(np.float32)
        Zr, Zc = orc.convert_raw_depth(raw["depth_ref"], 1.0 / 5000.0), orc.convert_raw_depth(raw["depth_cur"], 1.0 / 5000.0)
        ref, cur = orc.Pyramid(Ir, Zr, K, LEVELS), orc.Pyramid(Ic, Zc, K, LEVELS)
        out = dict(raw)
        out["intrinsics"] = np.array(K, dtype=np.float32)
        out["T_true"] = p["T_true"]
        ocfg = orc.config(first_level=2, last_level=0, max_iterations_per_level=50, precision=1e-4)
        for mname in ("faithful", "exact", "mirror"):
            r = orc.match(ref, cur, ocfg, orc.mode(mname))
            out[f"{mname}_T"] = r["T"]
            out[f"{mname}_information"] = r["information"]
            out[f"{mname}_ll"] = np.array(r["log_likelihood"])
            out[f"{mname}_levels"] = np.array([[l["id"], l["termination"], l["valid_pixels"], l["num_iterations"]] for l in r["levels"]])
            out[f"{mname}_iter_n"] = np.array([it["n"] for it in r["iterations"]])
            out[f"{mname}_iter_nll"] = np.array([it["nll"] for it in r["iterations"]])
        # one linearisation KAT per level at a fixed transform (90 % of the true motion)
        T = synth.se3_exp(p["xi"] * 0.9)
        out["kat_T"] = T
        pp = np.array([[2000.0, -30.0], [-30.0, 9000.0]], dtype=np.float32)
        out["kat_prev_precision"] = pp
        for mname in ("faithful", "mirror"):
            for lvl in range(LEVELS):
                for uw in (0, 1):
                    lin = orc.linearize(ref, cur, lvl, T, orc.mode(mname), bool(uw), pp)
                    key = f"kat_{mname}_l{lvl}_w{uw}"
                    out[key + "_n"] = np.array(lin["n"])
                    out[key + "_P"] = lin["precision"]
                    out[key + "_ll"] = np.array(lin["ll"])
                    out[key + "_A"] = lin["A"]
                    out[key + "_b"] = lin["b"]
        # pyramid checksums (sum of finite values per plane per level) and selection counts
        for lvl in range(LEVELS):
            pl = ref.planes(lvl)
            out[f"pyr_l{lvl}_sum"] = np.array([np.nansum(pl[c].astype(np.float64)) for c in range(6)])
            out[f"pyr_l{lvl}_nan"] = np.array([int(np.isnan(pl[c]).sum()) for c in range(6)])
            out[f"sel_l{lvl}"] = np.array(orc.select(ref, lvl)[0])
        np.savez_compressed(os.path.join(OUT, f"pair_{seed}.npz"), **out)
        print("wrote", seed, {m: out[f"{m}_levels"][:, 3].tolist() for m in ("faithful", "exact", "mirror")})


if __name__ == "__main__":
    main()
