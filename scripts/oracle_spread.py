"""FAITHFUL <-> MIRROR / EXACT pose spread of the oracle over seeded 640x480 pairs (sets the stated SE(3) tolerance)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from dvo_slam_b200 import synth
from oracle import oracle_py as orc
orc.lib()
def one(seed):
    p = synth.make_pair(seed); K = p["intrinsics"]
    a = {k: p[k].numpy() for k in ("I_ref", "Z_ref", "I_cur", "Z_cur")}
    r, c = orc.Pyramid(a["I_ref"], a["Z_ref"], K, 5), orc.Pyramid(a["I_cur"], a["Z_cur"], K, 5)
    cfg = orc.config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
    T = {m: orc.match(r, c, cfg, orc.mode(m))["T"] for m in ("faithful", "mirror", "exact")}
    out = []
    for m in ("mirror", "exact"):
        d = synth.se3_log(np.linalg.inv(T["faithful"]) @ T[m]); out += [np.abs(d[:3]).max(), np.abs(d[3:]).max()]
    gt = synth.se3_log(T["faithful"] @ p["T_true"]); out += [np.abs(gt[:3]).max(), np.abs(gt[3:]).max()]
    return out
seeds = range(int(sys.argv[1]), int(sys.argv[2]))
with ThreadPoolExecutor(8) as ex: res = np.array(list(ex.map(one, seeds)))
for i, name in enumerate(["mirror dT", "mirror dR", "exact dT", "exact dR", "faithful-vs-truth dT", "faithful-vs-truth dR"]):
    print("%-22s median %.2e  p90 %.2e  max %.2e" % (name, np.median(res[:, i]), np.percentile(res[:, i], 90), res[:, i].max()))
