import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
from oracle import oracle_py as orc
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = synth.make_pair(seed); K = p["intrinsics"]
a = {k: p[k].numpy() for k in ("I_ref", "Z_ref", "I_cur", "Z_cur")}
eng = Engine(0)
gr, gc = eng.pyramid(a["I_ref"], a["Z_ref"], K, 5), eng.pyramid(a["I_cur"], a["Z_cur"], K, 5)
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
r = eng.match_batch([gr] * N, [gc] * N, cfg, with_iterations=True)[N - 1]
if os.environ.get("DVO_ORACLE"):
    o = orc.match(orc.Pyramid(a["I_ref"], a["Z_ref"], K, 5), orc.Pyramid(a["I_cur"], a["Z_cur"], K, 5),
                  orc.config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4), orc.mode("mirror"))
    its = o["iterations"]
else:
    its = None
print("RPW", os.environ.get("DVO_B200_RPW"), "N", N, "ll", r.log_likelihood, [l["num_iterations"] for l in r.levels])
for k, it in enumerate(r.iterations):
    s = "  L%d it%d n=%d nll=%.3f P=%s" % (it["level"], it["id"], it["n"], it["nll"], np.array2string(it["precision"].ravel(), precision=3))
    if its and k < len(its): s += "   | oracle n=%d nll=%.3f P=%s" % (its[k]["n"], its[k]["nll"], np.array2string(its[k]["precision"].ravel(), precision=3))
    if it["level"] <= 1: print(s)
