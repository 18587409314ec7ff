import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import *
from dvo_slam_b200.engine import Engine, Config
from oracle import oracle_py as orc
eng = Engine(0)
g = load_golden(11); im = golden_images(g, orc)
gref = eng.pyramid(im["I_ref"], im["Z_ref"], g["K"], 3); gcur = eng.pyramid(im["I_cur"], im["Z_cur"], g["K"], 3)
oref = orc.Pyramid(im["I_ref"], im["Z_ref"], g["K"], 3); ocur = orc.Pyramid(im["I_cur"], im["Z_cur"], g["K"], 3)
for lvl in range(3):
    n_g, img_g = eng.residual_image(gref, gcur, lvl, g["kat_T"])
    n_o, img_o = orc.residual_image(oref, ocur, lvl, g["kat_T"], orc.mode("mirror"))
    print(lvl, n_g, n_o, nan_equal(img_g, img_o), np.isnan(img_g[0]).sum(), np.isnan(img_o[0]).sum())
    if not nan_equal(img_g, img_o):
        d = np.isnan(img_g[0]) != np.isnan(img_o[0]); ys, xs = np.nonzero(d); print("mismatch validity", len(ys), list(zip(ys[:8], xs[:8])))
    lg = eng.linearize(gref, gcur, lvl, g["kat_T"], True, g["kat_prev_precision"]); lo = orc.linearize(oref, ocur, lvl, g["kat_T"], orc.mode("mirror"), True, g["kat_prev_precision"])
    print("  lin n", lg["n"], lo["n"], "P", lg["precision"].ravel(), lo["precision"].ravel(), "ll", lg["ll"], lo["ll"])
cfg = Config(first_level=2, last_level=0, max_iterations_per_level=50, precision=1e-4)
r = eng.match(gref, gcur, cfg)
print([l["num_iterations"] for l in r.levels], g["mirror_levels"][:, 3])
