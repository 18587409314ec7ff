"""batch vs single vs repeated runs: where do they diverge?  (developer check, GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Config, Engine
eng = Engine(device=0)
pairs = [synth.make_pair(s) for s in range(4)]
K = pairs[0]["intrinsics"]
st = lambda k: np.stack([p[k].numpy() for p in pairs])
refs = eng.pyramid_batch(st("I_ref"), st("Z_ref"), K, 5); curs = eng.pyramid_batch(st("I_cur"), st("Z_cur"), K, 5)
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
def sig(r):
    return [(it["level"], it["id"], it["n"], float(it["nll"])) for it in r.iterations]
runs = [eng.match_batch(refs, curs, cfg, with_iterations=True) for _ in range(3)]
for i in range(4):
    single = [eng.match(refs[i], curs[i], cfg, with_iterations=True) for _ in range(2)]
    sigs = [sig(r[i]) for r in runs] + [sig(s) for s in single]
    same = [s == sigs[0] for s in sigs]
    print("pair", i, "batch runs equal:", same[:3], "single runs equal to batch[0]:", same[3:], "single==single:", sigs[3] == sigs[4])
    if not all(same):
        for k in range(max(len(s) for s in sigs)):
            row = [s[k] if k < len(s) else None for s in sigs]
            if any(x != row[0] for x in row):
                print("  first difference at iteration index", k)
                for x in row: print("    ", x)
                break
