"""80 pairs 640x480x5 through the raw-input pyramid path and one match_batch (fused two-segment launch), for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
B = 80
eng = Engine(0)
p = [synth.make_pair(i) for i in range(4)]          # rendered on the CPU: nothing but this library runs on the GPU
G = torch.empty((2 * B, 480, 640), dtype=torch.uint8).pin_memory(); D = torch.empty((2 * B, 480, 640), dtype=torch.uint16).pin_memory()
for i in range(B):
    q = p[i % 4]
    for k, (I, Z) in enumerate(((q["I_ref"], q["Z_ref"]), (q["I_cur"], q["Z_cur"]))):
        G[k * B + i].copy_(I.to(torch.uint8))
        D[k * B + i].copy_(torch.where(torch.isnan(Z), torch.zeros_like(Z), torch.round(Z * 5000.0)).to(torch.int32).to(torch.uint16))
pyr = eng.pyramid_raw_batch((G.data_ptr(), D.data_ptr(), 2 * B, 480, 640), 1.0 / 5000.0, synth.FR1_INTRINSICS, 5)
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
res = eng.match_batch(pyr[:B], pyr[B:], cfg, raw=True)
print("done", res[0].num_iterations_total, res[B - 1].num_iterations_total)
