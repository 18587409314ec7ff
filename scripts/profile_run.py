"""Short workload for ncu: a batch of 640x480 pairs, one warm match_batch, one profiled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
first = int(sys.argv[3]) if len(sys.argv) > 3 else 4
last = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = Engine(0)
dev = torch.device("cuda", 0)
I = torch.empty((2 * B, 480, 640), dtype=torch.float32).pin_memory(); Z = torch.empty_like(I).pin_memory()
for i in range(B):
    p = synth.make_pair(i, device=dev)
    I[i].copy_(p["I_ref"]); Z[i].copy_(p["Z_ref"]); I[B + i].copy_(p["I_cur"]); Z[B + i].copy_(p["Z_cur"])
torch.cuda.synchronize()
pyr = eng.pyramid_batch(None, None, synth.FR1_INTRINSICS, 5, host_ptrs=(I.data_ptr(), Z.data_ptr(), 2 * B, 480, 640))
cfg = Config(first_level=first, last_level=last, max_iterations_per_level=50, precision=1e-4)
for _ in range(reps + 1):
    res = eng.match_batch(pyr[:B], pyr[B:], cfg, raw=True)
print("done", res[0].num_iterations_total)
