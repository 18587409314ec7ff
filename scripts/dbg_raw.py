import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
eng = Engine(0); eng2 = Engine(0)
B = 8
I = torch.empty((2 * B, 480, 640), dtype=torch.float32).pin_memory(); Z = torch.empty_like(I).pin_memory()
for i in range(B):
    p = synth.make_pair(i, device="cuda:0")
    I[i].copy_(p["I_ref"]); Z[i].copy_(p["Z_ref"]); I[B + i].copy_(p["I_cur"]); Z[B + i].copy_(p["Z_cur"])
torch.cuda.synchronize()
G = I.to(torch.uint8).pin_memory()
raw = torch.where(torch.isnan(Z), torch.zeros_like(Z), torch.round(Z * 5000.0)).to(torch.int32)
D = raw.to(torch.uint16).pin_memory()
Z.copy_(torch.where(raw == 0, torch.full_like(Z, float("nan")), raw.to(torch.float32) * torch.tensor(1.0 / 5000.0, dtype=torch.float32)))
K = synth.FR1_INTRINSICS
pf = eng.pyramid_batch(None, None, K, 5, host_ptrs=(I.data_ptr(), Z.data_ptr(), 2 * B, 480, 640)); eng.synchronize()
pr = eng2.pyramid_raw_batch((G.data_ptr(), D.data_ptr(), 2 * B, 480, 640), 1.0 / 5000.0, K, 5); eng2.synchronize()
for l in (0, 2):
    a, b = pf[0].download(l), pr[0].download(l)
    same = [bool(np.array_equal(np.isnan(a[c]), np.isnan(b[c])) and np.array_equal(a[c][~np.isnan(a[c])], b[c][~np.isnan(b[c])])) for c in range(6)]
    print("level", l, same, "max |dZ|", np.nanmax(np.abs(a[1] - b[1])))
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
r1 = eng.match_batch(pf[:B], pf[B:], cfg, raw=True); r2 = eng2.match_batch(pr[:B], pr[B:], cfg, raw=True); r3 = eng.match_batch(pf[:4], pf[B:B+4], cfg, raw=True)
for i in range(3):
    print(i, r1[i].log_likelihood, r2[i].log_likelihood, r3[i].log_likelihood, [r1[i].levels[k].last_valid_constraints for k in range(5)], [r2[i].levels[k].last_valid_constraints for k in range(5)], [r3[i].levels[k].last_valid_constraints for k in range(5)])
