"""Developer probe: where does the e2e step go?  H2D bandwidth, pyramid build from raw host images, match per chunk."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
eng = Engine(0)
H, W = 480, 640
hG = torch.empty((2 * B, H, W), dtype=torch.uint8).pin_memory()
hD = torch.empty((2 * B, H, W), dtype=torch.uint16).pin_memory()
for i in range(B):
    p = synth.make_pair(i % 64, device=dev)
    for off, I, Z in ((i, p["I_ref"], p["Z_ref"]), (B + i, p["I_cur"], p["Z_cur"])):
        hG[off].copy_(I.to(torch.uint8))
        hD[off].copy_(torch.where(torch.isnan(Z), torch.zeros_like(Z), torch.round(Z * 5000.0)).to(torch.int32).to(torch.uint16))
torch.cuda.synchronize()
# 1. raw H2D bandwidth from pinned memory
dG = torch.empty_like(hG, device=dev); dD = torch.empty_like(hD, device=dev)
for _ in range(2):
    t0 = time.perf_counter(); dG.copy_(hG, non_blocking=True); dD.copy_(hD, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("H2D %.1f MB in %.2f ms -> %.1f GB/s" % ((hG.numel() + 2 * hD.numel()) / 1e6, (t1 - t0) * 1e3, (hG.numel() + 2 * hD.numel()) / (t1 - t0) / 1e9))
del dG, dD
K = synth.FR1_INTRINSICS
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
for nch in (1, 2, 4):
    CH = B // nch
    for rep in range(2):
        tp = tm = 0.0
        for k in range(nch):
            o = k * CH
            t0 = time.perf_counter()
            pr = eng.pyramid_raw_batch((hG[o].data_ptr(), hD[o].data_ptr(), CH, H, W), 1.0 / 5000.0, K, 5)
            pc = eng.pyramid_raw_batch((hG[B + o].data_ptr(), hD[B + o].data_ptr(), CH, H, W), 1.0 / 5000.0, K, 5)
            eng.synchronize(); t1 = time.perf_counter()
            eng.match_batch(pr, pc, cfg, raw=True)
            eng.synchronize(); t2 = time.perf_counter()
            for p in pr + pc:
                p.release()
            tp += t1 - t0; tm += t2 - t1
    print("chunks=%d: pyramids (H2D + build) %.2f ms, match %.2f ms per %d pairs" % (nch, tp * 1e3, tm * 1e3, B))
