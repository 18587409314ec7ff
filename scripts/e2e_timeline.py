"""Developer probe: timeline of the e2e pipeline of bench.py (loader thread/context + tracker thread/context), host timestamps."""
import sys, os, time, queue, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
B = 512
dev = torch.device("cuda", 0)
tr, ld = Engine(0), Engine(0)
H, W = 480, 640
hG = torch.empty((2 * B, H, W), dtype=torch.uint8).pin_memory()
hD = torch.empty((2 * B, H, W), dtype=torch.uint16).pin_memory()
for i in range(B):
    p = synth.make_pair(i, device=dev)
    for off, I, Z in ((i, p["I_ref"], p["Z_ref"]), (B + i, p["I_cur"], p["Z_cur"])):
        hG[off].copy_(I.to(torch.uint8))
        hD[off].copy_(torch.where(torch.isnan(Z), torch.zeros_like(Z), torch.round(Z * 5000.0)).to(torch.int32).to(torch.uint16))
torch.cuda.synchronize()
K = synth.FR1_INTRINSICS
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
ev = []
T0 = time.perf_counter()
def now(): return (time.perf_counter() - T0) * 1e3
def loader(steps, q):
    for s in range(steps):
        t0 = now()
        pyr = ld.pyramid_raw_batch((hG.data_ptr(), hD.data_ptr(), 2 * B, H, W), 1.0 / 5000.0, K, 5)
        t1 = now()
        ld.synchronize()
        t2 = now()
        q.put(pyr)
        ev.append(("L", s, t0, t1, t2, now()))
def tracker(steps, q):
    for s in range(steps):
        t0 = now(); pyr = q.get(); t1 = now()
        tr.match_batch(pyr[:B], pyr[B:], cfg, raw=True); t2 = now()
        for p in pyr: p.release()
        ev.append(("T", s, t0, t1, t2, now()))
for sync_loader in (True,):
    q = queue.Queue(maxsize=1)
    a = threading.Thread(target=loader, args=(7, q)); b = threading.Thread(target=tracker, args=(7, q))
    a.start(); b.start(); a.join(); b.join()
for e in sorted(ev, key=lambda e: e[2]):
    if e[0] == "L": print("loader  step %d: enqueue %.2f..%.2f  gpu done %.2f  (upload+build %.2f ms)  put done %.2f" % (e[1], e[2], e[3], e[4], e[4] - e[2], e[5]))
    else: print("tracker step %d: wait %.2f..%.2f  match done %.2f (%.2f ms)  released %.2f (%.2f ms)" % (e[1], e[2], e[3], e[4], e[4] - e[3], e[5], e[5] - e[4]))
