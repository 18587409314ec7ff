"""Developer probe: does a pinned H2D copy overlap (a) a plain long kernel, (b) the persistent level kernels?"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
dev = torch.device("cuda", 0)
B = 256
eng = Engine(0)
H, W = 480, 640
I = torch.empty((2 * B, H, W), dtype=torch.float32).pin_memory(); Z = torch.empty_like(I).pin_memory()
for i in range(B):
    p = synth.make_pair(i % 32, device=dev)
    I[i].copy_(p["I_ref"]); Z[i].copy_(p["Z_ref"]); I[B + i].copy_(p["I_cur"]); Z[B + i].copy_(p["Z_cur"])
torch.cuda.synchronize()
pyr = eng.pyramid_batch(None, None, synth.FR1_INTRINSICS, 5, host_ptrs=(I.data_ptr(), Z.data_ptr(), 2 * B, H, W))
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
eng.match_batch(pyr[:B], pyr[B:], cfg, raw=True)
src = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream()
big = torch.randn(8192, 8192, device=dev)

def copy():
    with torch.cuda.stream(side):
        dst.copy_(src, non_blocking=True)
    side.synchronize()

def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3

def both(work):
    th = threading.Thread(target=copy); th.start(); work(); th.join()

def gemm():
    for _ in range(12):
        torch.mm(big, big)
def match():
    eng.match_batch(pyr[:B], pyr[B:], cfg, raw=True)
for name, work in (("gemm", gemm), ("match", match)):
    work()
    a = t(copy); b = t(work); c = t(lambda: both(work))
    print("%s: copy alone %.1f ms, work alone %.1f ms, together %.1f ms" % (name, a, b, c))
