"""Developer benchmark: resident 512-pair batch, level-kernel time per step for a list of plan overrides.
usage: python scripts/dev_bench.py [batch] [steps] [spc,spc,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Config, Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spcs = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
W, H = 640, 480
dev = torch.device("cuda", 0)
eng = Engine(device=0)
hI = torch.empty((2 * B, H, W), dtype=torch.float32).pin_memory(); hZ = torch.empty((2 * B, H, W), dtype=torch.float32).pin_memory()
for i in range(B):
    p = synth.make_pair(i, synth.SceneConfig(), device=dev)
    hI[i].copy_(p["I_ref"]); hZ[i].copy_(p["Z_ref"]); hI[B + i].copy_(p["I_cur"]); hZ[B + i].copy_(p["Z_cur"])
torch.cuda.synchronize()
pyrs = eng.pyramid_batch(None, None, synth.FR1_INTRINSICS, 5, host_ptrs=(hI.data_ptr(), hZ.data_ptr(), 2 * B, H, W))
refs, curs = pyrs[:B], pyrs[B:]
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
eng.synchronize()
for spc in spcs:
    if spc: os.environ["DVO_B200_STRIPS_PER_CTA"] = str(spc)
    else: os.environ.pop("DVO_B200_STRIPS_PER_CTA", None)
    for _ in range(2): res = eng.match_batch(refs, curs, cfg, raw=True)
    eng.profile_read(reset=True)
    eng.profile_enable(True)
    t0 = time.perf_counter()
    hashes = []
    for _ in range(steps):
        res = eng.match_batch(refs, curs, cfg, raw=True)
    eng.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    import hashlib
    for _ in range(2):     # run-to-run determinism of the batch (outside the timed region)
        rr = eng.match_batch(refs, curs, cfg)
        hashes.append(hashlib.md5(b"".join(np.asarray(r.transformation).tobytes() + np.asarray(r.information).tobytes() for r in rr)).hexdigest()[:12])
    print("result hashes of two more runs:", hashes, "deterministic" if len(set(hashes)) == 1 else "NOT DETERMINISTIC")
    pix = sum((W >> res[i].levels[l].id) * (H >> res[i].levels[l].id) * res[i].levels[l].num_iterations for i in range(B) for l in range(res[i].num_levels))
    its = np.array([[res[i].levels[l].num_iterations for l in range(res[i].num_levels)] for i in range(B)])
    print("iterations per level (coarse->fine): mean", np.round(its.mean(0), 2), "max", its.max(0), "p99", np.percentile(its, 99, axis=0))
    print(f"spc={spc} batch={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} align/s  algorithmic {40 * pix / ms / 1e6:.0f} GB/s  frac {40 * pix / ms / 1e6 / 6578.3:.3f}", flush=True)
    eng.profile_read(reset=True)     # DVO_B200_TIMING=1: per-level CTA-time breakdown on stderr
