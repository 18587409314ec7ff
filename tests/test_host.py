"""CPU tests of the host-side logic: synthetic generator, sharding, the result all-gather on a
world_size-2 gloo group (the N>1 path of bench.py without GPUs)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_synth_is_seeded_and_geometrically_consistent(small_scene):
    from dvo_slam_b200 import synth
    a, b = synth.make_pair(5, small_scene), synth.make_pair(5, small_scene)
    for k in ("I_ref", "Z_ref", "I_cur", "Z_cur"):
        x, y = a[k].numpy(), b[k].numpy()
        assert np.array_equal(np.isnan(x), np.isnan(y)) and np.array_equal(x[~np.isnan(x)], y[~np.isnan(y)])
    assert np.allclose(a["T_true"], synth.se3_exp(a["xi"]))
    assert np.abs(a["xi"][:3]).max() <= 0.03 and np.abs(a["xi"][3:]).max() <= 0.02
    Z = a["Z_ref"].numpy()
    assert 0.01 < np.isnan(Z).mean() < 0.25 and np.nanmin(Z) > 0.5 and np.nanmax(Z) <= 4.0
    I = a["I_ref"].numpy()
    assert I.min() >= 0 and I.max() <= 255 and np.array_equal(I, np.round(I))
    # depth is a multiple of 1/5000 m (TUM u16 quantisation, benchmark_slam.cpp:77)
    q = Z[~np.isnan(Z)].astype(np.float64) * 5000.0
    assert np.abs(q - np.round(q)).max() < 1e-2
    # warping a reference pixel with its depth and T_true lands on a current pixel with matching depth
    fx, fy, ox, oy = a["intrinsics"]
    v, u = 60, 80
    z = Z[v, u]
    if not np.isnan(z):
        p = a["T_true"] @ np.array([(u - ox) / fx * z, (v - oy) / fy * z, z, 1.0])
        uc, vc = fx * p[0] / p[2] + ox, fy * p[1] / p[2] + oy
        zc = a["Z_cur"].numpy()[int(round(vc)), int(round(uc))]
        assert np.isnan(zc) or abs(zc - p[2]) < 0.05


def test_se3_helpers_roundtrip():
    from dvo_slam_b200 import synth
    rng = np.random.default_rng(3)
    for _ in range(20):
        xi = rng.uniform(-0.3, 0.3, 6)
        assert np.allclose(synth.se3_log(synth.se3_exp(xi)), xi, atol=1e-10)


def test_shard_range_partitions_the_batch():
    from dvo_slam_b200.distributed import shard_range
    for total in (1, 7, 512, 4096, 4097):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_range(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 8, 3) == (1536, 2048)      # configs[3]: 512 pairs per GPU


def _gather_worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dvo_slam_b200.distributed import all_gather_results, results_to_tensor, shard_range, tensor_to_results
        from dvo_slam_b200.engine import CResult
        b, e = shard_range(total, world, rank)
        local = (CResult * (e - b))()
        for i in range(e - b):
            local[i].log_likelihood = float(b + i)
            local[i].transformation[3] = 0.5 * (b + i)
            local[i].num_levels = 5
            local[i].levels[4].num_iterations = (b + i) % 7
        g = all_gather_results(results_to_tensor(local, "cpu"), total)
        out = tensor_to_results(g)
        ok = len(out) == total and all(out[i].log_likelihood == float(i) and out[i].transformation[3] == 0.5 * i and
                                       out[i].levels[4].num_iterations == i % 7 for i in range(total))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_result_all_gather_world_size_2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(got) == [(0, True), (1, True)]


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference must print one JSON line with the contract's keys (no GPU needed)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--cpu-sample", "2"], text=True, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "alignments/s"
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in line
    from oracle import oracle_py as orc
    want = "reference" if (orc.ref_available("_O3") or orc.ref_available("")) else "port"
    assert line["cpu_baseline"]["kind"] == want and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_product_arm_fails_loudly_without_a_device():
    """The measured arm has no CPU fallback: without a CUDA device it exits non-zero and says why."""
    import subprocess
    import sys
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_c_abi_shard_range_is_the_partition_of_the_multi_process_path():
    """dvo_b200_shard_range (one process, one thread per device) and distributed.shard_range (one rank per GPU) own the
    same pair indices; without a device the sharded front end refuses to start (no CPU fallback)."""
    import ctypes as C
    import torch
    from dvo_slam_b200 import engine
    from dvo_slam_b200.distributed import shard_range
    L = engine.load_library()
    for total in (0, 1, 7, 512, 4096, 4099):
        for n in (1, 2, 3, 8):
            covered = []
            for k in range(n):
                b, e = C.c_int64(), C.c_int64()
                assert L.dvo_b200_shard_range(total, n, k, C.byref(b), C.byref(e)) == 0
                assert (b.value, e.value) == shard_range(total, n, k)
                covered += list(range(b.value, e.value))
            assert covered == list(range(total))
    b, e = C.c_int64(), C.c_int64()
    assert L.dvo_b200_shard_range(10, 2, 2, C.byref(b), C.byref(e)) != 0      # shard out of range
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no usable CUDA device"):
            engine.ShardedEngine([0, 0])


def test_scale_sum_segments_combine_like_the_reference_walk(tmp_path):
    """The pairwise scale sum of computeScaleSse (dense_tracking_impl.cpp:590-638: (w_2j + w_2j+1) r_2j r_2j^T per pair of the
    compacted list, odd tail alone) from the kernel's segment summaries: SegT / combine_seg of csrc/stages.cuh, compiled for
    the HOST, combined left to right, through random parenthesisations and in the kernel's row -> strip -> level order,
    against the sequential walk (tests/native/seg_monoid.cu)."""
    import shutil
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "seg_monoid")
    r = subprocess.run([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-w", "-I", os.path.join(ROOT, "dvo_slam_b200", "csrc"),
                        "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "native", "seg_monoid.cu")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:]
