"""N3 row: TUM RGB-D replay without ROS (dvo_slam_b200/host/tum_replay.cpp) -- association / ground-truth readers,
PNG loader, trajectory writer, and (GPU) the batched frame-to-frame odometry against a synthetic sequence."""
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dvo_slam_b200", "host")


@pytest.fixture(scope="module")
def replay_bin():
    import __graft_entry__ as ge
    ge.build_cuda()
    ge.build_host()
    return os.path.join(HOST, "tum_replay")


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, arr):
    """uint8 (h,w) / (h,w,3) or uint16 (h,w) -> PNG, cycling through all five scanline filters."""
    arr = np.asarray(arr)
    h, w = arr.shape[:2]
    ch = 1 if arr.ndim == 2 else arr.shape[2]
    bits = 16 if arr.dtype == np.uint16 else 8
    color = {1: 0, 3: 2, 4: 6}[ch]
    rows = arr.astype(">u2").tobytes() if bits == 16 else arr.astype(np.uint8).tobytes()
    bpp = ch * bits // 8
    stride = w * bpp
    raw = bytearray()
    prev = bytes(stride)
    for y in range(h):
        cur = rows[y * stride:(y + 1) * stride]
        f = y % 5
        out = bytearray(stride)
        for i in range(stride):
            a = cur[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            out[i] = (cur[i] - pred) & 255
        raw.append(f)
        raw += out
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    comp = zlib.compress(bytes(raw), 6)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, color, 0, 0, 0)))
        f.write(chunk(b"IDAT", comp[: len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b""))


def _quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])


def write_sequence(folder, frames_rgb, frames_depth_u16, stamps, poses=None):
    os.makedirs(os.path.join(folder, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(folder, "depth"), exist_ok=True)
    with open(os.path.join(folder, "assoc.txt"), "w") as f:
        f.write("# color images\n# timestamp filename timestamp filename\n")
        for k, t in enumerate(stamps):
            write_png(os.path.join(folder, "rgb", f"{t:.6f}.png"), frames_rgb[k])
            write_png(os.path.join(folder, "depth", f"{t + 0.01:.6f}.png"), frames_depth_u16[k])
            f.write(f"{t:.6f} rgb/{t:.6f}.png {t + 0.01:.6f} depth/{t + 0.01:.6f}.png\n")
    if poses is not None:
        with open(os.path.join(folder, "groundtruth.txt"), "w") as f:
            f.write("# ground truth trajectory\n# timestamp tx ty tz qx qy qz qw\n")
            f.write("%.4f 9 9 9 0 0 0 1\n" % (stamps[0] - 0.5))         # an entry before the first frame: must be skipped
            for t, P in zip(stamps, poses):
                q = _quat(P[:3, :3])
                f.write("%.4f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % (t + 0.001, P[0, 3], P[1, 3], P[2, 3], q[0], q[1], q[2], q[3]))
    return os.path.join(folder, "assoc.txt")


def test_readers_and_loader_without_a_device(replay_bin, tmp_path):
    rng = np.random.default_rng(3)
    h, w = 48, 64
    rgb = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for _ in range(3)]
    depth = [rng.integers(0, 40000, size=(h, w), dtype=np.uint16) for _ in range(3)]
    depth[0][rng.random((h, w)) < 0.1] = 0
    stamps = [1305031102.175304, 1305031102.211214, 1305031102.243211]
    P0 = np.eye(4); P0[:3, 3] = [1.25, -0.5, 0.75]
    c, s = np.cos(0.3), np.sin(0.3)
    P0[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    assoc = write_sequence(str(tmp_path), rgb, depth, stamps, [P0, P0, P0])
    r = subprocess.run([replay_bin, "--assoc", assoc, "--groundtruth", str(tmp_path / "groundtruth.txt"), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["pairs"] == 3 and out["groundtruth"] == 4 and out["gt_first"] == 1       # comments skipped, closest entry after the stamp
    # ros::Time::fromSec splits the parsed double (sec = floor, nsec = round of the rest) and prints sec.nsec with nine
    # digits: 1305031102.175304 comes out as ...175303936, exactly as in the reference's trajectory files
    sec = int(np.floor(stamps[0]))
    assert out["first_stamp"] == "%d.%09d" % (sec, int(round((stamps[0] - sec) * 1e9)))
    assert out["rgb"] == [w, h, 3, 8] and out["depth"] == [w, h, 1, 16]
    R, G, B = (rgb[0][..., k].astype(np.int64) for k in range(3))
    grey = (B * 1868 + G * 9617 + R * 4899 + 8192) >> 14                                # OpenCV 8-bit BGR2GRAY
    assert out["grey_sum"] == float(grey.sum())
    z = depth[0].astype(np.float32) * np.float32(1.0 / 5000.0)
    assert out["depth_nan"] == int((depth[0] == 0).sum())
    assert abs(out["depth_sum"] - float(z[depth[0] != 0].astype(np.float64).sum())) < 1e-6 * out["depth_sum"]
    assert np.allclose(out["pose0"][:3], P0[:3, 3]) and np.allclose(out["pose0"][3:], _quat(P0[:3, :3]), atol=1e-8)
    # error behaviour: a missing association file is an error, not an empty trajectory
    r = subprocess.run([replay_bin, "--assoc", str(tmp_path / "nope.txt")], capture_output=True, text=True)
    assert r.returncode == 2 and "no entries" in r.stderr


@pytest.mark.gpu
def test_replay_follows_a_synthetic_sequence(replay_bin, tmp_path, oracle):
    from dvo_slam_b200 import synth
    from helpers import POSE_TOL_R, POSE_TOL_T, pose_delta
    cfg = synth.SceneConfig(width=320, height=240, intrinsics=tuple(v / 2 for v in synth.FR1_INTRINSICS))
    n = 7
    frames, poses = synth.make_sequence(5, n, cfg)
    rgb = [np.repeat(f[0].numpy().astype(np.uint8)[..., None], 3, axis=2) for f in frames]      # grey stored as R=G=B
    depth = [np.where(np.isnan(f[1].numpy()), 0, np.round(f[1].numpy() * 5000.0)).astype(np.uint16) for f in frames]
    stamps = [100.0 + 0.033 * k for k in range(n)]
    assoc = write_sequence(str(tmp_path), rgb, depth, stamps, poses)
    traj = str(tmp_path / "traj.txt")
    r = subprocess.run([replay_bin, "--assoc", assoc, "--groundtruth", str(tmp_path / "groundtruth.txt"), "--out", traj, "--batch", "4",
                        "--first", "2", "--last", "0", "--intrinsics"] + [repr(float(v)) for v in cfg.intrinsics],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    summary = json.loads(r.stderr.strip().splitlines()[-1])
    assert summary["frames"] == n and summary["alignments"] == n - 1 and summary["failed"] == 0
    lines = open(traj).read().splitlines()
    assert len(lines) == n - 1 and all(l.endswith(" ") and len(l.split()) == 8 for l in lines)   # `ts tx ty tz qx qy qz qw `
    for k, l in enumerate(lines, start=1):
        v = [float(x) for x in l.split()]
        assert abs(v[0] - stamps[k]) < 1e-6
        # against the TRUE camera path (not the oracle): the estimator itself is a few mm off per alignment at 320x240
        assert np.linalg.norm(np.array(v[1:4]) - poses[k][:3, 3]) < 5e-3 * k
        q = _quat(poses[k][:3, :3])
        assert min(np.linalg.norm(np.array(v[4:]) - q), np.linalg.norm(np.array(v[4:]) + q)) < 2e-3 * k
    # ... and against an oracle-FAITHFUL replay of the same files' content: the reference's loader arithmetic (8-bit grey ->
    # float, u16 * (1/5000)f, 0 -> NaN; benchmark_slam.cpp:58-77), DenseTracker defaults (MaxIterationsPerLevel 100,
    # Precision 5e-7), trajectory = trajectory * Result.Transformation (benchmark.cpp:463).  Per alignment the GPU engine and
    # FAITHFUL agree to the stated SE(3) tolerance; the accumulated pose may drift by that much per frame.
    ocfg = oracle.config(first_level=2, last_level=0, max_iterations_per_level=100, precision=5e-7)
    fa = oracle.mode("faithful")
    pyr = []
    for k in range(n):
        grey = rgb[k][..., 0].astype(np.float32)
        z = np.where(depth[k] == 0, np.float32("nan"), depth[k].astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
        pyr.append(oracle.Pyramid(grey, z, cfg.intrinsics, 3))
    traj_o = np.eye(4)          # first ground-truth pose of the file = identity
    prev = np.eye(4)
    for k in range(1, n):
        rel = oracle.match(pyr[k - 1], pyr[k], ocfg, fa)["T"]
        traj_o = traj_o @ rel
        v = [float(x) for x in lines[k - 1].split()]
        q = _quat(traj_o[:3, :3])
        assert np.linalg.norm(np.array(v[1:4]) - traj_o[:3, 3]) < POSE_TOL_T * k
        assert min(np.linalg.norm(np.array(v[4:]) - q), np.linalg.norm(np.array(v[4:]) + q)) < POSE_TOL_R * k
        # the relative motion of this frame alone: replay's pose_k = pose_{k-1} * rel_gpu
        x, y, zq, w = v[4:]
        Rg = np.array([[1 - 2 * (y * y + zq * zq), 2 * (x * y - zq * w), 2 * (x * zq + y * w)],
                       [2 * (x * y + zq * w), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - x * w)],
                       [2 * (x * zq - y * w), 2 * (y * zq + x * w), 1 - 2 * (x * x + y * y)]])
        Tg = np.eye(4); Tg[:3, :3] = Rg; Tg[:3, 3] = v[1:4]
        rel_g = np.linalg.inv(prev) @ Tg
        dt, dr = pose_delta(rel, rel_g)
        assert dt < POSE_TOL_T and dr < POSE_TOL_R
        prev = Tg
