"""The C++ adapter (include/dvo/, libdvo_core_b200.so): the reference's class surface on the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import POSE_TOL_R, POSE_TOL_T, pose_delta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dvo_slam_b200", "host")


@pytest.fixture(scope="module")
def selftest_bin():
    import __graft_entry__ as ge
    ge.build_cuda()
    ge.build_host()
    return os.path.join(HOST, "selftest")


def _write_pair(tmp_path, pair):
    path = tmp_path / "pair.bin"
    with open(path, "wb") as f:
        for k in ("I_ref", "Z_ref", "I_cur", "Z_cur"):
            f.write(np.ascontiguousarray(pair[k].numpy(), dtype=np.float32).tobytes())
    return str(path)


def test_adapter_builds_and_fails_loudly_without_a_device(selftest_bin, tmp_path, small_scene):
    import torch
    from dvo_slam_b200 import synth
    assert os.path.exists(os.path.join(ROOT, "dvo_slam_b200", "libdvo_core_b200.so"))
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pair = synth.make_pair(3, small_scene)
    K = small_scene.intrinsics
    r = subprocess.run([selftest_bin, _write_pair(tmp_path, pair), str(small_scene.width), str(small_scene.height)] + [repr(float(v)) for v in K] + ["2", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr      # no CPU fallback behind the class API


def test_adapter_headers_keep_the_reference_surface():
    """Names a dvo_benchmark / dvo_slam translation unit uses must exist in the adapter headers."""
    hdr = open(os.path.join(ROOT, "include", "dvo", "dense_tracking.h")).read()
    for name in ("class DenseTracker", "struct Config", "struct TerminationCriteria", "struct IterationStats", "struct LevelStats", "struct Result",
                 "getDefaultConfig", "configuration()", "void configure(", "computeIntensityErrorImage", "HasIterationWithIncrement",
                 "LastIterationWithIncrement", "clearStatistics", "isNaN", "setIdentity", "MaxIterationsPerLevel", "UseInitialEstimate",
                 "IntensityDerivativeThreshold", "InfluenceFuntionType", "ScaleEstimatorParam"):
        assert name in hdr, name
    ev = open(os.path.join(ROOT, "include", "dvo_slam", "tracking_result_evaluation.h")).read()
    for name in ("class TrackingResultEvaluation", "class LogLikelihoodTrackingResultEvaluation", "class NormalizedLogLikelihoodTrackingResultEvaluation",
                 "class EntropyRatioTrackingResultEvaluation", "ratioWithFirst", "ratioWithAverage", "void add("):
        assert name in ev, name
    assert hdr.count("bool match(") == 4
    img = open(os.path.join(ROOT, "include", "dvo", "core", "rgbd_image.h")).read()
    for name in ("class RgbdCameraPyramid", "class RgbdImagePyramid", "class RgbdImage", "RgbdImagePyramidPtr create(", "void build(", "void compute(",
                 "RgbdImage& level(", "double timestamp", "cv::Mat intensity", " rgb;"):
        assert name in img, name


@pytest.mark.gpu
def test_adapter_matches_like_the_reference_callers(selftest_bin, tmp_path, oracle):
    from dvo_slam_b200 import synth
    pair = synth.make_pair(21)
    K = pair["intrinsics"]
    r = subprocess.run([selftest_bin, _write_pair(tmp_path, pair), "640", "480"] + [repr(float(v)) for v in K] + ["3", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    T = np.array(out["T"]).reshape(4, 4)
    a = {k: pair[k].numpy() for k in ("I_ref", "Z_ref", "I_cur", "Z_cur")}
    oref, ocur = oracle.Pyramid(a["I_ref"], a["Z_ref"], K, 4), oracle.Pyramid(a["I_cur"], a["Z_cur"], K, 4)
    fa = oracle.match(oref, ocur, oracle.config(first_level=3, last_level=1, max_iterations_per_level=50, precision=1e-4), oracle.mode("faithful"))
    dt, dr = pose_delta(fa["T"], T)
    assert dt < 2 * POSE_TOL_T and dr < POSE_TOL_R          # stops at level 1 (320x240)
    assert out["nan"] == 0 and [l["id"] for l in out["levels"]] == [3, 2, 1]
    assert [l["valid"] for l in out["levels"]] == [l["valid_pixels"] for l in fa["levels"]]
    assert all(l["n_last"] > 1000 for l in out["levels"])
    assert np.allclose(out["second_t"], T[:3, 3], atol=1e-12)        # a copy-constructed tracker gives the same answer
    assert out["err_sum"] > 0 and out["level1_w"] == 320
    assert "Level: 3" in r.stderr and "Termination:" in r.stderr      # operator<< of Stats
    # N1: LocalTracker's pair of alignments and the validator's proposal loop as one batched call each
    assert out["batch_equal"] == 1 and out["proposals_equal"] == 1
    # N4: keyframe-selection scores (tracking_result_evaluation.cpp:26-62) on identical results are exactly 1
    assert out["entropy_ratio_first"] == 1.0 and out["entropy_ratio_avg"] == 1.0 and out["ll_ratio"] == 1.0 and out["nll_ratio"] == 1.0
    info = np.array(fa["information"])
    # Information follows the (chaotic, pair-bug-sensitive) scale estimate of the last iteration: entries agree to ~1e-2 of
    # the largest one with FAITHFUL, i.e. log det to a few percent (measured 89.35 vs 90.49)
    expected = np.log(np.linalg.det(info))
    assert np.isfinite(out["logdet"]) and abs(out["logdet"] - expected) < 0.03 * abs(expected)
