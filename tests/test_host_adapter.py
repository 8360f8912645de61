"""The reference-shaped C++ adapter (dynamic-visual-slam_b200/host/ORBextractor.hpp) over the C ABI.
CPU: it compiles against include/orbx.h, links liborbx.so and fails loudly without a GPU.
GPU: driven like the frontend drives the reference's extractor and matcher, results bit-identical to the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "dynamic-visual-slam_b200", "lib")


@pytest.fixture(scope="module")
def adapter_exe(built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adapter") / "adapter_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "adapter_check.cpp"),
                           "-o", exe, "-L" + LIBDIR, "-lorbx", "-Wl,-rpath," + LIBDIR])
    return exe


def _inputs(oracle, tmp_path, w, h):
    g0, g1 = oracle.synth_gray(5, 0, w, h), oracle.synth_gray(5, 1, w, h)
    d1 = oracle.synth_depth(5, 1, w, h)
    paths = [str(tmp_path / n) for n in ("g0.raw", "g1.raw", "d1.raw", "out.bin")]
    g0.tofile(paths[0]); g1.tofile(paths[1]); d1.tofile(paths[2])
    return g0, g1, d1, paths


def test_adapter_builds_and_fails_loudly_without_gpu(adapter_exe, oracle, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _, _, _, paths = _inputs(oracle, tmp_path, 320, 240)
    r = subprocess.run([adapter_exe, "320", "240"] + paths, capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_adapter_matches_oracle(adapter_exe, oracle, tmp_path):
    w, h = 640, 480
    g0, g1, d1, paths = _inputs(oracle, tmp_path, w, h)
    r = subprocess.run([adapter_exe, str(w), str(h)] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(paths[3], "rb").read()
    hdr = np.frombuffer(raw, np.int32, 8)
    n0, n1, n1f, nm, ng, empty_rc, levels, npyr = hdr.tolist()
    off = 32
    k0 = np.frombuffer(raw, oracle.KP_DTYPE, n0, off); off += 28 * n0
    d0 = np.frombuffer(raw, np.uint8, 32 * n0, off).reshape(n0, 32); off += 32 * n0
    k1 = np.frombuffer(raw, oracle.KP_DTYPE, n1, off); off += 28 * n1
    dd1 = np.frombuffer(raw, np.uint8, 32 * n1, off).reshape(n1, 32); off += 32 * n1
    m = np.frombuffer(raw, oracle.DM_DTYPE, nm, off); off += 16 * nm
    good = np.frombuffer(raw, oracle.DM_DTYPE, ng, off); off += 16 * ng
    sf = np.frombuffer(raw, np.float32, 8, off); off += 32
    nb = int(np.frombuffer(raw, np.int32, 1, off)[0]); off += 4
    bk = np.frombuffer(raw, oracle.KP_DTYPE, nb, off); off += 28 * nb
    bd = np.frombuffer(raw, np.uint8, 32 * nb, off).reshape(nb, 32)
    orc = oracle.COracle()
    r0, r1 = orc.extract(g0), orc.extract(g1)
    assert empty_rc == -1 and levels == 8 and npyr == 8
    assert np.array_equal(k0.view(np.uint8), r0["kps"].view(np.uint8)) and np.array_equal(d0, r0["desc"])
    assert np.array_equal(k1.view(np.uint8), r1["kps"].view(np.uint8)) and np.array_equal(dd1, r1["desc"])
    fk, _, _ = oracle.filter_depth(r1["kps"], r1["desc"], d1)
    assert n1f == len(fk) and 0 < n1f < n1
    mo = oracle.match(r1["desc"], r0["desc"])
    assert np.array_equal(m.view(np.uint8), mo.view(np.uint8))
    assert np.array_equal(good.view(np.uint8), mo[mo["distance"] < 50.0].view(np.uint8))
    assert np.array_equal(sf, orc.scale)
    gq = mo[mo["distance"] < 50.0]["queryIdx"]
    idx = oracle.cull_keyframe(r1["kps"]["response"], gq)
    assert nb == len(idx) > len(gq) and np.array_equal(bk.view(np.uint8), r1["kps"][idx].view(np.uint8)) and np.array_equal(bd, r1["desc"][idx])


def test_adapter_opencv_build_mode_compiles(built, tmp_path):
    """-DORBX_WITH_OPENCV (the branch the ROS nodes compile: cv::InputArray / cv::OutputArray / cv::noArray() call shape) against the
    OpenCV-shaped header shim (oracle/ref_shim) — the image has no OpenCV C++ headers, so this is the closest compile check available."""
    src = tmp_path / "ocv_mode.cpp"
    src.write_text('#define ORBX_WITH_OPENCV\n#include "dynamic-visual-slam_b200/host/ORBextractor.hpp"\n'
                   'int main() { orbx::ORBextractor e(1000, 1.2f, 8, 20, 7); orbx::BFMatcher m(e); cv::Mat g, d; std::vector<cv::KeyPoint> k;\n'
                   '  std::vector<int> lap = {0, 0}; int n = e(g, cv::noArray(), k, d, lap); std::vector<cv::DMatch> mm; m.match(d, d, mm);\n'
                   '  orbx::LandmarkDB db(e, 16); return n + (int)db.rows(); }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + ROOT, str(src)])


@pytest.mark.gpu
def test_same_callsite_on_reference_class_and_dropin(built, oracle, tmp_path):
    """oracle/_ref/callsite_check: ONE templated call site (the frontend's operator() call, frontend.cpp:1094-1095) instantiated on the
    reference's own compiled ORB_SLAM3::ORBextractor and on the drop-in adapter, results compared in-process, incl. the vLappingArea
    mono / stereo split (ORBextractor.cpp:1152-1166), the empty-image return and the getters."""
    exe = os.path.join(ROOT, "oracle", "_ref", "callsite_check")
    assert os.path.exists(exe), "oracle/_ref/callsite_check was not built in the container (make -C oracle _ref)"
    for (w, h, seed) in [(640, 480, 5), (1280, 720, 3)]:
        p = str(tmp_path / ("g%d.raw" % w))
        oracle.synth_gray(seed, 0, w, h).tofile(p)
        r = subprocess.run([exe, str(w), str(h), p], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("equal") >= 7 and "DIFFERENT" not in r.stdout
