"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/orbx.h declares, keeps the
cv::KeyPoint / cv::DMatch layouts, and fails LOUDLY (status + message, no fallback) when no CUDA device exists."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "orbx.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_two_seams():
    names = declared_symbols()
    for must in ("orbx_create", "orbx_destroy", "orbx_extract", "orbx_extract_filtered", "orbx_extract_batch",
                 "orbx_extract_batch_device", "orbx_track_batch", "orbx_match", "orbx_match_device",
                 "orbx_db_create", "orbx_db_query_top2", "orbx_db_query_radius", "orbx_merge_top2_device",
                 "orbx_get_scale_factors", "orbx_get_pyramid_level", "orbx_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    import orbx
    L = orbx.load()
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, "declared in include/orbx.h but not exported by liborbx.so: %s" % missing


def test_binding_covers_every_declared_symbol(built):
    import orbx
    L = orbx.load()
    unbound = [n for n in declared_symbols() if getattr(L, n).argtypes is None]
    assert not unbound, "no ctypes signature for: %s" % unbound


def test_no_torch_or_opencv_types_in_the_abi():
    src = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for banned in ("at::", "torch::", "cv::", "std::", "cudaStream_t", "Tensor"):
        assert banned not in code, banned


def test_header_compiles_as_c_and_layouts(tmp_path):
    """The header is plain C (gcc -std=c99) and the POD layouts are cv::KeyPoint (28 B) / cv::DMatch (16 B)."""
    c = tmp_path / "t.c"
    c.write_text('#include "orbx.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                 'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(orbx_keypoint), sizeof(orbx_dmatch), sizeof(orbx_box),'
                 'sizeof(orbx_top2), sizeof(orbx_params), offsetof(orbx_keypoint, octave), offsetof(orbx_dmatch, distance));return 0;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out == ["28", "16", "40", "16", "64", "20", "12"]
    import orbx
    assert ct.sizeof(orbx.Params) == 64
    assert orbx.KP_DTYPE.itemsize == 28 and orbx.DM_DTYPE.itemsize == 16


def test_default_params_are_the_reference_literals(built):
    """frontend.cpp:205-211 (1000, 1.2f, 8, 20, 7) and :241-242 (0.3, 3.0)."""
    import orbx
    L = orbx.load()
    p = orbx.Params()
    L.orbx_default_params(ct.byref(p))
    assert (p.nfeatures, p.nlevels, p.ini_th_fast, p.min_th_fast) == (1000, 8, 20, 7)
    assert np.float32(p.scale_factor) == np.float32(1.2) and np.float32(p.depth_min) == np.float32(0.3) and p.depth_max == 3.0
    assert (p.max_width, p.max_height, p.max_batch) == (1280, 720, 1)


def test_create_fails_loudly_without_a_gpu(built):
    """No CPU fallback: without a CUDA device orbx_create returns a status and a message, and the Python mirror raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import orbx
    L = orbx.load()
    p = orbx.Params()
    L.orbx_default_params(ct.byref(p))
    h = ct.c_void_p()
    st = L.orbx_create(ct.byref(p), ct.byref(h))
    assert st == orbx.E_CUDA and not h.value
    assert b"no CPU fallback" in L.orbx_last_error(None)
    with pytest.raises(orbx.OrbxError):
        orbx.ORBextractor()
    # argument validation happens before any device work
    p.nlevels = 0
    assert L.orbx_create(ct.byref(p), ct.byref(h)) == orbx.E_INVALID
    assert L.orbx_create(None, ct.byref(h)) == orbx.E_INVALID
    # null handles are rejected, not dereferenced
    assert L.orbx_sync(None) == orbx.E_INVALID
    assert L.orbx_extract(None, None, 0, 0, 0, None, None, 0, None) == orbx.E_INVALID
    assert L.orbx_launch_count(None) == 0


def test_product_does_not_link_the_oracle(built):
    """The product library must not depend on, or contain, the oracle."""
    import orbx
    out = subprocess.check_output(["ldd", orbx.LIB_PATH]).decode()
    assert "orb_oracle" not in out and "opencv" not in out.lower()
    syms = subprocess.check_output(["nm", "-D", "--defined-only", orbx.LIB_PATH]).decode()
    assert "orc_" not in syms
    csrc = os.path.join(ROOT, "dynamic-visual-slam_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".h")):
            code = re.sub(r"//[^\n]*|/\*.*?\*/", "", open(os.path.join(csrc, f)).read(), flags=re.S)
            assert "oracle" not in code.lower() and "orc_" not in code, f
            if f == "orbx_comm.cu":                      # the one run-time loaded library is NCCL (the path's single collective)
                assert set(re.findall(r'"([\w.]+\.so[\w.]*)"', code)) == {"libnccl.so.2", "libnccl.so"}, f
            else:
                assert "dlopen" not in code, f


def test_option_numbers_of_the_binding_match_the_header():
    """every ORBX_OPT_* of include/orbx.h has its own number, and the ctypes mirror's set_* helpers pass exactly those numbers"""
    hdr = open(os.path.join(ROOT, "include", "orbx.h")).read()
    opts = {name: int(num) for name, num in re.findall(r"#define\s+ORBX_OPT_(\w+)\s+(\d+)", hdr)}
    assert len(set(opts.values())) == len(opts) >= 8 and opts["FILTER_FIRST"] == 8
    src = open(os.path.join(ROOT, "dynamic-visual-slam_b200", "python", "orbx", "__init__.py")).read()
    used = {}
    for meth, doc_opt, num in re.findall(r"def (set_\w+)\(self[^\n]*\n\s+\"\"\"ORBX_OPT_(\w+):.*?orbx_set_option\(self\._h, (\d+),", src, flags=re.S):
        used[doc_opt] = int(num)
    assert used and all(opts[k] == v for k, v in used.items()), (opts, used)
    assert set(used) == set(opts), (sorted(set(opts) - set(used)), "options without a helper in the binding")
