"""The header-only C++ `vs::Stabilizer` shim (include/video/Stabilizer.h) compiles against a minimal
cv::Mat stand-in, links against the C-ABI library, fails loudly without a GPU, and on a GPU produces
exactly the frames the Python mirror produces."""
import os
import subprocess
import zlib

import numpy as np

import synthclip
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_exe(tmp_path_factory):
    import __graft_entry__
    __graft_entry__.build()
    d = tmp_path_factory.mktemp("shim")
    exe = str(d / "shim_main")
    pkg = os.path.join(ROOT, "video-stab_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "fake_opencv"), os.path.join(ROOT, "tests", "cpp_shim_main.cpp"),
           "-L", pkg, "-lvstab_b200", "-Wl,-rpath," + pkg, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(shim_exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    clip = np.zeros((2, 36, 64, 3), np.uint8)
    p = tmp_path / "clip.raw"
    clip.tofile(p)
    r = subprocess.run([shim_exe, str(p), "64", "36", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_shim_matches_python_mirror(shim_exe, tmp_path):
    import video_stab_b200 as vsb
    w, h, n = 640, 360, 12
    clip = synthclip.make_clip(w, h, n, 31)
    p = tmp_path / "clip.raw"
    clip.tofile(p)
    r = subprocess.run([shim_exe, str(p), str(w), str(h), str(n), "5"], capture_output=True, text=True, check=True)
    lines = [ln.split() for ln in r.stdout.strip().splitlines()]
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5))
    outs = [o for o in (st.stabilize(f) for f in clip) if o is not None]
    while True:
        o = st.flush()
        if o is None:
            break
        outs.append(o)
    assert len(lines) == len(outs) == n
    for ln, o in zip(lines, outs):
        assert int(ln[0], 16) == (zlib.crc32(o.tobytes()) & 0xFFFFFFFF)
        assert (int(ln[1]), int(ln[2])) == (o.shape[1], o.shape[0])
