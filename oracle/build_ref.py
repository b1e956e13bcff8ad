"""Recipe for oracle/_ref/ — the reference's OWN translation units, compiled where they lie.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            # builds oracle/_ref/libvideostab_ref.so (and the other _ref libraries)

What is compiled: `/root/reference/src/Stabilizer.cpp` (all 2688 lines, unmodified, read in place — nothing is
copied into this repository), plus `oracle/ref_shim.cpp` (C entry points).  What it is compiled against:
`oracle/mini_cv/` — a stand-in for the OpenCV *headers* (this container has no OpenCV C++ development files,
SURVEY.md §8c) whose containers are implemented locally and whose image operations forward to callbacks that
`oracle/ref_lib.py` implements with the real OpenCV 4.13 of the cv2 wheel.  None of the reference's
`HAVE_OPENCV_CUDA*` macros is defined, so the TU is the CPU branch (`useCuda=false`) — the parity target.

`-ffp-contract=off` and no `-march`: the host float arithmetic is plain IEEE single/double in source order, which is
also what the GPU kernels are specified to (`-fmad=false`).

Outputs go to `oracle/_ref/` only (git-ignored; it still travels to the GPU box with the snapshot, where
`/root/reference` does not exist — tests and the bench there use the prebuilt library).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VSTAB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

TARGETS = {
    # output                     reference TUs (read in place)         shim
    "libvideostab_ref.so": (["src/Stabilizer.cpp"], "ref_shim.cpp"),
    # the shim #includes src/RollCorrection.cpp and src/AutoZoomCrop.cpp (one TU, so their file-scope statics can be read back)
    "libstages_ref.so": ([], "ref_stages_shim.cpp"),
}

CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w"]


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "Stabilizer.cpp"))


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> list[str]:
    """Build every _ref library whose sources exist.  Returns the list of libraries present afterwards.
    Without /root/reference (the GPU box) nothing is compiled and the prebuilt files are used as they are."""
    os.makedirs(OUT, exist_ok=True)
    built = []
    hdrs = [os.path.join(HERE, "mini_cv", "opencv2", f) for f in os.listdir(os.path.join(HERE, "mini_cv", "opencv2"))]
    for name, (tus, shim) in TARGETS.items():
        out = os.path.join(OUT, name)
        shim_path = os.path.join(HERE, shim)
        srcs = [os.path.join(REF, t) for t in tus]
        if reference_present() and os.path.exists(shim_path) and all(os.path.exists(s) for s in srcs):
            deps = srcs + [shim_path] + hdrs + [os.path.join(REF, "include", "video", "Stabilizer.h"),
                                                os.path.join(REF, "src", "RollCorrection.cpp"), os.path.join(REF, "src", "AutoZoomCrop.cpp")]
            if force or _stale(out, deps):
                cmd = ["g++"] + CXXFLAGS + ["-I", os.path.join(HERE, "mini_cv"), "-I", os.path.join(REF, "include"),
                                            "-I", os.path.join(REF, "src"), "-o", out] + srcs + [shim_path]
                res = subprocess.run(cmd, capture_output=True, text=True)
                if verbose or res.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError(f"building {name} failed")
        if os.path.exists(out):
            built.append(out)
    return built


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv, verbose=True)))
