"""oracle/ — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.

CPU restatement of the per-frame stabilization hot path of OmerMersin/video-stab
(`vs::Stabilizer::stabilize`, reference `src/Stabilizer.cpp`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from here, and only as the checker or the reported CPU
baseline — never as the thing shipped.  The product path (`video-stab_b200/`)
never imports this package and fails loudly when its CUDA library is missing.

Where the arithmetic lives
--------------------------
The reference is a thin host layer over **OpenCV**, which is un-vendored and not
version-pinned (`CMakeLists.txt:11` `find_package(OpenCV REQUIRED)`).  The oracle
therefore has two layers:

* `oracle.stabilizer_ref` — a line-by-line Python restatement of the reference's
  CPU branch (`useCuda=false`) host logic, calling the real OpenCV through the
  Python `cv2` **4.13.0** wheel for the seven library operations on the path.
* `oracle.cv_models` — plain numpy restatements of those seven OpenCV operations
  (gray, resize, pyrDown, Scharr/LK, min-eigenvalue/GFTT, RANSAC partial affine,
  warpAffine), each written from OpenCV's published algorithm.  These are what the
  CUDA kernels are specified against, and they are *pinned* against `cv2` itself
  by `tests/test_oracle_models.py` (bit-exact on every op).

* `oracle/_ref/libvideostab_ref.so` — **the reference itself**: `/root/reference/src/Stabilizer.cpp`,
  all 2688 lines, unmodified and read in place, compiled by `oracle/build_ref.py` against
  `oracle/mini_cv/` (a stand-in for the OpenCV *headers*; its image operations call back into the
  real OpenCV of the cv2 wheel, `oracle/ref_lib.py`).  `oracle.ref_lib.RefStabilizer` drives it.

Parity pinning status: PINNED ON THE REFERENCE'S OWN CODE
---------------------------------------------------------
The reference ships **no tests, golden vectors or fixtures** for this path, so
there are no reference-held vectors to pin on.  The pin is instead the strongest one available:
outputs of the reference itself, run here.  `tests/test_ref_pin.py` runs the compiled reference
(`oracle/_ref`) and the Python restatement on the same seeded clips — 19 configurations covering every
live flag (box / gaussian / kalman, horizon lock, the five border modes, crop-n-zoom, fade, drone mode,
adaptive smoothing, custom corner parameters, 720p / 1080p / odd sizes), degenerate inputs, and clips
built to reach every motion intent — and requires bit equality of every transform, path sample,
corner list, LK status, RANSAC mask, smoothed sample, adaptive radius, motion intent, warp matrix and
output pixel.  The pure host functions are also driven on their own with random inputs, and the
committed goldens (`tests/golden/*.npz`, generator `oracle/make_golden.py`) are re-derived from the
compiled reference and must match field for field.

OpenCV itself is taken from the cv2 4.13.0 wheel with `cv2.setUseOptimized(False)` (OpenCV's portable
baseline code path — the IPP/AVX2-dispatched path differs in float rounding of the Shi-Tomasi map by
<=1e-8 and is not bit-stable across CPUs).  Residual risk: the OpenCV version on the author's Jetson is
unknown (the reference does not pin one).
"""


def reference_available() -> bool:
    """True when the compiled reference (oracle/_ref) can be used on this machine."""
    from . import ref_lib
    return ref_lib.available()


def oracle_kind() -> str:
    return "reference" if reference_available() else "port"


def gaussian_reads_out_of_bounds(params) -> bool:
    """Quirk B-Q8: `gaussianFilterConvolve` pads with `path[center - i]` and
    `path[size - 1 - i]` without checking `size > center` (Stabilizer.cpp:1392-1401).  The first frame is popped when
    the path holds gate-1 samples, so with `clamp(smoothingRadius,5,35) - 1 <= center` the reference reads outside
    the vector — undefined behaviour, whatever the heap holds (observed: 164 LSB of garbage on the first outputs).
    The drop-in DEFINES that case (box filter until the path is longer than `center`), which is what the Python
    restatement implements; the compiled reference cannot be an oracle there."""
    import math
    import numpy as np
    if getattr(params, "smoothingMethod", "box") != "gaussian":
        return False
    if getattr(params, "adaptiveSmoothing", False):
        return True                                   # the gate moves with the data: be conservative
    sigma = np.float32(params.gaussianSigma)
    k = max(3, int(math.ceil(np.float32(6) * sigma)))
    k += k % 2 == 0
    gate = max(5, min(int(params.smoothingRadius), 35))
    return gate - 1 <= k // 2


def run_clip(frames, params, flush: bool = True, use_optimized: bool = False):
    """Push every frame, then flush — through the compiled reference when oracle/_ref is present (it travels to
    the GPU box with the snapshot), else through the Python restatement (bit-identical, tests/test_ref_pin.py).
    Returns (outputs, stabilizer) with `.frame_records`, `.output_records`, `.first_corners` on either."""
    if reference_available() and not gaussian_reads_out_of_bounds(params):
        from . import ref_lib
        return ref_lib.run_clip(frames, params, flush=flush, use_optimized=use_optimized)
    from . import stabilizer_ref
    return stabilizer_ref.run_clip(frames, params, flush=flush, use_optimized=use_optimized)
