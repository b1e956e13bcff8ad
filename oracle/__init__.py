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

Parity pinning status
---------------------
The reference ships **no tests, golden vectors or fixtures** for this path
(SURVEY.md §4, §8c), so parity cannot be pinned on reference-held vectors.  It is
pinned instead on outputs of the reference's own dependency run here: `cv2` 4.13.0
with `cv2.setUseOptimized(False)` (OpenCV's portable baseline code path — the
IPP/AVX2-dispatched path differs in float rounding of the Shi-Tomasi map by <=1e-8
and is not bit-stable across CPUs).  Golden fixtures generated from that run are
committed under `tests/golden/` with the generating script `oracle/make_golden.py`.
"""
