"""Seeded synthetic shaky-clip generator shared by the tests, the bench and the oracle
(SURVEY.md §8d).  Data tooling only — not on the stabilization path.

Base texture: uniform uint8 noise (seed S), (H+2m)x(W+2m)x3, Gaussian-blurred sigma=2
and min-max normalised to 0..255 (gives several hundred well-spread Shi-Tomasi corners).
Frame k is the base seen through a ground-truth camera pose: slow sinusoidal pan
(<= ~1 px/frame) + per-frame jitter N(0, 3^2) px and N(0, 0.004^2) rad (clipped so the
view never leaves the margin m=64), RNG numpy.default_rng(S+1).
"""
from __future__ import annotations

import numpy as np

MARGIN = 64


def base_texture(width: int, height: int, seed: int) -> np.ndarray:
    import cv2  # data tooling only
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (height + 2 * MARGIN, width + 2 * MARGIN, 3), dtype=np.uint8)
    blur = cv2.GaussianBlur(noise, (0, 0), 2.0).astype(np.float32)
    lo, hi = blur.min(), blur.max()
    return ((blur - lo) / (hi - lo) * 255.0).astype(np.uint8)


def camera_path(n_frames: int, seed: int) -> np.ndarray:
    """(n,3) float64 ground-truth pose (x, y, angle) per frame."""
    rng = np.random.default_rng(seed + 1)
    k = np.arange(n_frames, dtype=np.float64)
    px = 40.0 * np.sin(2 * np.pi * k / 240.0)
    py = 20.0 * np.sin(2 * np.pi * k / 180.0 + 1.0)
    jx = np.clip(rng.normal(0.0, 3.0, n_frames), -10, 10)
    jy = np.clip(rng.normal(0.0, 3.0, n_frames), -10, 10)
    ja = np.clip(rng.normal(0.0, 0.004, n_frames), -0.012, 0.012)
    return np.stack([px + jx, py + jy, ja], axis=1)


def render_frame(base: np.ndarray, pose, width: int, height: int) -> np.ndarray:
    import cv2  # data tooling only
    x, y, a = (float(v) for v in pose)
    cx = base.shape[1] / 2.0
    cy = base.shape[0] / 2.0
    c, s = np.cos(a), np.sin(a)
    # rotate about the base centre, translate, then shift so the output is the centre crop
    m = np.array([[c, -s, cx - c * cx + s * cy + x - MARGIN],
                  [s, c, cy - s * cx - c * cy + y - MARGIN]], np.float64)
    return cv2.warpAffine(base, m, (width, height), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT)


def make_clip(width: int, height: int, n_frames: int, seed: int) -> np.ndarray:
    """(n, H, W, 3) uint8 BGR clip."""
    base = base_texture(width, height, seed)
    poses = camera_path(n_frames, seed)
    out = np.empty((n_frames, height, width, 3), np.uint8)
    for k in range(n_frames):
        out[k] = render_frame(base, poses[k], width, height)
    return out


def shake_clip(width: int, height: int, n_frames: int, seed: int, theta: float = 0.008, jump: float = 8.0, every: int = 7) -> np.ndarray:
    """A clip built to reach the reference's SHAKE_REMOVAL / FOLLOW_ACTION motion intents (Stabilizer.cpp:1676-1719):
    an alternating roll of +-theta about the frame ORIGIN (so the translation part of the fitted similarity stays
    near zero while |da| is large) with a translation jump every `every` frames (erratic magnitudes: low consistency)."""
    import cv2  # data tooling only
    base = base_texture(width + 4 * MARGIN, height, seed)
    out = np.empty((n_frames, height, width, 3), np.uint8)
    for k in range(n_frames):
        th = theta * (1.0 if k % 2 == 0 else -1.0)
        offx = MARGIN + jump * (k // every)
        offy = float(MARGIN)
        c, s = np.cos(th), np.sin(th)
        m = np.array([[c, -s, -(c * offx - s * offy)], [s, c, -(s * offx + c * offy)]], np.float64)
        out[k] = cv2.warpAffine(base, m, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    return out


def pan_clip(width: int, height: int, n_frames: int, seed: int, speed: float = 14.0) -> np.ndarray:
    """A steady horizontal pan (constant direction, consistent magnitude): reaches DELIBERATE_PAN."""
    base = base_texture(width + int(speed * n_frames) + 2 * MARGIN, height, seed)
    cx = (base.shape[1] - width) / 2.0 - MARGIN
    return np.stack([render_frame(base, (-cx + speed * k, 0.0, 0.0), width, height) for k in range(n_frames)])


class DeviceClip:
    """An arbitrarily long clip generated ON THE DEVICE from a seed (a 10-minute 1080p clip is 112 GB raw, so it cannot
    be built on the host): integer-offset crops of five pre-rotated copies of one seeded texture, following a slow
    sinusoidal pan plus seeded per-frame jitter (N(0, 3^2) px, rotation from {0, +-0.002, +-0.004} rad).  Every rank of a
    multi-GPU run builds the same clip from the same seed and materialises only its own frames.  Needs torch + CUDA."""

    PAD = 96

    def __init__(self, width: int, height: int, n_frames: int, seed: int, device):
        import cv2  # data tooling only
        import torch
        self.torch, self.w, self.h, self.n, self.device = torch, width, height, n_frames, device
        m = self.PAD
        base = base_texture(width + 2 * m - 2 * MARGIN, height + 2 * m - 2 * MARGIN, seed)
        cx, cy = base.shape[1] / 2.0, base.shape[0] / 2.0
        self.rots = []
        for a in (-0.004, -0.002, 0.0, 0.002, 0.004):
            c, s = np.cos(a), np.sin(a)
            mat = np.array([[c, -s, cx - c * cx + s * cy], [s, c, cy - s * cx - c * cy]])
            self.rots.append(torch.from_numpy(cv2.warpAffine(base, mat, (base.shape[1], base.shape[0]))).to(device))
        rng = np.random.default_rng(seed + 1)
        k = np.arange(n_frames)
        self.xs = np.clip(np.rint(m + 40 * np.sin(2 * np.pi * k / 240.0) + rng.normal(0, 3, n_frames)), 0, 2 * m).astype(int)
        self.ys = np.clip(np.rint(m + 20 * np.sin(2 * np.pi * k / 180.0 + 1.0) + rng.normal(0, 3, n_frames)), 0, 2 * m).astype(int)
        self.rs = rng.integers(0, len(self.rots), n_frames)

    def frames(self, a: int, b: int, out=None):
        """frames [a, b) as a contiguous (b-a, H, W, 3) uint8 tensor on the device"""
        if out is None:
            out = self.torch.empty((b - a, self.h, self.w, 3), dtype=self.torch.uint8, device=self.device)
        for i in range(a, b):
            out[i - a] = self.rots[self.rs[i]][self.ys[i]:self.ys[i] + self.h, self.xs[i]:self.xs[i] + self.w]
        return out


def horizon_clip(width: int, height: int, n_frames: int, seed: int, roll_deg=None) -> np.ndarray:
    """Frames with a strong straight horizon (bright sky over textured ground plus a few long straight structures) seen
    through a slowly varying camera ROLL: input for RollCorrection (Canny + Hough find the horizon lines)."""
    import cv2  # data tooling only
    rng = np.random.default_rng(seed)
    big_w, big_h = width + 4 * MARGIN, height + 4 * MARGIN
    tex = base_texture(big_w - 2 * MARGIN, big_h - 2 * MARGIN, seed)
    scene = (tex // 4 + 40).astype(np.uint8)
    hy = big_h // 2
    scene[:hy] = (scene[:hy] // 2 + 150).astype(np.uint8)                      # sky
    cv2.line(scene, (0, hy), (big_w - 1, hy), (250, 250, 250), 3)
    for k in range(3):
        y = hy + 60 + 70 * k
        cv2.line(scene, (0, y), (big_w - 1, y), (int(rng.integers(0, 80)),) * 3, 2)
    if roll_deg is None:
        k = np.arange(n_frames)
        roll_deg = 3.0 * np.sin(2 * np.pi * k / 40.0) + rng.normal(0, 0.2, n_frames)
    out = np.empty((n_frames, height, width, 3), np.uint8)
    cx, cy = big_w / 2.0, big_h / 2.0
    for i in range(n_frames):
        M = cv2.getRotationMatrix2D((cx, cy), float(roll_deg[i]), 1.0)
        M[0, 2] -= (big_w - width) / 2.0
        M[1, 2] -= (big_h - height) / 2.0
        out[i] = cv2.warpAffine(scene, M, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    return out


def black_corner_frame(width: int, height: int, seed: int, angle_deg: float, shift=(0.0, 0.0), hole: bool = False) -> np.ndarray:
    """A frame as a roll correction / stabilizer with BORDER_CONSTANT leaves it: never-black content rotated about the centre and
    shifted, black where nothing maps (input for AutoZoomCrop).  `hole` paints a black object inside the content."""
    import cv2  # data tooling only
    img = np.maximum(base_texture(width, height, seed)[MARGIN:-MARGIN, MARGIN:-MARGIN], 8)
    M = cv2.getRotationMatrix2D((width / 2.0, height / 2.0), float(angle_deg), 1.0)
    M[0, 2] += shift[0]
    M[1, 2] += shift[1]
    fr = cv2.warpAffine(img, M, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    if hole:
        fr[height // 3: height // 3 + 40, width // 2: width // 2 + 60] = 0
    return fr
