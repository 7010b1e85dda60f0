"""Seeded synthetic chimney-swift video (numpy twin of csrc/synth.cu).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every pixel is a pure function of ``(seed, video, t, y, x)`` through a
stateless 32-bit integer hash, so the CPU oracle and the CUDA generator
(``swb_synth_frames``) produce bit-identical frames without copying video
across PCIe.  The reference ships no sample video
(its .gitignore excludes videos/ *.mp4 *.jpg).

Frame model
* background: static per-channel gradient  ``140 + 10*c + (64*x)//W - (48*y)//H``
* noise: per pixel, per frame, per channel in {-2..+2} from the hash
* birds: ``n_birds`` dark rectangles/ellipses (value ``40 + 5*c`` + noise),
  5..12 px wide, 7..16 px tall, moving 3..8 px/frame horizontally and up to
  4 px/frame vertically (16ths-of-a-pixel fixed point, never axis aligned),
  wrapping around the frame edges so the count stays constant.
"""

import numpy as np

M32 = 0xFFFFFFFF
GOLD = 0x9E3779B1
K_T = 0x85EBCA6B
BIRD_TAG = 0xB1D50000


def mix32(x):
    """lowbias32 integer finaliser on a Python int."""
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def mix32_np(x):
    """Same finaliser on a uint32 ndarray (wrap-around arithmetic)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def video_key(seed, video):
    return mix32((seed * GOLD + video) & M32)


def frame_key(seed, video, t):
    return mix32(video_key(seed, video) ^ ((t * K_T) & M32))


def bird_params(seed, video, b, width, height):
    """(x0, y0, vx16, vy16, bw, bh, ellipse) for bird ``b``."""
    vk = video_key(seed, video)
    hk = [mix32(vk ^ ((BIRD_TAG + b * 8 + k) & M32)) for k in range(7)]
    x0 = hk[0] % width
    y0 = hk[1] % height
    vx16 = 48 + hk[2] % 80
    if hk[2] & 0x80000000:
        vx16 = -vx16
    vy16 = ((hk[3] % 129) - 64) | 1
    bw = 5 + hk[4] % 8
    bh = 7 + hk[5] % 10
    ellipse = hk[6] & 1
    return x0, y0, vx16, vy16, bw, bh, ellipse


def bird_origin(x0, y0, vx16, vy16, t, width, height):
    cx = ((x0 * 16 + vx16 * t) >> 4) % width
    cy = ((y0 * 16 + vy16 * t) >> 4) % height
    return cx, cy


def bird_shape(bw, bh, ellipse):
    """bool (bh, bw) footprint: full rectangle or inscribed ellipse
    ``(2dx+1-bw)^2 * bh^2 + (2dy+1-bh)^2 * bw^2 <= bw^2 * bh^2``."""
    if not ellipse:
        return np.ones((bh, bw), dtype=bool)
    dx = 2 * np.arange(bw, dtype=np.int64) + 1 - bw
    dy = 2 * np.arange(bh, dtype=np.int64) + 1 - bh
    return (dx[None, :] ** 2 * bh * bh + dy[:, None] ** 2 * bw * bw
            <= bw * bw * bh * bh)


def synth_frame(seed, video, t, height, width, n_birds):
    """One (H, W, 3) uint8 BGR frame."""
    fk = frame_key(seed, video, t)
    idx = np.arange(height * width, dtype=np.uint32)
    h = mix32_np(np.uint32(fk) + idx * np.uint32(GOLD)).reshape(height, width)
    noise = np.empty((height, width, 3), dtype=np.int32)
    for c in range(3):
        noise[..., c] = ((((h >> np.uint32(8 * c)) & np.uint32(0xFF))
                          * np.uint32(5)) >> np.uint32(8)).astype(np.int32) - 2
    xs = np.arange(width, dtype=np.int64)
    ys = np.arange(height, dtype=np.int64)
    grad = ((64 * xs) // width)[None, :] - ((48 * ys) // height)[:, None]
    frame = np.empty((height, width, 3), dtype=np.int32)
    for c in range(3):
        frame[..., c] = 140 + 10 * c + grad + noise[..., c]
    bird = np.zeros((height, width), dtype=bool)
    for b in range(n_birds):
        x0, y0, vx16, vy16, bw, bh, ell = bird_params(seed, video, b, width, height)
        cx, cy = bird_origin(x0, y0, vx16, vy16, t, width, height)
        rows = (cy + np.arange(bh)) % height
        cols = (cx + np.arange(bw)) % width
        sub = bird[np.ix_(rows, cols)]
        bird[np.ix_(rows, cols)] = sub | bird_shape(bw, bh, ell)
    for c in range(3):
        frame[..., c] = np.where(bird, 40 + 5 * c + noise[..., c], frame[..., c])
    return frame.astype(np.uint8)


def synth_video(seed, video, t0, n_frames, height, width, n_birds):
    """(n_frames, H, W, 3) uint8, frames t0 .. t0+n_frames-1."""
    return np.stack([synth_frame(seed, video, t0 + i, height, width, n_birds)
                     for i in range(n_frames)])
