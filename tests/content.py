"""Deterministic synthetic I420 content (SURVEY.md 8(d) "Inputs").

The reference's own clip (sequence/foreman.zip) is not in the mount, so every
config uses seeded synthetic content:

  * ``panning``  : box-blurred noise texture panned (3,2) px/frame + a moving
                   high-contrast rectangle + +-2 uniform noise; chroma = subsampled
                   offsets of the same texture.  Exercises skip / inter / intra.
  * ``chessboard``: the reference CLI's ``--gen`` rotating chessboard (T:407-452),
                   luma only, chroma = 128.

Only numpy integer / float64 ops with fixed seeds are used, so the frames are
bit-identical on every machine (MD5s of the small golden clips are committed in
tests/golden/manifest.json).
"""
import numpy as np


def _box_blur(a, r):
    """Separable box blur with wrap-around, integer arithmetic (deterministic)."""
    a = a.astype(np.int64)
    for axis in (0, 1):
        acc = np.zeros_like(a)
        for d in range(-r, r + 1):
            acc += np.roll(a, d, axis=axis)
        a = acc // (2 * r + 1)
    return a


def make_texture(seed, th, tw):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(th, tw), dtype=np.int64)
    fine = _box_blur(base, 1)
    coarse = _box_blur(rng.integers(0, 256, size=(th, tw), dtype=np.int64), 6)
    tex = (fine + 3 * coarse) // 4
    # stretch contrast a little
    tex = np.clip((tex - 128) * 2 + 128, 0, 255)
    return tex.astype(np.uint8)


def panning(width, height, nframes, seed=1234, noise=2):
    """Returns uint8 array [nframes, width*height*3//2] (planar I420)."""
    th, tw = height + 4 * nframes + 64, width + 4 * nframes + 64
    th += th & 1
    tw += tw & 1
    tex = make_texture(seed, th, tw)
    rng = np.random.default_rng(seed + 1)
    out = np.empty((nframes, width * height * 3 // 2), dtype=np.uint8)
    cw, ch = width // 2, height // 2
    for f in range(nframes):
        ox, oy = 3 * f, 2 * f
        y = tex[oy:oy + height, ox:ox + width].astype(np.int16)
        # moving high-contrast rectangle
        rw, rh = max(16, width // 6), max(16, height // 6)
        rx = (7 * f + width // 5) % max(1, width - rw)
        ry = (5 * f + height // 4) % max(1, height - rh)
        y[ry:ry + rh, rx:rx + rw] = 235 if (f // 8) % 2 == 0 else 20
        if noise:
            y = y + rng.integers(-noise, noise + 1, size=y.shape, dtype=np.int16)
        y = np.clip(y, 0, 255).astype(np.uint8)
        u = tex[(oy // 2 + 17):(oy // 2 + 17) + 2 * ch:2, (ox // 2 + 5):(ox // 2 + 5) + 2 * cw:2]
        v = tex[(oy // 2 + 41):(oy // 2 + 41) + 2 * ch:2, (ox // 2 + 29):(ox // 2 + 29) + 2 * cw:2]
        u = (u.astype(np.int16) // 2 + 64).astype(np.uint8)
        v = (v.astype(np.int16) // 2 + 64).astype(np.uint8)
        if rw and rh:
            u = u.copy(); v = v.copy()
            u[ry // 2:(ry + rh) // 2, rx // 2:(rx + rw) // 2] = 90
            v[ry // 2:(ry + rh) // 2, rx // 2:(rx + rw) // 2] = 200
        out[f, :width * height] = y.reshape(-1)
        out[f, width * height:width * height + cw * ch] = u.reshape(-1)
        out[f, width * height + cw * ch:] = v.reshape(-1)
    return out


def noise_frames(width, height, nframes, seed=7):
    """Uniform random frames: stresses transform / CAVLC escapes / intra."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(nframes, width * height * 3 // 2), dtype=np.uint8)


def flat_frames(width, height, nframes, value=128):
    return np.full((nframes, width * height * 3 // 2), value, dtype=np.uint8)


def chessboard(width, height, nframes):
    """The reference CLI's --gen content (T:407-452), vectorised in float64."""
    out = np.empty((nframes, width * height * 3 // 2), dtype=np.uint8)
    hw, hh = width >> 1, height >> 1
    c = np.arange(width, dtype=np.float64)[None, :] - hw
    r = np.arange(height, dtype=np.float64)[:, None] - hh
    for f in range(nframes):
        co, si = np.cos(.01 * f), np.sin(.01 * f)
        x = co * c + si * r
        y = -si * c + co * r
        mid = (np.abs(x) < 4) & (np.abs(y) < 4)
        i = np.trunc(x).astype(np.int64)
        j = np.trunc(y).astype(np.int64)
        # C integer division truncates toward zero
        black = np.where(mid, 128, np.trunc(i / 16).astype(np.int64))
        white = np.where(mid, 128, 255 - np.trunc(j / 16).astype(np.int64))

        def cell(ii, jj):
            return np.where((((ii >> 4) + (jj >> 4)) & 1) != 0, white, black).astype(np.float64)
        c00, c01, c10, c11 = cell(i, j), cell(i + 1, j), cell(i, j + 1), cell(i + 1, j + 1)
        fx, fy = x - i, y - j
        s = np.trunc((c00 * (1 - fx) + c01 * fx) * (1 - fy) + (c10 * (1 - fx) + c11 * fx) * fy + 0.5)
        s = np.clip(s, 0, 255).astype(np.uint8)
        out[f, :width * height] = s.reshape(-1)
        out[f, width * height:] = 128
    return out


def multi_motion(width, height, nframes, seed=77):
    """Textured background panning fast (11,-7) px/frame with three textured objects moving
    with different large velocities: keeps the reference's running MV "clusters" changing
    inside a frame, which exercises the speculate/verify/repair path (csrc/h264_wave.h)."""
    pad = 16 * nframes + 64
    tex = make_texture(seed, height + 2 * pad, width + 2 * pad)
    objs = []
    rng = np.random.default_rng(seed + 5)
    for k in range(3):
        ow, oh = max(32, width // (3 + k)), max(32, height // (4 + k))
        objs.append((make_texture(seed + 10 + k, oh, ow), int(rng.integers(0, max(1, width - ow))),
                     int(rng.integers(0, max(1, height - oh))), int(rng.integers(-14, 15)), int(rng.integers(-9, 10))))
    out = np.empty((nframes, width * height * 3 // 2), dtype=np.uint8)
    cw, ch = width // 2, height // 2
    for f in range(nframes):
        ox, oy = pad + 11 * f, pad - 7 * f
        y = tex[oy:oy + height, ox:ox + width].copy()
        for (ot, x0, y0, vx, vy) in objs:
            oh, ow = ot.shape
            px = (x0 + vx * f) % max(1, width - ow)
            py = (y0 + vy * f) % max(1, height - oh)
            y[py:py + oh, px:px + ow] = ot
        y = np.clip(y.astype(np.int16) + rng.integers(-2, 3, size=y.shape, dtype=np.int16), 0, 255).astype(np.uint8)
        u = (y[0::2, 0::2] // 2 + 64).astype(np.uint8)
        v = (y[1::2, 1::2] // 2 + 32).astype(np.uint8)
        out[f, :width * height] = y.reshape(-1)
        out[f, width * height:width * height + cw * ch] = u.reshape(-1)
        out[f, width * height + cw * ch:] = v.reshape(-1)
    return out
