"""Seeded synthetic frames: value-noise background translating by
(1.5, -0.75) px/frame + 6 textured discs on fixed trajectories.  Pure numpy, deterministic,
shared by the tests, bench.py and the golden generator.  No arithmetic of the hot path lives here: it only
manufactures input frames (there is no network for datasets), like seeded_weights.py manufactures weights."""
import numpy as np


def _value_noise(h, w, ox, oy, rng_seed):
    out = np.zeros((h, w), dtype=np.float64)
    amp = 1.0
    tot = 0.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    for octave in range(5):
        cell = 128 >> octave
        gx = (xx + ox) / cell
        gy = (yy + oy) / cell
        x0 = np.floor(gx).astype(np.int64)
        y0 = np.floor(gy).astype(np.int64)
        fx = gx - x0
        fy = gy - y0
        fx = fx * fx * (3 - 2 * fx)
        fy = fy * fy * (3 - 2 * fy)

        def hsh(ix, iy):
            n = (ix * 374761393 + iy * 668265263 + (rng_seed + octave) * 2147483647) & 0xFFFFFFFF
            n = ((n ^ (n >> 13)) * 1274126177) & 0xFFFFFFFF
            return ((n ^ (n >> 16)) & 0xFFFF) / 65535.0

        v = (hsh(x0, y0) * (1 - fx) + hsh(x0 + 1, y0) * fx) * (1 - fy) + \
            (hsh(x0, y0 + 1) * (1 - fx) + hsh(x0 + 1, y0 + 1) * fx) * fy
        out += amp * v
        tot += amp
        amp *= 0.5
    return out / tot


def synthetic_frame(h, w, t, seed=0):
    """Frame t of the seeded synthetic clip: HxWx3 u8 RGB."""
    ox, oy = 1.5 * t, -0.75 * t
    img = np.stack([_value_noise(h, w, ox + 1000 * c, oy + 777 * c, seed + 11 * c) for c in range(3)], axis=-1)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    rng = np.random.RandomState(seed + 12345)
    for d in range(6):
        cx0, cy0 = rng.uniform(0.15, 0.85) * w, rng.uniform(0.15, 0.85) * h
        vx, vy = rng.uniform(-4, 4), rng.uniform(-3, 3)
        rad = rng.uniform(0.04, 0.10) * min(h, w)
        col = rng.uniform(0.1, 0.9, size=3)
        cx = (cx0 + vx * t) % w
        cy = (cy0 + vy * t) % h
        r2 = (xx - cx) ** 2 + (yy - cy) ** 2
        m = r2 < rad * rad
        tex = 0.75 + 0.25 * np.sin((xx - cx) * 0.35 + d) * np.cos((yy - cy) * 0.29 - d)
        for c in range(3):
            ch = img[..., c]
            ch[m] = (col[c] * tex)[m]
    return np.clip(img * 255.0, 0, 255).astype(np.uint8)
