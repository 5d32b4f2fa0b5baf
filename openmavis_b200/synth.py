"""Deterministic synthetic inputs: textured camera images and descriptor sets.

Pure numpy so the same bytes come out in the build container and on the GPU box.
"""
import numpy as np


def _upsample_linear(grid, h, w):
    gh, gw = grid.shape
    ys = np.linspace(0, gh - 1, h)
    xs = np.linspace(0, gw - 1, w)
    y0 = np.floor(ys).astype(np.int64); y1 = np.minimum(y0 + 1, gh - 1); fy = (ys - y0)[:, None]
    x0 = np.floor(xs).astype(np.int64); x1 = np.minimum(x0 + 1, gw - 1); fx = (xs - x0)[None, :]
    a = grid[y0][:, x0] * (1 - fx) + grid[y0][:, x1] * fx
    b = grid[y1][:, x0] * (1 - fx) + grid[y1][:, x1] * fx
    return a * (1 - fy) + b * fy


def synth_image(w, h, seed, flat_region=True):
    """Textured u8 image: 4 octaves of value noise (period 64/16/4/1 px, amplitude 40/30/18/6),
    120 random rectangles (side 4-40 px, offset +-60) and one flattened region (top-left 25% x 20%
    scaled x0.05) so that the minThFAST retry and the empty-cell paths fire."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float64)
    for period, amp in ((64, 40.0), (16, 30.0), (4, 18.0), (1, 6.0)):
        gh, gw = h // period + 2, w // period + 2
        g = rng.standard_normal((gh, gw))
        img += amp * (_upsample_linear(g, h, w) if period > 1 else g[:h, :w])
    for _ in range(120):
        rw, rh = rng.integers(4, 41, size=2)
        x0 = int(rng.integers(0, max(1, w - rw))); y0 = int(rng.integers(0, max(1, h - rh)))
        img[y0:y0 + rh, x0:x0 + rw] += rng.uniform(-60, 60)
    if flat_region:
        fh, fw = int(h * 0.20), int(w * 0.25)
        img[:fh, :fw] *= 0.05
    return np.clip(np.rint(img + 128.0), 0, 255).astype(np.uint8)


def synth_rig(n_cam, w, h, config, frame):
    """One rig frame: n_cam images, seeds 1000*config + 10*frame + cam."""
    return np.stack([synth_image(w, h, 1000 * config + 10 * frame + c) for c in range(n_cam)])


def synth_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def synth_query_from_train(train, seed, p_flip=0.08, frac_match=0.8):
    """Queries for the brute-force 2-NN workload (config 5): query i = a random train row with
    Binomial(256, p_flip) bit flips for frac_match of the rows, fresh random rows otherwise."""
    rng = np.random.default_rng(seed)
    n = train.shape[0]
    src = rng.integers(0, n, size=n)
    flips = (rng.random((n, 256)) < p_flip)
    flip_bytes = np.packbits(flips, axis=1, bitorder='little')
    q = train[src] ^ flip_bytes
    fresh = rng.random(n) >= frac_match
    q[fresh] = rng.integers(0, 256, size=(int(fresh.sum()), 32), dtype=np.uint8)
    return np.ascontiguousarray(q)
