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


# ---- fisheye stereo rig (Kannala-Brandt), for the stereo post-filter workload
KB8_CAM = (190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504,
           0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182)   # TUM-VI-like 512x512


def kb8_project_f64(cam, X):
    """KannalaBrandt8::project in float64 (numpy restatement, used to synthesise observations)."""
    X = np.asarray(X, np.float64)
    r_xy = np.hypot(X[..., 0], X[..., 1])
    theta = np.arctan2(r_xy, X[..., 2]); psi = np.arctan2(X[..., 1], X[..., 0])
    r = theta + cam[4] * theta ** 3 + cam[5] * theta ** 5 + cam[6] * theta ** 7 + cam[7] * theta ** 9
    return np.stack([cam[0] * r * np.cos(psi) + cam[2], cam[1] * r * np.sin(psi) + cam[3]], axis=-1)


def synth_fisheye_pairs(n, seed, baseline=0.101, noise_px=0.3, outlier_frac=0.15, far_frac=0.1):
    """n keypoint pairs of a two-camera fisheye rig: 3-D points in front of camera 1 (depth 0.3-4 m, `far_frac` of them at
    60-400 m so that the parallax gate fires), projected into both cameras with `noise_px` pixel noise; `outlier_frac` of the
    second observations are replaced by random pixels (reprojection / cheirality gates fire).
    Returns dict(cam1, cam2, R12, t12, xy1[n,2], xy2[n,2], X[n,3], octave1[n], octave2[n])."""
    rng = np.random.default_rng(seed)
    cam1 = np.array(KB8_CAM, np.float64); cam2 = cam1 * (1 + 1e-3 * rng.standard_normal(8))
    ang = np.deg2rad(rng.uniform(-1.5, 1.5, 3))
    cx, cy, cz = np.cos(ang); sx, sy, sz = np.sin(ang)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    R12 = Rz @ Ry @ Rx                                        # camera-2 coordinates -> camera-1 coordinates (Frame::mRlr)
    t12 = np.array([baseline, 0.002, -0.001]) + 1e-3 * rng.standard_normal(3)
    depth = rng.uniform(0.3, 4.0, n)
    far = rng.random(n) < far_frac
    depth[far] = rng.uniform(60.0, 400.0, int(far.sum()))
    theta = rng.uniform(0.0, 1.2, n); psi = rng.uniform(-np.pi, np.pi, n)
    ray = np.stack([np.sin(theta) * np.cos(psi), np.sin(theta) * np.sin(psi), np.cos(theta)], axis=1)
    X1 = ray * (depth / ray[:, 2])[:, None]
    X2 = (X1 - t12) @ R12                                      # R12^T (X1 - t12)
    xy1 = kb8_project_f64(cam1, X1) + noise_px * rng.standard_normal((n, 2))
    xy2 = kb8_project_f64(cam2, X2) + noise_px * rng.standard_normal((n, 2))
    out = rng.random(n) < outlier_frac
    xy2[out] = rng.uniform(20, 490, (int(out.sum()), 2))
    return dict(cam1=cam1.astype(np.float32), cam2=cam2.astype(np.float32), R12=R12.astype(np.float32), t12=t12.astype(np.float32),
                xy1=xy1.astype(np.float32), xy2=xy2.astype(np.float32), X=X1, octave1=rng.integers(0, 8, n).astype(np.int32),
                octave2=rng.integers(0, 8, n).astype(np.int32))
