"""CPU: pins the oracle's fisheye stereo post-filter (oracle/orb_oracle.cpp: kb8_*, null_vector4, orc_stereo_fisheye_matches;
reference src/CameraModels/KannalaBrandt8.cpp:48-126, 319-427, src/Frame.cc:1409-1459).

The reference holds no tests or golden vectors for this function and its SVD lives in Eigen (not in this image): the
oracle's null vector is pinned against numpy.linalg.svd (LAPACK), its camera model and the whole TriangulateMatches against
an independent float64 numpy restatement."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import KP_DTYPE, lib, p
from openmavis_b200.synth import kb8_project_f64, synth_descriptors, synth_fisheye_pairs, synth_query_from_train

RIG = np.dtype([('cam1', 'f4', 8), ('cam2', 'f4', 8), ('R12', 'f4', 9), ('t12', 'f4', 3), ('precision1', 'f4'), ('precision2', 'f4')])


def make_rig(s):
    r = np.zeros(1, RIG)
    r['cam1'][0] = s['cam1']; r['cam2'][0] = s['cam2']; r['R12'][0] = s['R12'].reshape(9); r['t12'][0] = s['t12']
    r['precision1'] = np.float32(1e-6); r['precision2'] = np.float32(1e-6)
    return r


def oracle_triangulate(rig, xy1, xy2, s1, s2):
    n = xy1.shape[0]
    depth = np.zeros(n, np.float32); p3d = np.zeros((n, 3), np.float32); margin = np.zeros(n, np.float32)
    lib().orc_kb8_triangulate_matches(p(rig), p(xy1), p(xy2), p(s1), p(s2), n, p(depth), p(p3d), p(margin))
    return depth, p3d, margin


def np_unproject(cam, xy):
    cam = cam.astype(np.float64); xy = xy.astype(np.float64)
    pw = (xy - cam[2:4]) / cam[0:2]
    td = np.clip(np.hypot(pw[:, 0], pw[:, 1]), -np.pi / 2, np.pi / 2)
    th = td.copy()
    for _ in range(20):
        t2 = th * th
        f = th * (1 + cam[4] * t2 + cam[5] * t2 ** 2 + cam[6] * t2 ** 3 + cam[7] * t2 ** 4) - td
        df = 1 + 3 * cam[4] * t2 + 5 * cam[5] * t2 ** 2 + 7 * cam[6] * t2 ** 3 + 9 * cam[7] * t2 ** 4
        th = th - f / df
    scale = np.where(td > 1e-8, np.tan(th) / np.maximum(td, 1e-300), 1.0)
    return np.concatenate([pw * scale[:, None], np.ones((xy.shape[0], 1))], axis=1)


def np_triangulate(s, xy1, xy2, s1, s2):
    """float64 restatement of TriangulateMatches with numpy's SVD."""
    R = s['R12'].astype(np.float64); t = s['t12'].astype(np.float64)
    r1 = np_unproject(s['cam1'], xy1); r2 = np_unproject(s['cam2'], xy2)
    out = np.zeros(xy1.shape[0]); P = np.zeros((xy1.shape[0], 3))
    Rt = R.T; T2 = np.concatenate([Rt, (-Rt @ t)[:, None]], axis=1)
    for i in range(xy1.shape[0]):
        r21 = R @ r2[i]
        if r1[i] @ r21 / (np.linalg.norm(r1[i]) * np.linalg.norm(r21)) > 0.9998:
            out[i] = -1; continue
        A = np.array([[-1, 0, r1[i, 0], 0], [0, -1, r1[i, 1], 0], r2[i, 0] * T2[2] - T2[0], r2[i, 1] * T2[2] - T2[1]])
        v = np.linalg.svd(A)[2][3]
        x = v[:3] / v[3]
        if x[2] <= 0:
            out[i] = -2; continue
        z2 = T2[2, :3] @ x + T2[2, 3]
        if z2 <= 0:
            out[i] = -3; continue
        e1 = kb8_project_f64(s['cam1'].astype(np.float64), x) - xy1[i]
        if e1 @ e1 > 5.991 * s1[i]:
            out[i] = -4; continue
        e2 = kb8_project_f64(s['cam2'].astype(np.float64), T2[:, :3] @ x + T2[:, 3]) - xy2[i]
        if e2 @ e2 > 5.991 * s2[i]:
            out[i] = -5; continue
        out[i] = x[2]; P[i] = x
    return out, P


def test_null_vector_matches_lapack():
    rng = np.random.default_rng(11)
    L = lib()
    worst = 0.0
    for k in range(400):
        if k % 2:
            a = rng.standard_normal((4, 4)).astype(np.float32)
        else:   # DLT-shaped: two rows of camera 1, two of camera 2, near rank 3
            x = np.append(rng.uniform(-2, 2, 2), [rng.uniform(0.5, 10), 1.0])
            a = rng.standard_normal((4, 4))
            a -= np.outer(a @ x, x) / (x @ x)
            a = (a + 1e-4 * rng.standard_normal((4, 4))).astype(np.float32)
        v = np.zeros(4)
        L.orc_null_vector4(p(a), p(v))
        u, sv, vt = np.linalg.svd(a.astype(np.float64))
        ref = vt[3]
        assert abs(np.linalg.norm(v) - 1) < 1e-12
        err = min(np.linalg.norm(v - ref), np.linalg.norm(v + ref))
        gap = (sv[2] - sv[3]) / sv[0]
        worst = max(worst, err * gap)
        assert err * gap < 1e-10, (k, err, gap)
    assert worst < 1e-10


def test_camera_model_against_float64():
    s = synth_fisheye_pairs(2000, 3)
    L = lib()
    rays = np.zeros((2000, 3), np.float32)
    L.orc_kb8_unproject(p(s['cam1']), C.c_float(1e-6), p(s['xy1']), 2000, p(rays))
    ref = np_unproject(s['cam1'], s['xy1'])
    assert np.allclose(rays, ref, rtol=2e-5, atol=2e-6)
    # project(unproject(p)) == p
    uv = np.zeros((2000, 2), np.float32)
    L.orc_kb8_project(p(s['cam1']), p(rays), 2000, p(uv))
    assert np.abs(uv - s['xy1']).max() < 2e-3
    assert np.allclose(uv, kb8_project_f64(s['cam1'].astype(np.float64), rays.astype(np.float64)), rtol=0, atol=2e-3)
    # the principal point unprojects to the optical axis (theta_d <= 1e-8 branch)
    c = np.array([[s['cam1'][2], s['cam1'][3]]], np.float32)
    r = np.zeros((1, 3), np.float32)
    L.orc_kb8_unproject(p(s['cam1']), C.c_float(1e-6), p(c), 1, p(r))
    assert r.tolist() == [[0.0, 0.0, 1.0]]


def test_triangulate_matches_against_float64_restatement():
    s = synth_fisheye_pairs(3000, 5)
    rig = make_rig(s)
    sig = (np.float32(1.2) ** (2 * s['octave1'])).astype(np.float32); sig2 = (np.float32(1.2) ** (2 * s['octave2'])).astype(np.float32)
    depth, p3d, margin = oracle_triangulate(rig, s['xy1'], s['xy2'], sig, sig2)
    ref, P = np_triangulate(s, s['xy1'], s['xy2'], sig, sig2)
    codes = {c: int((depth == c).sum()) for c in (-1, -2, -3, -4, -5)}
    assert all(v > 0 for v in codes.values()), codes                 # every reject path of the reference is exercised
    ok = depth > 0
    assert ok.sum() > 1200
    clear = margin > 1e-3                                            # away from every gate the decisions must agree
    assert clear.mean() > 0.97
    same = np.where(ok, ref > 0, ref == depth)
    assert same[clear].all()
    both = ok & (ref > 0)
    assert np.allclose(depth[both], ref[both], rtol=2e-3)
    assert np.allclose(p3d[both], P[both], rtol=2e-3, atol=1e-4)
    # accepted points are the true points
    near = both & (s["X"][:, 2] < 4.5)
    rel = np.abs(depth[near] - s['X'][near, 2]) / s['X'][near, 2]
    assert np.median(rel) < 0.06                                   # 0.3 px noise on a 10 cm baseline


def fisheye_frame(seed, n_left=900, n_right=950, mono_left=120, mono_right=140):
    """keypoints + descriptors of a fisheye stereo frame: the lapping blocks hold true correspondences (flipped descriptor
    bits), duplicates (two left keypoints with the same best right one) and unmatched rows."""
    rng = np.random.default_rng(seed)
    nq, nt = n_left - mono_left, n_right - mono_right
    s = synth_fisheye_pairs(nt, seed + 1)
    kr = np.zeros(n_right, KP_DTYPE); kl = np.zeros(n_left, KP_DTYPE)
    dr = synth_descriptors(n_right, seed + 2)
    kr['x'][mono_right:] = s['xy2'][:, 0]; kr['y'][mono_right:] = s['xy2'][:, 1]; kr['octave'][mono_right:] = s['octave2']
    kr['x'][:mono_right] = rng.uniform(0, 512, mono_right); kr['y'][:mono_right] = rng.uniform(0, 512, mono_right)
    src = rng.integers(0, nt, nq)                                   # left row q observes the point of right row src[q] (duplicates happen)
    flips = np.packbits(rng.random((nq, 256)) < 0.06, axis=1, bitorder='little')
    dl = synth_descriptors(n_left, seed + 3)
    dl[mono_left:] = dr[mono_right:][src] ^ flips
    fresh = rng.random(nq) < 0.2
    dl[mono_left:][fresh] = synth_descriptors(int(fresh.sum()), seed + 4)
    kl['x'][mono_left:] = s['xy1'][src, 0] + 0.1 * rng.standard_normal(nq); kl['y'][mono_left:] = s['xy1'][src, 1] + 0.1 * rng.standard_normal(nq)
    kl['octave'][mono_left:] = s['octave1'][src]
    kl['x'][:mono_left] = rng.uniform(0, 512, mono_left); kl['y'][:mono_left] = rng.uniform(0, 512, mono_left)
    sigma2 = np.array([np.float32(1.2) ** (2 * l) for l in range(8)], np.float32)
    return dict(rig=make_rig(s), s=s, kl=kl, dl=dl, kr=kr, dr=dr, mono_left=mono_left, mono_right=mono_right, sigma2=sigma2, src=src)


def oracle_fisheye_matches(f, ratio):
    kl, kr = f['kl'], f['kr']
    l2r = np.zeros(kl.size, np.int32); r2l = np.zeros(kr.size, np.int32)
    depth = np.zeros(kl.size, np.float32); p3d = np.zeros((kl.size, 3), np.float32); margin = np.zeros(kl.size, np.float32)
    dm = C.c_int(0)
    L = lib()
    L.orc_stereo_fisheye_matches.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 6
    n = L.orc_stereo_fisheye_matches(p(f['rig']), p(kl), p(f['dl']), kl.size, f['mono_left'], p(kr), p(f['dr']), kr.size, f['mono_right'],
                                     p(f['sigma2']), ratio, p(l2r), p(r2l), p(depth), p(p3d), p(margin), C.cast(C.byref(dm), C.c_void_p))
    return dict(nMatches=n, descMatches=dm.value, l2r=l2r, r2l=r2l, depth=depth, p3d=p3d, margin=margin)


@pytest.mark.parametrize('ratio', [0.7, 0.8])
def test_stereo_fisheye_matches_bookkeeping(ratio):
    f = fisheye_frame(40)
    o = oracle_fisheye_matches(f, ratio)
    assert o['nMatches'] > 200 and o['descMatches'] >= o['nMatches']
    l2r, r2l = o['l2r'], o['r2l']
    assert (l2r[:f['mono_left']] == -1).all() and (r2l[:f['mono_right']] == -1).all()
    m = np.nonzero(l2r >= 0)[0]
    assert m.size == o['nMatches'] and (o['depth'][m] > 1e-4).all() and (o['depth'][l2r < 0] == -1).all()
    # mvRightToLeftMatch keeps the LAST left keypoint that chose a right one (src/Frame.cc:1450-1451)
    for r in np.unique(l2r[m]):
        assert r2l[r] == m[l2r[m] == r].max()
    assert (np.bincount(l2r[m]).max() >= 2)                         # the fixture does contain such duplicates
    # most accepted matches are the planted correspondences
    planted = f['src'][m - f['mono_left']] + f['mono_right']
    assert (planted == l2r[m]).mean() > 0.95
