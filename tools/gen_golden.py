#!/usr/bin/env python
"""Generates tests/golden/*.npz — golden vectors for the hot path, produced in the BUILD container from the reference's own
third-party arithmetic (OpenCV, through the cv2 4.13.0 wheel; the reference's C++ sources cannot be compiled here, see
DESIGN.md §oracle) and committed so that they travel to the GPU box, where neither cv2's presence nor /root/reference is assumed.

  extract_256x192.npz   image, per level: cv2.resize chain + cv2.copyMakeBorder(REFLECT_101) planes, per-cell cv2 FAST candidates
                        (ini 20 -> min 7 retry) in vToDistributeKeys order, cv2.GaussianBlur(7x7, 2) levels; final keypoints /
                        descriptors / return value of ORBextractor::operator() for two lapping areas.  The final vectors come
                        from the CPU oracle AFTER it was checked against the cv2-derived stages above in this very script
                        (the quadtree order and rBRIEF glue have no cv2 equivalent).
  atan2.npz             cv2.fastAtan2 on 4096 integer moment pairs
  knn2.npz              cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) on a tie-heavy descriptor set

usage: python tools/gen_golden.py        (needs cv2; deterministic)"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib as O                                     # noqa: E402
from openmavis_b200.synth import synth_image, synth_descriptors   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
W, H, NF, NL, INI, MIN = 256, 192, 400, 4, 20, 7


def cell_candidates(L, ini, mn):
    """src/ORBextractor.cc:704-794 with cv2.FAST on every cell ROI."""
    hh, ww = L.shape
    minB, maxBX, maxBY = 16, ww - 16, hh - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    out = []
    for i in range(nRows):
        iniY = minB + i * hCell
        if iniY >= maxBY - 3:
            continue
        maxY = min(iniY + hCell + 6, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            if iniX >= maxBX - 6:
                continue
            maxX = min(iniX + wCell + 6, maxBX)
            cell = np.ascontiguousarray(L[iniY:maxY, iniX:maxX])
            ks = cv2.FastFeatureDetector_create(ini, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(cell)
            if not ks:
                ks = cv2.FastFeatureDetector_create(mn, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(cell)
            out += [(int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)) for k in ks]
    return np.array(out, np.int32).reshape(-1, 3)


def main():
    os.makedirs(OUT, exist_ok=True)
    img = synth_image(W, H, 424242)
    E = O.OracleExtractor(NF, 1.2, NL, INI, MIN)
    t = E.tables()
    data = {'image': img, 'params': np.array([NF, NL, INI, MIN], np.int32)}
    cur = img
    res = {}
    for lap in ((0, 0), (0, 1000), (100, 180)):
        res[lap] = E(img, lap)
    for l in range(NL):
        if l:
            dw = int(np.rint(np.float32(W) * t['inv_sf'][l])); dh = int(np.rint(np.float32(H) * t['inv_sf'][l]))
            cur = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        plane = cv2.copyMakeBorder(cur, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        cand = cell_candidates(cur, INI, MIN)
        blur = cv2.GaussianBlur(cur.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        # the oracle must agree with OpenCV on every stage before its final vectors are accepted as golden
        assert np.array_equal(E.level(l), plane), l
        xs, ys, sc = E.level_candidates(l)
        assert np.array_equal(np.stack([xs, ys, sc], 1), cand), l
        if len(E.level_keypoints(l)):
            assert np.array_equal(E.level_blur(l), blur), l
        for k in E.level_keypoints(l)[:64]:
            x, y = int(k['x']), int(k['y'])
            m10 = m01 = 0
            for v in range(-15, 16):
                d = t['umax'][abs(v)]
                row = cur[y + v, x - d:x + d + 1].astype(np.int64)
                m10 += int((np.arange(-d, d + 1) * row).sum()); m01 += v * int(row.sum())
            assert np.float32(cv2.fastAtan2(float(m01), float(m10))) == k['angle']
        data['plane%d' % l] = plane; data['cand%d' % l] = cand; data['blur%d' % l] = blur
        data['lvl_kps%d' % l] = E.level_keypoints(l)
    for lap, (mono, kps, desc) in res.items():
        tag = '%d_%d' % lap
        data['mono_' + tag] = np.int32(mono); data['kps_' + tag] = kps; data['desc_' + tag] = desc
    np.savez_compressed(os.path.join(OUT, 'extract_256x192.npz'), **data)

    rng = np.random.default_rng(99)
    y = rng.integers(-2400000, 2400000, size=4096).astype(np.float32)
    x = rng.integers(-2400000, 2400000, size=4096).astype(np.float32)
    y[:6] = [0, 0, 5, -5, 7, -7]; x[:6] = [0, -5, 0, 0, 7, -7]
    a = np.array([cv2.fastAtan2(float(p), float(q)) for p, q in zip(y, x)], np.float32)
    np.savez_compressed(os.path.join(OUT, 'atan2.npz'), y=y, x=x, angle=a)

    tr = synth_descriptors(384, 1); tr[:128, 4:] = 0
    q = np.concatenate([tr[rng.integers(0, 384, 192)], synth_descriptors(64, 3)])
    q[:, :2] ^= rng.integers(0, 4, size=(256, 2), dtype=np.uint8)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, tr, 2)
    idx = np.array([[m.trainIdx for m in r] for r in ref], np.int32)
    dist = np.array([[int(m.distance) for m in r] for r in ref], np.int32)
    np.savez_compressed(os.path.join(OUT, 'knn2.npz'), query=q, train=tr, idx=idx, dist=dist)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
