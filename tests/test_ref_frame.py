"""oracle/_ref/libref_frame.so — the reference's OWN Frame::ComputeStereoMatches (src/Frame.cc:1057-1219),
MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483, N4) and ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2575-2591):
function bodies lifted by line range at build time and compiled unmodified inside stand-in classes (oracle/ref_frame_driver.cpp) —
against the restated oracle (orc_compute_stereo_matches, orc_distinctive_descriptors), which is what the device paths
orbx_compute_stereo_matches / orbx_distinctive_descriptors are tested against on the GPU box.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_descriptors, synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def RF():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'oracle', '_ref', 'libref_frame.so')
        if not os.path.exists(so):
            O.ref_lib()                                      # runs make when /root/reference is mounted
        if not os.path.exists(so):
            raise RuntimeError('%s is missing and /root/reference is not mounted to build it' % so)
        _lib = C.CDLL(so)
        _lib.reff_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib.reff_distinctive_descriptors.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]
        _lib.reff_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        _lib.reff_norm_l1.restype = C.c_double
        _lib.reff_norm_l1.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    return _lib


p = O.p


def stereo_pair(w, h, seed, base, amp):
    left = synth_image(w, h, seed)
    xs = np.arange(w)
    right = left[:, np.clip(xs + (base + amp * np.sin(xs / 60.0)).astype(np.int64), 0, w - 1)].copy()
    return left, right


@pytest.mark.parametrize('w,h,nf,seed,base,amp,mb,mbf', [(376, 240, 600, 2101, 9, 6, 0.11, 47.9), (376, 240, 600, 2102, 9, 6, 0.05, 20.0),
                                                        (752, 480, 1200, 2103, 14, 10, 0.11, 47.9), (512, 384, 900, 2104, 3, 3, 0.5, 12.0),
                                                        (640, 360, 1000, 2105, 25, 12, 0.11, 90.0)])
def test_compute_stereo_matches_reference_equals_oracle(w, h, nf, seed, base, amp, mb, mbf):
    """Row-band candidate tables, level gate, disparity window, Hamming best, the 11 x 11 SAD slide on the pyramid level of the left
    keypoint (cv::norm L1), parabola fit, disparity gates, the median-based outlier cut: mvuRight / mvDepth to the last bit."""
    L = RF(); Lo = O.lib()
    left, right = stereo_pair(w, h, seed, base, amp)
    EL = O.OracleExtractor(nf, 1.2, 8, 20, 7); ER = O.OracleExtractor(nf, 1.2, 8, 20, 7)
    _, kl, dl = EL(left, (0, 0)); _, kr, dr = ER(right, (0, 0))
    tab = EL.tables()
    levL = [EL.level(l) for l in range(8)]; levR = [ER.level(l) for l in range(8)]            # bordered planes (19 px frame)
    lw = np.array([a.shape[1] - 38 for a in levL], np.int32); lh = np.array([a.shape[0] - 38 for a in levL], np.int32)
    ls = np.array([a.strides[0] for a in levL], np.int32)
    PL = (C.c_void_p * 8)(*[a.ctypes.data + 19 * a.strides[0] + 19 for a in levL])               # the level interior, as mvImagePyramid[l] views it
    PR = (C.c_void_p * 8)(*[a.ctypes.data + 19 * a.strides[0] + 19 for a in levR])
    N = len(kl)
    ru = np.zeros(N, np.float32); rd = np.zeros(N, np.float32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    nr = L.reff_compute_stereo_matches(p(kl), p(dl), N, p(kr), p(dr), len(kr), p(tab['sf']), p(tab['inv_sf']), 8, PL, PR, p(lw), p(lh), p(ls), mb, mbf,
                                       p(ru), p(rd))
    ou = np.zeros(N, np.float32); od = np.zeros(N, np.float32)
    Lo.orc_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    no = Lo.orc_compute_stereo_matches(EL.h, ER.h, mb, mbf, p(ou), p(od))
    assert nr == no and nr > 40, (nr, no)
    assert np.array_equal(ru.view(np.uint32), ou.view(np.uint32))
    assert np.array_equal(rd.view(np.uint32), od.view(np.uint32))
    assert (ru[rd > 0] >= 0).all() and ((rd > 0) == (ru != -1)).all()


def test_distinctive_descriptors_reference_equals_oracle():
    """Observations of up to four cameras per keyframe, bad keyframes, bad map points, single observations, duplicates: the descriptor
    the reference stores in mDescriptor is the one the oracle's batched restatement selects (first minimum of the medians)."""
    L = RF(); Lo = O.lib()
    rng = np.random.default_rng(11)
    G = 400
    nkf = rng.integers(0, 13, G).astype(np.int32); nkf[::17] = 1; nkf[5::23] = 0
    ptr4 = np.concatenate([[0], np.cumsum(nkf)]).astype(np.int32)
    Q = int(ptr4[-1])
    kf_rows = rng.integers(4, 9, Q).astype(np.int32)
    kf_row0 = np.concatenate([[0], np.cumsum(kf_rows)])[:-1].astype(np.int32)
    total = int(kf_rows.sum())
    # descriptors of one map point are noisy copies of a common pattern (real medians), every keyframe holds a few unrelated rows too
    desc = synth_descriptors(total, 12)
    idx = np.full((Q, 4), -1, np.int32)
    bad = (rng.random(Q) < 0.1).astype(np.uint8)
    for g in range(G):
        base = synth_descriptors(1, 1000 + g)[0]
        for q in range(ptr4[g], ptr4[g + 1]):
            ncam = rng.choice([1, 1, 2, 4])
            cams = sorted(rng.choice(4, ncam, replace=False))
            rows = rng.choice(kf_rows[q], ncam, replace=False)
            for c, r in zip(cams, rows):
                idx[q, c] = r
                desc[kf_row0[q] + r] = base ^ np.packbits(rng.random(256) < rng.choice([0.03, 0.1, 0.3]), bitorder='little')
    # exact duplicates inside a group (ties of the medians: the first one wins)
    for g in range(0, G, 9):
        qs = [q for q in range(ptr4[g], ptr4[g + 1]) if not bad[q]]
        if len(qs) >= 2:
            a, b = qs[0], qs[-1]
            desc[kf_row0[b] + idx[b][idx[b] >= 0][0]] = desc[kf_row0[a] + idx[a][idx[a] >= 0][0]]
    desc = np.ascontiguousarray(desc); idx = np.ascontiguousarray(idx)
    every = 10
    best_ref = np.zeros(G, np.int32)
    L.reff_distinctive_descriptors(p(desc), p(ptr4), p(kf_row0), p(kf_rows), p(idx), p(bad), G, every, p(best_ref))
    # the oracle takes the descriptors of a group in the order the reference pushes them: keyframes in map order, cameras left .. side-right
    rows_of = []
    for g in range(G):
        rs = []
        if g % every != every - 1:
            for q in range(ptr4[g], ptr4[g + 1]):
                if not bad[q]:
                    rs += [int(kf_row0[q] + idx[q, c]) for c in range(4) if idx[q, c] != -1]
        rows_of.append(rs)
    gptr = np.concatenate([[0], np.cumsum([len(r) for r in rows_of])]).astype(np.int32)
    gdesc = np.ascontiguousarray(desc[[r for rs in rows_of for r in rs]]) if gptr[-1] else np.zeros((0, 32), np.uint8)
    best = np.zeros(G, np.int32)
    Lo.orc_distinctive_descriptors(p(gdesc), p(gptr), G, p(best))
    expect = np.array([rows_of[g][best[g]] if best[g] >= 0 else -1 for g in range(G)], np.int32)
    # (two rows with the same bytes are the same descriptor: compare bytes where the row numbers differ)
    same = best_ref == expect
    for g in np.flatnonzero(~same):
        assert best_ref[g] >= 0 and expect[g] >= 0 and np.array_equal(desc[best_ref[g]], desc[expect[g]]), g
    assert (expect >= 0).sum() > 250 and (expect < 0).sum() > 40
    sizes = np.diff(gptr)
    assert (sizes == 1).any() and (sizes >= 20).any()


def test_descriptor_distance_reference_is_the_popcount():
    L = RF()
    a = synth_descriptors(400, 3); b = synth_descriptors(400, 4)
    b[:5] = a[:5]; b[5:10] = ~a[5:10]
    for i in range(400):
        assert L.reff_descriptor_distance(p(a[i]), p(b[i])) == O.hamming(a[i], b[i])


def test_shim_norm_l1_is_cv2_norm():
    """The one OpenCV primitive ComputeStereoMatches adds (cv::norm(IL, IR, NORM_L1) on 11 x 11 views, src/Frame.cc:1167): the stand-in the
    lifted body is compiled against returns what cv2.norm returns, on strided views."""
    cv2 = pytest.importorskip('cv2')
    L = RF()
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (64, 80), dtype=np.uint8); B = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    A[:8] = 255; B[:8] = 0
    for (y, x, yb, xb, r, c) in [(0, 0, 0, 0, 11, 11), (5, 7, 9, 30, 11, 11), (20, 3, 1, 60, 11, 11), (0, 0, 0, 0, 64, 80), (40, 40, 40, 40, 1, 1)]:
        a = A[y:y + r, x:x + c]; b = B[yb:yb + r, xb:xb + c]
        got = L.reff_norm_l1(a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0], r, c)
        assert got == cv2.norm(np.ascontiguousarray(a), np.ascontiguousarray(b), cv2.NORM_L1)
