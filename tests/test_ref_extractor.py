"""oracle/_ref — the reference's OWN src/ORBextractor.cc, compiled unmodified from /root/reference against oracle/cvshim
(oracle/Makefile) — against the restated oracle.  This is what pins the oracle to the reference itself: the quadtree order
(ORBextractor.cc:496-702, std::list + std::sort), the IC_Angle / rBRIEF glue (:19-90), the cell loop (:704-794), the lapping
partition (:1024-1068), the ctor tables (:351-414) are the reference's object code here; only cv::resize / copyMakeBorder /
FAST / GaussianBlur / fastAtan2 come from the cv2-pinned oracle primitives (tests/test_oracle_vs_cv2.py).
CPU only; the GPU tests repeat the comparison CUDA-vs-_ref (tests/test_gpu_extract.py::compare_with_reference)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_image

KPF = ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')

# BASELINE.json configs 1-5 (+ the yaml-faithful Hilti parameters): (w, h, nfeatures, iniTh, minTh, lapping, seed)
CONFIGS = [
    (752, 480, 1000, 20, 7, (0, 1000), 1000), (752, 480, 1000, 20, 7, (0, 0), 1000), (752, 480, 1000, 20, 7, (300, 500), 1001),
    (752, 480, 1200, 20, 7, (0, 0), 2000), (752, 480, 1200, 20, 7, (0, 0), 2001),
    (720, 540, 1500, 20, 7, (0, 720), 3000), (720, 540, 1500, 20, 7, (0, 0), 3002), (720, 540, 500, 15, 7, (0, 720), 3010),
    (720, 540, 1000, 15, 7, (0, 0), 3012),
    (512, 512, 1500, 20, 7, (0, 511), 4000), (512, 512, 1500, 20, 7, (0, 511), 4001),
    (1280, 1024, 2000, 20, 7, (0, 0), 5000),
]


def assert_same(ref_out, orc_out, what=''):
    rm, rk, rd = ref_out
    om, ok, od = orc_out
    assert rm == om and len(rk) == len(ok), '%s: mono %d/%d, n %d/%d' % (what, rm, om, len(rk), len(ok))
    for f in KPF:
        assert np.array_equal(rk[f].view(np.uint32), ok[f].view(np.uint32)), '%s: field %s' % (what, f)
    assert np.array_equal(rd, od), '%s: %d descriptor rows differ' % (what, int((rd != od).any(axis=1).sum()))


@pytest.mark.parametrize('w,h,nf,ini,mn,lap,seed', CONFIGS)
def test_reference_equals_oracle_on_baseline_configs(w, h, nf, ini, mn, lap, seed):
    img = synth_image(w, h, seed)
    R = O.RefExtractor(nf, 1.2, 8, ini, mn)
    E = O.OracleExtractor(nf, 1.2, 8, ini, mn)
    assert_same(R(img, lap), E(img, lap), 'operator()')
    for l in range(8):
        assert np.array_equal(R.level(l), E.level(l)), 'mvImagePyramid[%d] incl. frame' % l
    for l, rk in enumerate(R.stages(img)):                  # after ComputeKeyPointsOctTree
        ok = E.level_keypoints(l)
        assert len(rk) == len(ok), l
        for f in KPF:
            assert np.array_equal(rk[f].view(np.uint32), ok[f].view(np.uint32)), (l, f)
    t, to = R.tables(), E.tables()
    for k in t:
        assert np.array_equal(t[k], to[k]), k


@pytest.mark.parametrize('nf,sf,nl', [(1000, 1.2, 8), (500, 1.2, 8), (2000, 1.2, 8), (5000, 1.2, 8), (900, 1.3, 6), (600, 2.0, 3), (1500, 1.1, 10)])
def test_constructor_tables(nf, sf, nl):
    t, to = O.RefExtractor(nf, sf, nl, 20, 7).tables(), O.OracleExtractor(nf, sf, nl, 20, 7).tables()
    for k in t:
        assert np.array_equal(t[k].view(np.uint32), to[k].view(np.uint32)), k


@pytest.mark.parametrize('kind', ['zeros', 'const255', 'checker', 'noise', 'halfflat', 'gradient'])
def test_adversarial_images(kind):
    w, h = 640, 480
    rng = np.random.default_rng(9)
    if kind == 'zeros':
        img = np.zeros((h, w), np.uint8)
    elif kind == 'const255':
        img = np.full((h, w), 255, np.uint8)
    elif kind == 'checker':                                   # massive score ties
        yy, xx = np.mgrid[0:h, 0:w]
        img = ((((yy // 8) + (xx // 8)) % 2) * 200 + 20).astype(np.uint8)
    elif kind == 'noise':                                     # 60k+ candidates
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    elif kind == 'gradient':
        img = np.tile((np.arange(w) % 256).astype(np.uint8), (h, 1))
    else:
        img = synth_image(w, h, 17); img[:, : w // 2] = 90
    R = O.RefExtractor(1000, 1.2, 8, 20, 7)
    E = O.OracleExtractor(1000, 1.2, 8, 20, 7)
    out = R(img, (100, 400))
    assert_same(out, E(img, (100, 400)), kind)
    if kind in ('zeros', 'const255'):
        assert out[0] == 0 and len(out[1]) == 0
    for l in range(8):
        assert np.array_equal(R.level(l), E.level(l))


def test_empty_image_and_strided_view():
    R = O.RefExtractor(1000, 1.2, 8, 20, 7)
    assert R(None)[0] == -1                                   # ORBextractor.cc:994-995
    big = synth_image(800, 500, 21)
    view = big[10:490, 24:776]                                # 752x480, row stride 800
    E = O.OracleExtractor(1000, 1.2, 8, 20, 7)
    assert_same(R(view, (0, 0)), E(np.ascontiguousarray(view), (0, 0)), 'strided')


def test_other_pyramid_shapes():
    for (w, h, nl, sf, nf) in [(751, 479, 8, 1.2, 800), (333, 247, 5, 1.2, 500), (1000, 341, 6, 1.3, 900), (640, 480, 3, 2.0, 600)]:
        img = synth_image(w, h, 9000 + w)
        assert_same(O.RefExtractor(nf, sf, nl, 20, 7)(img, (w // 3, w // 2)), O.OracleExtractor(nf, sf, nl, 20, 7)(img, (w // 3, w // 2)), str((w, h)))


def test_distribute_octtree_against_reference_member():
    """ORBextractor::DistributeOctTree (the protected member, called through a derived class) on synthetic candidate sets:
    tie-heavy coordinates and responses, 1-3 root nodes, N from tiny to larger than the set."""
    rng = np.random.default_rng(77)
    R = O.RefExtractor(1000, 1.2, 8, 20, 7)
    for trial in range(300):
        W = int(rng.integers(60, 900)); H = int(rng.integers(60, 500))
        if W < H * 0.6:
            W = int(H * 0.6) + 1                              # nIni = round(W/H) must be >= 1 (the reference divides by it)
        n = int(rng.integers(1, 1500))
        step = int(rng.choice([1, 1, 2, 5]))                  # coarse lattices create many (size, UL.x) ties
        xs = (rng.integers(0, max(1, W // step), n) * step).astype(np.int32)
        ys = (rng.integers(0, max(1, H // step), n) * step).astype(np.int32)
        sc = rng.integers(7, 60 if trial % 2 else 255, n).astype(np.int32)
        N = int(rng.choice([1, 5, 31, 60, 109, 217, 326, 434, 2000]))
        ox, oy, osc = R.distribute(xs, ys, sc, 16, 16 + W, 16, 16 + H, N)
        sel = O.distribute(xs, ys, sc, 16, 16 + W, 16, 16 + H, N)
        assert len(sel) == len(ox), (trial, len(sel), len(ox))
        assert np.array_equal(xs[sel], ox) and np.array_equal(ys[sel], oy) and np.array_equal(sc[sel], osc), trial


def test_fma_contraction_question():
    """g++ contracts GET_VALUE's x*b + y*a (ORBextractor.cc:56-57) under the reference's own flags.  Both builds
    of the unmodified file — -ffp-contract=off and the default — must equal the oracle in the matching fma_mode, and the two
    modes are compared on 6 frames to record how often a contraction actually flips a descriptor bit."""
    rows = diff = 0
    for seed in range(6):
        img = synth_image(752, 480, 7100 + seed)
        E = O.OracleExtractor(1200, 1.2, 8, 20, 7)
        a = O.RefExtractor(1200, 1.2, 8, 20, 7, fma=False)(img, (0, 0))
        b = O.RefExtractor(1200, 1.2, 8, 20, 7, fma=True)(img, (0, 0))
        assert_same(a, E(img, (0, 0), fma_mode=0), 'contract=off')
        assert_same(b, E(img, (0, 0), fma_mode=1), 'contract=fast')
        rows += len(a[2]); diff += int((a[2] != b[2]).any(axis=1).sum())
    assert diff <= rows * 1e-3, (diff, rows)                 # the north star's ">= 99.9 % descriptors bit-identical" allowance


def test_golden_fixture_is_reference_output():
    """tests/golden/extract_256x192.npz (committed, travels to the GPU box) equals what the compiled reference produces."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'extract_256x192.npz'))
    nf, nl, ini, mn = g['params'].tolist()
    R = O.RefExtractor(nf, 1.2, nl, ini, mn)
    for lap in ((0, 0), (0, 1000), (100, 180)):
        mono, kps, desc = R(g['image'], lap)
        tag = '%d_%d' % lap
        assert mono == int(g['mono_' + tag]) and np.array_equal(kps, g['kps_' + tag]) and np.array_equal(desc, g['desc_' + tag])
    for l in range(nl):
        assert np.array_equal(R.level(l), g['plane%d' % l])
    for l, k in enumerate(R.stages(g['image'])):
        assert np.array_equal(k, g['lvl_kps%d' % l])
