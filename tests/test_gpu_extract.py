"""GPU parity tests (run on the B200 box): the CUDA extraction path through the C ABI vs the CPU oracle,
stage by stage and end to end, bit-exact.  Reference path: src/ORBextractor.cc:987-1071."""
import os

import numpy as np
import pytest

import oracle_lib as O
import openmavis_b200 as ob
from openmavis_b200.synth import synth_image

pytestmark = pytest.mark.gpu


def _need_gpu():
    if ob.device_count() == 0:
        pytest.fail('no CUDA device: -m gpu tests must run on the GPU box')


def compare_image(rig, b, img, nf, ini, mn, lap, result, fma_mode=0, stages=True):
    E = O.OracleExtractor(nf, 1.2, 8, ini, mn)
    mono, kps, desc = E(img, lap, fma_mode)
    if stages:
        for l in range(8):
            ref = E.level(l)
            got = rig.pyramid_level(b, l, with_border=True)
            assert np.array_equal(ref, got), 'pyramid level %d differs (%d px)' % (l, int((ref != got).sum()))
            assert np.array_equal(ref[19:-19, 19:-19], rig.pyramid_level(b, l))
            xs, ys, sc = E.level_candidates(l)
            gx, gy, gs = rig.debug_candidates(b, l)
            assert len(xs) == len(gx), 'level %d: %d vs %d candidates' % (l, len(xs), len(gx))
            assert np.array_equal(xs, gx) and np.array_equal(ys, gy) and np.array_equal(sc, gs), 'candidates level %d' % l
            rb = E.level_blur(l)
            if rb is not None:
                assert np.array_equal(rb, rig.debug_blur_level(b, l)), 'blur level %d' % l
            rk = E.level_keypoints(l)
            gk = rig.debug_level_keypoints(b, l)
            assert len(rk) == len(gk), 'level %d: %d vs %d keypoints' % (l, len(rk), len(gk))
            for f in ('x', 'y', 'response', 'octave', 'size'):
                assert np.array_equal(rk[f], gk[f]), 'level %d keypoint field %s' % (l, f)
            assert np.array_equal(rk['angle'].view(np.uint32), gk['angle'].view(np.uint32)), 'angle bits level %d' % l
    gmono, gk, gd = result
    assert gmono == mono
    assert len(gk) == len(kps)
    for f in KPF:
        assert np.array_equal(kps[f].view(np.uint32), gk[f].view(np.uint32)), 'final keypoint field %s' % f
    assert np.array_equal(desc, gd), '%d descriptor rows differ' % int((desc != gd).any(axis=1).sum())
    compare_with_reference(rig, b, img, nf, ini, mn, lap, result, fma_mode, stages)
    return len(kps)


def compare_with_reference(rig, b, img, nf, ini, mn, lap, result, fma_mode=0, stages=True):
    """The CUDA result against oracle/_ref — the reference's own src/ORBextractor.cc compiled unmodified (oracle/Makefile):
    operator() outputs, and with `stages` the public mvImagePyramid planes incl. their 19-px frame and the per-level
    keypoints ComputeKeyPointsOctTree leaves (position, response, octave, size, angle bits)."""
    R = O.RefExtractor(nf, 1.2, 8, ini, mn, fma=bool(fma_mode))
    rmono, rk, rd = R(img, lap)
    gmono, gk, gd = result
    assert gmono == rmono and len(gk) == len(rk), 'reference: %d/%d keypoints, mono %d/%d' % (len(gk), len(rk), gmono, rmono)
    for f in KPF:
        assert np.array_equal(rk[f].view(np.uint32), gk[f].view(np.uint32)), 'reference: final keypoint field %s' % f
    assert np.array_equal(rd, gd), 'reference: %d descriptor rows differ' % int((rd != gd).any(axis=1).sum())
    if stages:
        for l in range(8):
            assert np.array_equal(R.level(l), rig.pyramid_level(b, l, with_border=True)), 'reference: pyramid level %d' % l
        for l, rkl in enumerate(R.stages(img)):
            gkl = rig.debug_level_keypoints(b, l)
            assert len(rkl) == len(gkl), 'reference: level %d keypoint count' % l
            for f in ('x', 'y', 'response', 'octave', 'size', 'angle'):
                assert np.array_equal(rkl[f].view(np.uint32), gkl[f].view(np.uint32)), 'reference: level %d field %s' % (l, f)


KPF = ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')


# config 1 (BASELINE.json): one 752x480 EuRoC-mono frame, nF=1000, 8 levels, 1.2, FAST 20/7; mono lapping {0,1000} (D7) and {0,0}
@pytest.mark.parametrize('lap', [(0, 1000), (0, 0), (300, 500)])
def test_config1_euroc_mono(lap):
    _need_gpu()
    img = synth_image(752, 480, 1000)
    rig = ob.RigExtractor(752, 480, 1, 1000)
    res = rig.extract([img], ob.make_params(1, 1000, 20, 7, lap))
    n = compare_image(rig, 0, img, 1000, 20, 7, lap, res[0])
    assert n >= 900


def test_operator_mirror_single_image():
    _need_gpu()
    img = synth_image(752, 480, 1001)
    ex = ob.ORBextractor(1000, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img, None, (0, 1000))
    rm, rk, rd = O.OracleExtractor(1000, 1.2, 8, 20, 7)(img, (0, 1000))
    assert mono == rm and np.array_equal(desc, rd)
    for f in KPF:
        assert np.array_equal(kps[f], rk[f])
    assert ex(None)[0] == -1 and ex(np.zeros((0, 0), np.uint8))[0] == -1
    pyr = ex.mvImagePyramid
    assert len(pyr) == 8 and pyr[0].shape == (480, 752) and np.array_equal(pyr[0], img)
    assert ex.GetLevels() == 8 and np.allclose(ex.GetScaleFactors()[1], 1.2)
    t = O.OracleExtractor(1000, 1.2, 8, 20, 7).tables()
    assert np.array_equal(ex.GetScaleFactors(), t['sf']) and np.array_equal(ex.GetInverseScaleSigmaSquares(), t['inv_sigma2'])


# config 2: EuRoC stereo pair 2x752x480, nF=1200
def test_config2_stereo_pair():
    _need_gpu()
    imgs = [synth_image(752, 480, 2000 + c) for c in range(2)]
    rig = ob.RigExtractor(752, 480, 2, 1200)
    res = rig.extract(imgs, ob.make_params(2, 1200, 20, 7, (0, 0)))
    for b in range(2):
        compare_image(rig, b, imgs[b], 1200, 20, 7, (0, 0), res[b])


# config 3: Hilti rig 5x720x540, nF=1500/camera; and the yaml-faithful 4-camera wiring 500/500/1000/1000, ini 15
def test_config3_rig_5cam():
    _need_gpu()
    imgs = [synth_image(720, 540, 3000 + c) for c in range(5)]
    laps = [(0, 720), (0, 720), (0, 0), (0, 0), (0, 0)]
    rig = ob.RigExtractor(720, 540, 5, 1500)
    res = rig.extract(imgs, ob.make_params(5, 1500, 20, 7, laps))
    for b in range(5):
        compare_image(rig, b, imgs[b], 1500, 20, 7, laps[b], res[b], stages=(b < 2))


def test_config3_rig_4cam_yaml_params():
    _need_gpu()
    imgs = [synth_image(720, 540, 3010 + c) for c in range(4)]
    nfs = [500, 500, 1000, 1000]
    laps = [(0, 720), (0, 720), (0, 0), (0, 0)]
    rig = ob.RigExtractor(720, 540, 4, 1000)
    p = ob.make_params(4, 500, 15, 7, laps); p['nfeatures'] = nfs
    res = rig.extract(imgs, p)
    for b in range(4):
        compare_image(rig, b, imgs[b], nfs[b], 15, 7, laps[b], res[b], stages=(b in (0, 3)))


# config 4: TUM-VI 2x512x512, nF=1500, lapping {0,511}
def test_config4_tumvi():
    _need_gpu()
    imgs = [synth_image(512, 512, 4000 + c) for c in range(2)]
    rig = ob.RigExtractor(512, 512, 2, 1500)
    res = rig.extract(imgs, ob.make_params(2, 1500, 20, 7, (0, 511)))
    for b in range(2):
        compare_image(rig, b, imgs[b], 1500, 20, 7, (0, 511), res[b])


# config 5 geometry: 1280x1024, nF=2000 (small batch here; the 64-frame batch is exercised by bench.py)
def test_config5_geometry_small_batch():
    _need_gpu()
    imgs = [synth_image(1280, 1024, 5000 + c) for c in range(3)]
    rig = ob.RigExtractor(1280, 1024, 3, 2000)
    res = rig.extract(imgs, ob.make_params(3, 2000, 20, 7, (0, 0)))
    for b in range(3):
        compare_image(rig, b, imgs[b], 2000, 20, 7, (0, 0), res[b], stages=(b == 0))


def test_fma_mode_switch():
    _need_gpu()
    img = synth_image(752, 480, 1003)
    rig = ob.RigExtractor(752, 480, 1, 1000, fma_mode=1)
    res = rig.extract([img], ob.make_params(1, 1000, 20, 7, (0, 0)))
    compare_image(rig, 0, img, 1000, 20, 7, (0, 0), res[0], fma_mode=1, stages=False)


@pytest.mark.parametrize('kind', ['zeros', 'const255', 'checker', 'noise', 'halfflat'])
def test_adversarial_images(kind):
    _need_gpu()
    w, h = 640, 480
    rng = np.random.default_rng(9)
    if kind == 'zeros':
        img = np.zeros((h, w), np.uint8)
    elif kind == 'const255':
        img = np.full((h, w), 255, np.uint8)
    elif kind == 'checker':
        yy, xx = np.mgrid[0:h, 0:w]
        img = ((((yy // 8) + (xx // 8)) % 2) * 200 + 20).astype(np.uint8)
    elif kind == 'noise':
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    else:
        img = synth_image(w, h, 17); img[:, : w // 2] = 90
    rig = ob.RigExtractor(w, h, 1, 1000)
    res = rig.extract([img], ob.make_params(1, 1000, 20, 7, (100, 400)))
    n = compare_image(rig, 0, img, 1000, 20, 7, (100, 400), res[0])
    if kind in ('zeros', 'const255'):
        assert n == 0 and res[0][2].shape == (0, 32) and res[0][0] == 0


def test_non_contiguous_input_and_handle_reuse():
    _need_gpu()
    big = synth_image(800, 500, 21)
    view = big[10:490, 24:776]           # 752x480 view with stride 800
    rig = ob.RigExtractor(752, 480, 2, 1000)
    res = rig.extract([view], ob.make_params(1, 1000, 20, 7, (0, 0)))
    compare_image(rig, 0, np.ascontiguousarray(view), 1000, 20, 7, (0, 0), res[0], stages=False)
    # second call on the same handle with other content and a smaller nfeatures: no state leaks between calls
    img2 = synth_image(752, 480, 22)
    res2 = rig.extract([img2, view], ob.make_params(2, 600, 12, 5, (0, 0)))
    compare_image(rig, 0, img2, 600, 12, 5, (0, 0), res2[0])
    compare_image(rig, 1, np.ascontiguousarray(view), 600, 12, 5, (0, 0), res2[1], stages=False)


def test_device_resident_path_matches_host_path():
    _need_gpu()
    import torch
    imgs = np.stack([synth_image(720, 540, 3100 + c) for c in range(4)])
    rig = ob.RigExtractor(720, 540, 4, 1500)
    p = ob.make_params(4, 1500, 20, 7, (0, 720))
    host = rig.extract(list(imgs), p)
    host = [(m, k.copy(), d.copy()) for m, k, d in host]
    t = torch.from_numpy(imgs).cuda()
    rig.set_stream(torch.cuda.current_stream().cuda_stream)
    rig.extract_device(t.data_ptr(), 4, p)
    dev = rig.fetch(4)
    for (m0, k0, d0), (m1, k1, d1) in zip(host, dev):
        assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)


def test_errors():
    _need_gpu()
    with pytest.raises(ob.OrbxError):
        ob.RigExtractor(100, 80, 1, 1000)             # too small for 8 levels
    rig = ob.RigExtractor(752, 480, 1, 1000)
    with pytest.raises(ob.OrbxError):
        rig.extract([synth_image(752, 480, 1)] * 2, ob.make_params(2, 1000, 20, 7))    # batch > max_images
    with pytest.raises(ob.OrbxError):
        rig.extract([synth_image(752, 480, 1)], ob.make_params(1, 5000, 20, 7))        # nfeatures > max_nfeatures
    with pytest.raises(ob.OrbxError):
        rig.extract([synth_image(640, 480, 1)], ob.make_params(1, 1000, 20, 7))        # wrong geometry


def test_chunked_pipeline_equals_oracle_and_single_chunk():
    """A batch larger than one chunk runs as overlapping chunks on several streams (host copies in / out included): every image
    must still equal the oracle, whatever chunk it fell into, for host buffers, device buffers and profiling (serial) mode."""
    _need_gpu()
    import torch
    n = 37                                                   # not a multiple of the chunk size: 16 + 16 + 5 host-side
    base = [synth_image(512, 384, 5200 + c) for c in range(5)]
    imgs = np.stack([np.roll(base[i % 5], (3 * i, 7 * i), axis=(0, 1)) for i in range(n)])
    laps = [(0, 0), (0, 512), (100, 300)]
    p = ob.make_params(n, 600, 20, 7, [laps[i % 3] for i in range(n)])
    p['nfeatures'][::4] = 450                                # per-image feature targets
    rig = ob.RigExtractor(512, 384, n, 600)
    host = [(m, k.copy(), d.copy()) for m, k, d in rig.extract(list(imgs), p)]
    for i in (0, 15, 16, 31, 32, 36):                        # chunk borders
        E = O.OracleExtractor(int(p['nfeatures'][i]), 1.2, 8, 20, 7)
        m, k, d = E(imgs[i], (int(p['lap0'][i]), int(p['lap1'][i])))
        assert host[i][0] == m and np.array_equal(host[i][1], k) and np.array_equal(host[i][2], d), i
    t = torch.from_numpy(imgs).cuda()
    for prof in (False, True):
        rig.set_profiling(prof)
        rig.extract_device(t.data_ptr(), n, p)
        dev = rig.fetch(n)
        for (m0, k0, d0), (m1, k1, d1) in zip(host, dev):
            assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)
    assert sum(rig.stage_ms().values()) > 0


# Frame::ComputeStereoMatches on the device-resident pyramids / keypoints / descriptors of a stereo pair
@pytest.mark.parametrize('bf', [(0.11, 47.9), (0.5, 12.0)])
def test_compute_stereo_matches(bf):
    _need_gpu()
    mb, mbf = bf
    left = synth_image(752, 480, 2001)
    # rectified right image: the left one shifted by a disparity that varies smoothly along x (3..35 px) plus mild noise
    xs = np.arange(752)
    disp = (19 + 16 * np.sin(xs / 90.0)).astype(np.int64)
    right = left[:, np.clip(xs + disp, 0, 751)]
    right = np.clip(right.astype(np.int32) + np.random.default_rng(3).integers(-2, 3, right.shape), 0, 255).astype(np.uint8)
    rig = ob.RigExtractor(752, 480, 2, 1200)
    res = rig.extract([left, right], ob.make_params(2, 1200, 20, 7, (0, 0)))
    n, u, d = rig.compute_stereo_matches(0, 1, mb, mbf)
    EL = O.OracleExtractor(1200, 1.2, 8, 20, 7); ER = O.OracleExtractor(1200, 1.2, 8, 20, 7)
    EL(left, (0, 0)); ER(right, (0, 0))
    N = len(res[0][1])
    ru = np.zeros(N, np.float32); rd = np.zeros(N, np.float32)
    O.lib().orc_compute_stereo_matches.argtypes = [O.C.c_void_p, O.C.c_void_p, O.C.c_float, O.C.c_float, O.C.c_void_p, O.C.c_void_p]
    rn = O.lib().orc_compute_stereo_matches(EL.h, ER.h, mb, mbf, O.p(ru), O.p(rd))
    assert rn == n and n > 100, (rn, n)
    assert np.array_equal(ru.view(np.uint32), u[:N].view(np.uint32)), int((ru != u[:N]).sum())
    assert np.array_equal(rd.view(np.uint32), d[:N].view(np.uint32))
    # the same pair held by two extractor handles (mpORBextractorLeft / mpORBextractorRight of the reference)
    a = ob.RigExtractor(752, 480, 1, 1200); b2 = ob.RigExtractor(752, 480, 1, 1200)
    a.extract([left], ob.make_params(1, 1200, 20, 7, (0, 0))); b2.extract([right], ob.make_params(1, 1200, 20, 7, (0, 0)))
    n2, u2, d2 = a.compute_stereo_matches(0, 0, mb, mbf, right=b2)
    assert n2 == n and np.array_equal(u2[:N].view(np.uint32), u[:N].view(np.uint32)) and np.array_equal(d2[:N].view(np.uint32), d[:N].view(np.uint32))


@pytest.mark.parametrize('seed', range(int(os.environ.get('ORBX_STEREO_N', '6'))))
def test_compute_stereo_matches_random_pairs(seed):
    """ComputeStereoMatches (src/Frame.cc:1057-1219) over random image sizes, feature budgets, disparity fields (incl. small and negative
    disparities, which the reference rejects), noise levels and baselines: match count, mvuRight and mvDepth bits == the oracle."""
    _need_gpu()
    rng = np.random.default_rng(31000 + seed)
    w = int(rng.integers(400, 900)); h = int(rng.integers(300, 600)); nf = int(rng.integers(300, 1800))
    left = synth_image(w, h, 7000 + seed)
    xs = np.arange(w)
    disp = (rng.uniform(-2, 25) + rng.uniform(2, 20) * np.sin(xs / rng.uniform(40, 200) + rng.uniform(0, 6))).astype(np.int64)
    right = left[:, np.clip(xs + disp, 0, w - 1)]
    amp = int(rng.integers(0, 6))
    right = np.clip(right.astype(np.int32) + rng.integers(-amp, amp + 1, right.shape), 0, 255).astype(np.uint8)
    if rng.random() < 0.3:
        right = np.roll(right, int(rng.integers(-2, 3)), axis=0)       # imperfect rectification: rows off by up to 2
    mb = float(rng.uniform(0.05, 0.6)); mbf = float(mb * rng.uniform(200, 500))
    rig = ob.RigExtractor(w, h, 2, nf)
    res = rig.extract([left, right], ob.make_params(2, nf, 20, 7, (0, 0)))
    n, u, d = rig.compute_stereo_matches(0, 1, mb, mbf)
    EL = O.OracleExtractor(nf, 1.2, 8, 20, 7); ER = O.OracleExtractor(nf, 1.2, 8, 20, 7)
    EL(left, (0, 0)); ER(right, (0, 0))
    N = len(res[0][1])
    ru = np.zeros(N, np.float32); rd = np.zeros(N, np.float32)
    O.lib().orc_compute_stereo_matches.argtypes = [O.C.c_void_p, O.C.c_void_p, O.C.c_float, O.C.c_float, O.C.c_void_p, O.C.c_void_p]
    rn = O.lib().orc_compute_stereo_matches(EL.h, ER.h, mb, mbf, O.p(ru), O.p(rd))
    assert rn == n, (seed, rn, n)
    assert np.array_equal(ru.view(np.uint32), u[:N].view(np.uint32)), (seed, int((ru != u[:N]).sum()))
    assert np.array_equal(rd.view(np.uint32), d[:N].view(np.uint32)), seed


@pytest.mark.parametrize('w,h,nl,sf,nf', [(751, 479, 8, 1.2, 800), (333, 247, 5, 1.2, 500), (1000, 341, 6, 1.3, 900), (640, 480, 3, 2.0, 600),
                                            (1083, 203, 3, 1.2, 900), (1119, 150, 2, 1.2, 600), (700, 1083, 2, 1.2, 900)])
def test_odd_geometries_and_pyramid_shapes(w, h, nl, sf, nf):
    """Widths / strides that are not multiples of 4 or 16, wide aspect ratios (2-3 quadtree roots), other level counts and
    scale factors, level widths whose last FAST cell is ONE pixel wide (1083, 1119): planes, tiles and taps are derived per handle, nothing
    may assume the benchmark geometry."""
    _need_gpu()
    img = synth_image(w, h, 9000 + w)
    if w in (1083, 1119):   # isolated bright pixels (FAST corners) in the one tested column of the last cell of level 0, and its mirror
        n = (w - 32) // 35      # (src/ORBextractor.cc:729-751: borders 16, nCols, wCell)
        xt = 16 + (n - 1) * -(-(w - 32) // n) + 3
        for y in range(30, h - 30, 23):
            for x in (xt, w - 1 - xt):
                img[y - 4:y + 5, x - 4:x + 5] = 20
                img[y, x] = 255
    if h == 1083:           # the same for the ONE tested row of the last cell row
        n = (h - 32) // 35
        yt = 16 + (n - 1) * -(-(h - 32) // n) + 3
        for x in range(30, w - 30, 23):
            img[yt - 4:yt + 5, x - 4:x + 5] = 20
            img[yt, x] = 255
    rig = ob.RigExtractor(w, h, 2, nf, sf, nl)
    res = rig.extract([img, img[:, ::-1].copy()], ob.make_params(2, nf, 20, 7, [(0, 0), (w // 3, w // 2)]))
    E = O.OracleExtractor(nf, sf, nl, 20, 7)
    for b, (im, lap) in enumerate(((img, (0, 0)), (img[:, ::-1].copy(), (w // 3, w // 2)))):
        mono, kps, desc = E(im, lap)
        for l in range(nl):
            assert np.array_equal(E.level(l), rig.pyramid_level(b, l, with_border=True)), (b, l)
            xs, ys, sc = E.level_candidates(l)
            gx, gy, gs = rig.debug_candidates(b, l)
            assert np.array_equal(xs, gx) and np.array_equal(ys, gy) and np.array_equal(sc, gs), (b, l)
        assert res[b][0] == mono and np.array_equal(res[b][1], kps) and np.array_equal(res[b][2], desc), b


def test_staging_fallback_paths_without_tma(monkeypatch):
    """The tiled kernels stage through TMA tensor copies when the driver provides cuTensorMapEncodeTiled and through cp.async /
    vector loads otherwise (or for geometries wider than a TMA box): both paths must give the same bits."""
    _need_gpu()
    img = synth_image(720, 540, 8123)
    p = ob.make_params(1, 1500, 20, 7, (0, 720))
    a = ob.RigExtractor(720, 540, 1, 1500).extract([img], p)[0]
    monkeypatch.setenv('ORBX_NO_TMA', '1')
    monkeypatch.setenv('ORBX_NO_PATCH_TMA', '1')
    rig = ob.RigExtractor(720, 540, 1, 1500)                     # the knobs are read when the handle is created
    b = rig.extract([img], p)[0]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    compare_image(rig, 0, img, 1500, 20, 7, (0, 720), b)


@pytest.mark.parametrize('ring,fma', [(3, 0), (4, 0), (4, 1), (5, 0)])
def test_batched_orientation_descriptor_kernel(monkeypatch, ring, fma):
    """k_orient_desc_batch (one warp per group of 16 keypoints, TMA patch ring, per-lane atan2 / sincos, conversion-free sample loop;
    opt-in through ORBX_DESC_BATCH because it was measured no faster than one warp per keypoint) gives the same bits: angles, keypoint
    records, descriptors — against the oracle and the compiled reference, in both float-contraction modes."""
    _need_gpu()
    monkeypatch.setenv('ORBX_DESC_BATCH', '1')
    monkeypatch.setenv('ORBX_DESC_RING', str(ring))
    imgs = [synth_image(720, 540, 8200 + c) for c in range(2)]
    rig = ob.RigExtractor(720, 540, 2, 1500, fma_mode=fma)      # the knobs are read when the handle is created
    res = rig.extract(imgs, ob.make_params(2, 1500, 20, 7, (0, 720)))
    for b in range(2):
        compare_image(rig, b, imgs[b], 1500, 20, 7, (0, 720), res[b], fma_mode=fma, stages=(b == 0))


def test_concurrent_extractor_threads_reference_usage():
    """The reference calls operator() from 2-4 std::threads, one per extractor instance (src/Frame.cc:1841-1862): four handles,
    four host threads, repeated small calls (the cached CUDA-graph path with thread-local capture) - every result must equal the
    result of the same call made alone."""
    import threading
    _need_gpu()
    w, h, nf = 720, 540, 1500
    imgs = [synth_image(w, h, 3300 + c) for c in range(4)]
    params = ob.make_params(1, nf, 20, 7, (0, 0))
    alone = []
    for c in range(4):
        rig = ob.RigExtractor(w, h, 1, nf)
        alone.append(rig.extract([imgs[c]], params)[0])
    rigs = [ob.RigExtractor(w, h, 1, nf) for _ in range(4)]
    out = [[None] * 6 for _ in range(4)]
    errors = []

    def worker(c):
        try:
            for it in range(6):
                out[c][it] = rigs[c].extract([imgs[(c + it) % 4]], params)[0]
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    ts = [threading.Thread(target=worker, args=(c,)) for c in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for c in range(4):
        for it in range(6):
            mono, kps, desc = out[c][it]
            rmono, rkps, rdesc = alone[(c + it) % 4]
            assert mono == rmono and np.array_equal(desc, rdesc)
            for f in KPF:
                assert np.array_equal(kps[f].view(np.uint32), rkps[f].view(np.uint32)), f


def test_graph_cache_survives_many_device_buffers():
    """Small device-resident calls replay a CUDA graph cached per input buffer; a caller that keeps handing in new buffers (80
    distinct ones here, more than the cache holds) must keep getting the same results."""
    _need_gpu()
    import torch
    w, h = 512, 384
    imgs = np.stack([synth_image(w, h, 3500 + c) for c in range(2)])
    rig = ob.RigExtractor(w, h, 2, 500)
    p = ob.make_params(2, 500, 20, 7, (0, 0))
    ref = [(m, k.copy(), d.copy()) for m, k, d in rig.extract(list(imgs), p)]
    rig.set_stream(torch.cuda.current_stream().cuda_stream)
    base = torch.from_numpy(imgs).cuda()
    keep = []
    for it in range(80):
        t = base.clone(); keep.append(t)                  # a new device address every call
        rig.extract_device(t.data_ptr(), 2, p)
        if it % 9 == 0 or it == 79:
            for (m0, k0, d0), (m1, k1, d1) in zip(ref, rig.fetch(2)):
                assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)


@pytest.mark.parametrize('w,h,nf', [(1280, 1024, 2000), (752, 480, 1000)])
def test_first_handle_of_a_process_at_any_geometry(w, h, nf):
    """Shared-memory opt-in limits are per kernel and per PROCESS: a geometry whose tiles sit just under 48 KB must also work as the
    first handle of a fresh process (no earlier handle has raised the limit)."""
    import subprocess
    import sys
    _need_gpu()
    code = ('import numpy as np, openmavis_b200 as ob\n'
            'from openmavis_b200.synth import synth_image\n'
            'rig = ob.RigExtractor(%d, %d, 2, %d)\n'
            'res = rig.extract([synth_image(%d, %d, 5000), synth_image(%d, %d, 5001)], ob.make_params(2, %d, 20, 7, (0, 0)))\n'
            'print(len(res[0][1]), len(res[1][1]))\n' % (w, h, nf, w, h, w, h, nf))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-800:]
    a, b = (int(x) for x in out.stdout.split())
    assert a >= nf * 0.9 and b >= nf * 0.9


def _fuzz_case(seed):
    """A random but supported extractor configuration (src/ORBextractor.cc:351-420 accepts any of them) and a random image."""
    rng = np.random.default_rng(77000 + seed)
    nl = int(rng.integers(1, 9))
    sf = float(rng.choice([1.1, 1.2, 1.25, 1.5, 2.0]))
    smallest = 110                                     # the smallest level keeps a few cells in both directions
    while smallest * sf ** (nl - 1) > 800: nl -= 1     # (bounded image sizes: the CPU oracle runs every case)
    lo = int(np.ceil(smallest * sf ** (nl - 1)))
    w = int(rng.integers(max(lo, 160), max(lo, 160) + 700))
    h = int(rng.integers(max(lo, 120), max(lo, 120) + 500))
    w = max(w, int(0.65 * h) + 45)                      # (nIni = round(width / height) is 0 on no level: the reference divides by it)
    nf = int(rng.integers(30, 2500))
    f = 1.0 / sf
    n0 = nf * (1 - f) / (1 - f ** nl) if nl > 1 else nf   # features asked of level 0 (src/ORBextractor.cc:373-384)
    if n0 > 1900: nf = int(nf * 1900 / n0)              # (the product's quadtree workspace holds about 2000 nodes per level)
    ini = int(rng.integers(3, 70)); mn = int(rng.integers(1, ini + 1))
    kind = rng.choice(['synth', 'synth', 'synth', 'noise', 'checker', 'halfflat', 'lowcontrast'])
    if kind == 'noise':
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    elif kind == 'checker':
        p = int(rng.integers(3, 12)); yy, xx = np.mgrid[0:h, 0:w]
        img = ((((yy // p) + (xx // p)) % 2) * int(rng.integers(10, 200)) + 20).astype(np.uint8)
    else:
        img = synth_image(w, h, 5000 + seed)
        if kind == 'halfflat': img[:, : w // 2] = 90
        if kind == 'lowcontrast': img = (img // 6 + 100).astype(np.uint8)
    a = int(rng.integers(0, w)); b = int(rng.integers(0, w))
    lap = (0, 0) if rng.random() < 0.4 else (min(a, b), max(a, b))
    return w, h, nl, sf, nf, ini, mn, img, lap, kind


@pytest.mark.parametrize('seed', range(int(os.environ.get('ORBX_FUZZ_N', '48'))))
def test_random_configurations_against_the_oracle(seed):
    """Differential test over random geometries (level counts, scale factors, image sizes that are multiples of nothing), feature
    budgets, FAST thresholds, lapping areas and image kinds: pyramid planes incl. frame, FAST candidates per level (position,
    order, response) and the final keypoints / descriptors / return value, bit-exact against the CPU oracle and against the compiled reference
    (ORBX_FUZZ_N widens the sweep: 800 configurations were run green on the B200 in round 2)."""
    _need_gpu()
    w, h, nl, sf, nf, ini, mn, img, lap, kind = _fuzz_case(seed)
    rig = ob.RigExtractor(w, h, 1, nf, sf, nl)
    res = rig.extract([img], ob.make_params(1, nf, ini, mn, lap))
    E = O.OracleExtractor(nf, sf, nl, ini, mn)
    mono, kps, desc = E(img, lap)
    ctx = (seed, w, h, nl, sf, nf, ini, mn, lap, kind)
    for l in range(nl):
        assert np.array_equal(E.level(l), rig.pyramid_level(0, l, with_border=True)), ('pyramid', l) + ctx
        xs, ys, sc = E.level_candidates(l)
        gx, gy, gs = rig.debug_candidates(0, l)
        assert np.array_equal(xs, gx) and np.array_equal(ys, gy) and np.array_equal(sc, gs), ('candidates', l) + ctx
    gmono, gk, gd = res[0]
    assert gmono == mono and len(gk) == len(kps), ('counts', gmono, mono, len(gk), len(kps)) + ctx
    for f in KPF:
        assert np.array_equal(kps[f].view(np.uint32), gk[f].view(np.uint32)), ('field', f) + ctx
    assert np.array_equal(desc, gd), ('descriptors',) + ctx
    # and against the reference's own object code (src/ORBextractor.cc compiled unmodified, oracle/_ref)
    rmono, rk, rd = O.RefExtractor(nf, sf, nl, ini, mn, fma=False)(img, lap)
    assert gmono == rmono and len(gk) == len(rk), ('reference counts',) + ctx
    for f in KPF:
        assert np.array_equal(rk[f].view(np.uint32), gk[f].view(np.uint32)), ('reference field', f) + ctx
    assert np.array_equal(rd, gd), ('reference descriptors',) + ctx
