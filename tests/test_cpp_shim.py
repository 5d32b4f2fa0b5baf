"""The C++ drop-in classes (openmavis_b200/csrc/host: ORB_SLAM3::ORBextractor / ORBmatcher on the C ABI), driven the way
src/Frame.cc drives the reference classes, through tests/host_harness/libshim_harness.so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_image, synth_descriptors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def shim():
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'openmavis_b200', 'csrc')], stdout=subprocess.DEVNULL)
    d = os.path.join(ROOT, 'tests', 'host_harness')
    subprocess.check_call(['make', '-C', d], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(d, 'libshim_harness.so'))
    L.shim_extractor_new.restype = C.c_void_p
    L.shim_extractor_new.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.shim_extractor_delete.argtypes = [C.c_void_p]
    L.shim_extractor_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.shim_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.shim_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def test_shim_getters_match_reference_tables(shim):
    # constructor arithmetic (src/ORBextractor.cc:362-376) — no CUDA needed
    e = shim.shim_extractor_new(1000, 1.2, 8, 20, 7)
    a = [np.zeros(8, np.float32) for _ in range(4)]
    shim.shim_extractor_tables(e, *[O.p(x) for x in a])
    t = O.OracleExtractor(1000, 1.2, 8, 20, 7).tables()
    for got, key in zip(a, ('sf', 'inv_sf', 'sigma2', 'inv_sigma2')):
        assert np.array_equal(got.view(np.uint32), t[key].view(np.uint32))
    shim.shim_extractor_delete(e)


def test_shim_empty_image_returns_minus_one(shim):
    e = shim.shim_extractor_new(1000, 1.2, 8, 20, 7)
    n = C.c_int(7)
    assert shim.shim_extract(e, None, 0, 0, 0, 0, 0, None, None, 0, C.byref(n)) == -1    # src/ORBextractor.cc:994-995
    shim.shim_extractor_delete(e)


@pytest.mark.gpu
@pytest.mark.parametrize('lap', [(0, 0), (0, 1000), (200, 500)])
def test_shim_operator_matches_oracle(shim, lap):
    img = synth_image(752, 480, 77)
    padded = np.zeros((480, 800), np.uint8); padded[:, :752] = img       # non-continuous cv::Mat (row stride 800)
    e = shim.shim_extractor_new(1000, 1.2, 8, 20, 7)
    shim.shim_set_pyramid_download(C.c_void_p(e), 1)                      # off by default (only Frame::ComputeStereoMatches reads them)
    cap = 1200
    kps = np.zeros(cap, O.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
    mono = shim.shim_extract(e, O.p(padded), 752, 480, 800, lap[0], lap[1], O.p(kps), O.p(desc), cap, C.byref(n))
    rm, rk, rd = O.OracleExtractor(1000, 1.2, 8, 20, 7)(img, lap)
    assert mono == rm and n.value == len(rk)
    assert np.array_equal(kps[:n.value], rk) and np.array_equal(desc[:n.value], rd)
    E = O.OracleExtractor(1000, 1.2, 8, 20, 7); E(img, lap)
    for l in (0, 3, 7):                                                   # mvImagePyramid host mirrors
        w, h = C.c_int(), C.c_int()
        assert shim.shim_pyramid_level(e, l, None, C.byref(w), C.byref(h)) == 1
        buf = np.zeros((h.value, w.value), np.uint8)
        shim.shim_pyramid_level(e, l, O.p(buf), C.byref(w), C.byref(h))
        assert np.array_equal(buf, E.level(l)[19:-19, 19:-19])
    shim.shim_extractor_delete(e)


@pytest.mark.gpu
def test_shim_matcher_matches_oracle(shim):
    t = synth_descriptors(500, 1); q = synth_descriptors(300, 2); q[:100] = t[:100]; q[:100, 0] ^= 3
    assert shim.shim_descriptor_distance(O.p(q[5]), O.p(t[9])) == O.hamming(q[5], t[9])
    idx = np.zeros((300, 2), np.int32); dist = np.zeros((300, 2), np.int32); keep = np.zeros(300, np.uint8)
    consts = shim.shim_knn_ratio(O.p(q), 300, O.p(t), 500, C.c_double(0.8), O.p(idx), O.p(dist), O.p(keep))
    assert consts == 50 * 1000 + 100
    ridx, rdist = O.knn2(q, t)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    assert np.array_equal(keep.astype(bool), O.ratio_test(rdist, 0.8))


@pytest.mark.gpu
def test_shim_search_for_initialization_matches_oracle(shim):
    """The shim class method (std::vector<cv::KeyPoint>, cv::Point2f) against the oracle restatement of src/ORBmatcher.cc:895-1004."""
    import openmavis_b200 as ob
    from openmavis_b200.synth import synth_image
    rng = np.random.default_rng(5)
    rig = ob.RigExtractor(512, 512, 1, 1200)
    res = rig.extract([synth_image(512, 512, 6100)], ob.make_params(1, 1200, 20, 7, (0, 511)))
    k2 = res[0][1].copy(); d2 = res[0][2].copy()
    n2 = k2.size; n1 = 900
    src = rng.integers(0, n2, n1)
    k1 = k2[src].copy()
    k1['x'] += rng.uniform(-5, 5, n1).astype(np.float32); k1['y'] += rng.uniform(-5, 5, n1).astype(np.float32)
    d1 = d2[src] ^ np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little')
    prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1).astype(np.float32)); prev_ref = prev.copy()
    m12 = np.zeros(n1, np.int32); ref = np.zeros(n1, np.int32)
    nm = shim.shim_search_for_initialization(O.p(k1), O.p(d1), n1, O.p(k2), O.p(d2), n2, O.p(prev), 20, C.c_float(0.9), 1, O.p(m12))
    L = O.lib()
    g2 = L.orc_grid_create(O.p(k2), n2, C.c_float(0), C.c_float(512), C.c_float(0), C.c_float(512))
    rn = L.orc_search_for_initialization(O.p(k1), O.p(d1), n1, C.c_void_p(g2), O.p(k2), O.p(d2), n2, O.p(prev_ref), 20, C.c_float(0.9), 1, O.p(ref))
    L.orc_grid_destroy(C.c_void_p(g2))
    assert nm == rn and np.array_equal(m12, ref) and np.array_equal(prev, prev_ref) and nm > 50


@pytest.mark.gpu
def test_shim_rig_rendezvous_four_camera_threads(shim):
    """src/Frame.cc:1841-1862: four extractor objects (500 / 500 / 1000 / 1000 features as src/Tracking.cc:583-587 creates them), one
    std::thread per camera per rig frame.  Every camera's result must equal the reference's, and after the first frames the four
    calls of a rig frame must travel as ONE device batch (the rendezvous learns how many calls arrive together)."""
    w, h, ncam, frames = 720, 540, 4, 12
    nf = np.array([500, 500, 1000, 1000], np.int32)
    laps = np.array([[0, 720], [0, 720], [0, 0], [0, 0]], np.int32)
    imgs = np.stack([synth_image(w, h, 3400 + (f % 3) * 10 + c) for f in range(frames) for c in range(ncam)])
    cap = 1200
    kps = np.zeros((ncam, cap), O.KP_DTYPE); desc = np.zeros((ncam, cap, 32), np.uint8)
    n = np.zeros(ncam, np.int32); mono = np.zeros(ncam, np.int32)
    b0, i0, b1, i1 = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
    shim.shim_rig_stats(C.byref(b0), C.byref(i0))
    shim.shim_extract_rig(ncam, frames, O.p(imgs), w, h, O.p(nf), O.p(laps), 15, 7, O.p(kps), O.p(desc), cap, O.p(n), O.p(mono))
    shim.shim_rig_stats(C.byref(b1), C.byref(i1))
    last = imgs[(frames - 1) * ncam:]
    for c in range(ncam):
        rm, rk, rd = O.RefExtractor(int(nf[c]), 1.2, 8, 15, 7)(last[c], tuple(laps[c]))
        assert mono[c] == rm and n[c] == len(rk), (c, mono[c], rm, n[c], len(rk))
        assert np.array_equal(kps[c, :n[c]], rk) and np.array_equal(desc[c, :n[c]], rd), c
    images, batches = i1.value - i0.value, b1.value - b0.value
    assert images == ncam * frames
    assert batches <= frames + 6, (batches, images)          # one batch per rig frame once learned (the first frame may take up to 4)
