"""GPU parity of the fisheye stereo post-filter against the CPU oracle:
KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:319-388) and Frame::ComputeStereoFishEyeMatches /
ComputeMultiFishEyeMatches (src/Frame.cc:1409-1525).

Float geometry: integer outputs (match tables, counts) must be identical wherever the oracle's decision is not within 1e-3
(relative) of a gate; depths and 3-D points agree to TOL = 1e-4 relative (the only differences are CUDA's atan2f / tanf /
cosf / sinf, <= 2 ulp from glibc's; the DLT null vector is computed in fp64 on both sides)."""
import numpy as np
import pytest

import openmavis_b200 as ob
from openmavis_b200.synth import synth_fisheye_pairs
from test_oracle_fisheye import fisheye_frame, make_rig, oracle_fisheye_matches, oracle_triangulate

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope='module')
def M():
    if ob.device_count() == 0:
        pytest.fail('no CUDA device: -m gpu tests must run on the GPU box')
    return ob.ORBmatcher(0.8)


def rig_of(s):
    return ob.kb8_stereo(s['cam1'], s['cam2'], s['R12'], s['t12'])


@pytest.mark.parametrize('n,seed', [(1, 1), (33, 2), (5000, 3), (20000, 4)])
def test_triangulate_matches(M, n, seed):
    s = synth_fisheye_pairs(n, seed)
    sig1 = (np.float32(1.2) ** (2 * s['octave1'])).astype(np.float32); sig2 = (np.float32(1.2) ** (2 * s['octave2'])).astype(np.float32)
    ref_d, ref_p, margin = oracle_triangulate(make_rig(s), s['xy1'], s['xy2'], sig1, sig2)
    d, p3 = M.TriangulateMatches(rig_of(s), s['xy1'], s['xy2'], sig1, sig2)
    assert M.last_launches() == 1
    clear = margin > 1e-3
    if n >= 5000:
        assert clear.mean() > 0.97
        assert all(((ref_d == c) & clear).any() for c in (-1, -2, -4))      # reject paths covered
    rej = ref_d <= 0
    assert np.array_equal(d[clear & rej], ref_d[clear & rej])              # same reject code
    ok = clear & ~rej
    assert (d[ok] > 0).all()
    assert np.allclose(d[ok], ref_d[ok], rtol=TOL, atol=0)
    assert np.allclose(p3[ok], ref_p[ok], rtol=TOL, atol=TOL * 1e-2)
    assert (p3[d <= 0] == 0).all()


def test_triangulate_matches_empty(M):
    s = synth_fisheye_pairs(4, 9)
    d, p3 = M.TriangulateMatches(rig_of(s), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32))
    assert d.size == 0 and p3.shape == (0, 3)


@pytest.mark.parametrize('ratio,seed,shape', [(0.7, 40, (900, 950, 120, 140)), (0.8, 41, (1500, 1500, 0, 0)), (0.8, 42, (300, 280, 299, 10)),
                                              (0.7, 43, (64, 70, 0, 68))])
def test_stereo_fisheye_matches(M, ratio, seed, shape):
    f = fisheye_frame(seed, *shape)
    o = oracle_fisheye_matches(f, ratio)
    s = f['s']
    g = M.ComputeStereoFishEyeMatches(rig_of(s), f['kl'], f['dl'], f['mono_left'], f['kr'], f['dr'], f['mono_right'], f['sigma2'], ratio)
    assert g['descMatches'] == o['descMatches']                            # 2-NN + ratio test: integer, exact
    near_gate = o['margin'] <= 1e-3                                         # (1e30 for queries that never reached the triangulation)
    clear = ~near_gate
    assert np.array_equal(g['mvLeftToRightMatch'][clear], o['l2r'][clear])
    if not near_gate.any():
        assert g['nMatches'] == o['nMatches']
        assert np.array_equal(g['mvRightToLeftMatch'], o['r2l'])
    m = clear & (o['l2r'] >= 0)
    assert np.allclose(g['mvDepth'][m], o['depth'][m], rtol=TOL, atol=0)
    assert np.allclose(g['mvStereo3Dpoints'][m], o['p3d'][m], rtol=TOL, atol=TOL * 1e-2)
    un = clear & (o['l2r'] < 0)
    assert (g['mvDepth'][un] == -1).all() and (g['mvStereo3Dpoints'][un] == 0).all()
    if shape[0] >= 900:
        assert o['nMatches'] > 150


def test_stereo_fisheye_matches_degenerate(M):
    f = fisheye_frame(50, 40, 41, 40, 3)                                   # empty left lapping block
    g = M.ComputeStereoFishEyeMatches(rig_of(f['s']), f['kl'], f['dl'], 40, f['kr'], f['dr'], 3, f['sigma2'], 0.7)
    assert g['nMatches'] == 0 and (g['mvLeftToRightMatch'] == -1).all() and (g['mvRightToLeftMatch'] == -1).all()
    f = fisheye_frame(51, 40, 41, 0, 40)                                   # one right row: knnMatch yields a single neighbour
    g = M.ComputeStereoFishEyeMatches(rig_of(f['s']), f['kl'], f['dl'], 0, f['kr'], f['dr'], 40, f['sigma2'], 0.7)
    o = oracle_fisheye_matches(f, 0.7)
    assert g['nMatches'] == 0 == o['nMatches'] and (g['mvDepth'] == -1).all()
    with pytest.raises(ob.OrbxError):
        bad = f['kl'].copy(); bad['octave'][5] = 9
        M.ComputeStereoFishEyeMatches(rig_of(f['s']), bad, f['dl'], 0, f['kr'], f['dr'], 0, f['sigma2'], 0.7)


def test_device_resident_chain_equals_host_call(M):
    """Extraction -> stereo matching without the descriptors leaving the GPU (orbx_stereo_fisheye_matches_device on the rows the
    extraction left on the device) gives exactly what the host-pointer call gives on the fetched rows."""
    from openmavis_b200.synth import synth_image
    w = h = 512
    rig_x = ob.RigExtractor(w, h, 2, 1500)
    left = synth_image(w, h, 7100)
    xs = np.arange(w)
    right = left[:, np.clip(xs + (9 + 6 * np.sin(xs / 70.0)).astype(np.int64), 0, w - 1)].copy()
    res = rig_x.extract([left, right], ob.make_params(2, 1500, 20, 7, (0, w - 1)))       # lapping area = the whole image
    (ml, kl, dl), (mr, kr, dr) = res[0][:3], res[1][:3]
    s = synth_fisheye_pairs(4, 3)
    rig = ob.kb8_stereo(s['cam1'], s['cam1'], np.eye(3, dtype=np.float32), np.array([0.1, 0, 0], np.float32))
    g_dev = rig_x.compute_stereo_fisheye_matches(M, rig, 0, 1, (kl.size, ml), (kr.size, mr), ratio=0.8)
    g_host = M.ComputeStereoFishEyeMatches(rig, kl, dl, ml, kr, dr, mr, rig_x.sigma2, 0.8)
    assert g_dev['descMatches'] == g_host['descMatches'] > 100
    for k in ('mvLeftToRightMatch', 'mvRightToLeftMatch', 'mvDepth', 'mvStereo3Dpoints'):
        assert np.array_equal(g_dev[k], g_host[k]), k
    assert g_dev['nMatches'] == g_host['nMatches']
