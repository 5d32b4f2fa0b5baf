"""Drop-in test of the matcher glue (openmavis_b200/csrc/host/ORBmatcher_glue.cc): ORB_SLAM3::ORBmatcher methods with the REFERENCE's
signatures, compiled against the reference's own include/ORBmatcher.h and stub Frame / KeyFrame / MapPoint objects, running their
searches on the GPU through liborbx.so — against the reference's own object code (oracle/_ref/libref_matcher.so = src/ORBmatcher.cc
compiled unmodified).  Both libraries are driven by the SAME C driver (oracle/ref_matcher_driver.cpp) with the same arrays, so what is
compared is exactly "swap the method bodies, keep everything else"."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import openmavis_b200 as ob
import test_ref_matcher as R
from openmavis_b200.synth import synth_descriptors

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = R.p
TRIALS_X = int(os.environ.get('ORBX_GLUE_TRIALS_X', '1'))   # multiplies the random scenarios per test (a wide sweep before a release)
_glue = None


def G():
    global _glue
    if ob.device_count() == 0:
        pytest.fail('no CUDA device: -m gpu tests must run on the GPU box')
    if _glue is None:
        so = os.path.join(ROOT, 'tests', 'host_harness', 'libglue_matcher.so')
        if not os.path.exists(so):
            raise RuntimeError('%s is missing (built by tests/host_harness/Makefile where /root/reference is mounted)' % so)
        ob.lib()
        _glue = C.CDLL(so)
    return _glue


def test_descriptor_distance_host_side():
    a = synth_descriptors(500, 1); b = synth_descriptors(500, 2)
    for i in range(500):
        assert G().refm_descriptor_distance(p(a[i]), p(b[i])) == R.RM().refm_descriptor_distance(p(a[i]), p(b[i]))


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_projection_frame_mappoints(ncam):
    rng = np.random.default_rng(200 + ncam)
    for trial in range(6 * TRIALS_X):
        S = R.sbp_scenario(rng, ncam, int(rng.integers(200, 1500)), crowded=trial % 2 == 1)
        th = float([1.0, 3.0, 6.0][trial % 3]); far = trial % 4 == 3
        N = int(S['cam_off'][-1])
        outs = []
        for L in (R.RM(), G()):
            a = np.zeros(N, np.int32)
            n = L.refm_search_by_projection(ncam, p(S['cam_off']), p(S['k']), p(S['desc']), p(S['bounds']), p(S['sf']), 8, p(S['occupied']),
                                            p(S['l2r']), p(S['r2l']), p(S['u_right']), S['nmp'], p(S['in_view']), p(S['px']), p(S['py']), p(S['lvl']),
                                            p(S['vc']), p(S['pxr']), p(S['depth']), p(S['bad']), p(S['has_obs']), p(S['md']), C.c_float(th), int(far),
                                            C.c_float(40.0), C.c_float(0.8), p(a))
            outs.append((n, a))
        assert outs[0][0] == outs[1][0] and outs[0][0] > 20, (trial, outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1], outs[1][1]), (trial, int((outs[0][1] != outs[1][1]).sum()))


def test_search_for_initialization():
    rng = np.random.default_rng(18)
    for trial in range(8 * TRIALS_X):
        n1, n2 = int(rng.integers(100, 1500)), int(rng.integers(100, 1500))
        k2 = R.make_kps(n2, rng, lattice=6 if trial % 2 else None); k2['octave'][rng.random(n2) < 0.6] = 0
        d2 = synth_descriptors(n2, 1900 + trial)
        src = rng.integers(0, n2, n1)
        k1 = k2[src].copy(); k1['x'] += rng.uniform(-6, 6, n1).astype(np.float32); k1['y'] += rng.uniform(-6, 6, n1).astype(np.float32)
        k1['angle'] = (k1['angle'] + rng.choice([0, 0, 0, 45, 200], n1)) % 360
        d1 = d2[src] ^ np.packbits(rng.random((n1, 256)) < 0.06, axis=1, bitorder='little')
        bounds = np.array([0, 640, 0, 480], np.float32)
        prev = np.stack([k1['x'], k1['y']], 1).astype(np.float32)
        for chk in (0, 1):
            outs = []
            for L in (R.RM(), G()):
                pa = prev.copy(); a = np.zeros(n1, np.int32)
                n = L.refm_search_for_initialization(p(k1), p(d1), n1, p(k2), p(d2), n2, p(bounds), p(pa), 10 + 30 * (trial % 2), C.c_float(0.9), chk, p(a))
                outs.append((n, a, pa))
            assert outs[0][0] == outs[1][0], (trial, chk)
            assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2]), (trial, chk)


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_bow_keyframe_frame(ncam):
    rng = np.random.default_rng(300 + ncam)
    for trial in range(6 * TRIALS_X):
        nF = int(rng.integers(300, 1500)); nKF = int(rng.integers(300, 1500))
        dF = synth_descriptors(nF, 2900 + trial)
        dF[5:40] = dF[5]                                     # equal distances
        src = rng.integers(0, nF, nKF)
        flips = np.packbits(rng.random((nKF, 256)) < 0.05, axis=1, bitorder='little'); flips[:, 0] &= 0xE0
        dKF = dF[src] ^ flips
        dKF[::7] = synth_descriptors(len(dKF[::7]), 77 + trial)
        aF = rng.uniform(0, 360, nF).astype(np.float32)
        aKF = ((aF[src] + rng.choice([0, 0, 0, 30, 190], nKF)) % 360).astype(np.float32)
        fvK, fvF = R._feature_vector(dKF), R._feature_vector(dF)
        skip = (rng.random(nKF) < 0.25).astype(np.uint8)

        def split(n):
            cuts = np.sort(rng.integers(0, n + 1, ncam - 1)) if ncam > 1 else np.zeros(0, np.int64)
            return np.concatenate([[0], cuts, [n]]).astype(np.int32)
        offK, offF = split(nKF), split(nF)
        for chk in (0, 1):
            outs = []
            for L in (R.RM(), G()):
                m = np.zeros(nF, np.int32)
                n = L.refm_search_by_bow_frame(len(fvK[0]), p(fvK[0]), p(fvK[1]), p(fvK[2]), len(fvF[0]), p(fvF[0]), p(fvF[1]), p(fvF[2]), p(dKF), nKF,
                                               p(dF), nF, p(skip), p(aKF), p(aF), ncam, p(offK), p(offF), C.c_float(0.7), chk, p(m))
                outs.append((n, m))
            assert outs[0][0] == outs[1][0], (trial, chk, outs[0][0], outs[1][0])
            assert np.array_equal(outs[0][1], outs[1][1]), (trial, chk, int((outs[0][1] != outs[1][1]).sum()))


def test_search_by_bow_keyframe_keyframe():
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), src/ORBmatcher.cc:1006-1129, through the glue."""
    for S in R.bow_keyframes_scenarios():
        for chk in (0, 1):
            na, a = R.call_bow_keyframes(R.RM(), S, chk)
            nb, b = R.call_bow_keyframes(G(), S, chk)
            assert na == nb and na > 10, (S[0], chk, na, nb)
            assert np.array_equal(a, b), (S[0], chk, int((a != b).sum()))


@pytest.mark.parametrize('ncam', [1, 2, 4])
def test_search_by_projection_current_last(ncam):
    """The tracking matcher SearchByProjection(CurrentFrame, LastFrame, th, bMono), src/ORBmatcher.cc:1985-2413: the glue computes the
    projections on the host with the reference's expressions (poses, camera->project, Trl) and runs the search on the device."""
    seen = [0, 0]
    for S in R.sbp_last_scenarios(ncam):
        na, a = R.call_sbp_last(R.RM(), S)
        nb, b = R.call_sbp_last(G(), S)
        assert na == nb and na > 20, (S['trial'], na, nb)
        # (-3 = a keypoint that still holds the unobserved point it held: untouched, as opposed to claimed and cleared by the rotation filter,
        #  which the reference sets to NULL — the glue tells them apart through orbx_search_by_projection_last_cleared)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))
        seen[0] += int((a == -3).sum()); seen[1] += int(((a == -1) & (S['occupied'] == 2)).sum())
    assert seen[0] > 20 and seen[1] > 0


def test_search_by_projection_relocalisation():
    """SearchByProjection(CurrentFrame, KeyFrame, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2415-2535, through the glue."""
    for S in R.reloc_scenarios():
        na, a = R.call_reloc(R.RM(), S)
        nb, b = R.call_reloc(G(), S)
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))


@pytest.mark.parametrize('with_kfs', [0, 1])
def test_search_by_projection_keyframe_sim3(with_kfs):
    """Both loop-closing overloads SearchByProjection(KeyFrame*, Sim3f&, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming,
    cameraID), src/ORBmatcher.cc:668-893, through the glue (incl. the vpMatchedKF bookkeeping of the second one)."""
    for S in R.kf_sim3_scenarios(with_kfs):
        na, a, oka = R.call_kf_sim3(R.RM(), S)
        nb, b, okb = R.call_kf_sim3(G(), S)
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))
        assert oka == okb and (not with_kfs or okb == 1)


def test_fuse_keyframe_mappoints():
    """Fuse(KeyFrame*, vpMapPoints, th, cameraID), src/ORBmatcher.cc:1458-1647, through the glue: one device call for all searches, then the
    graph updates in point order — Replace in both directions, AddObservation / AddMapPoint, bad points, a map point that is listed twice."""
    for S in R.fuse_graph_scenarios():
        ra = R.call_fuse_graph(R.RM(), S)
        rb = R.call_fuse_graph(G(), S)
        assert ra[0] == rb[0] and ra[0] > 20, (S['trial'], ra[0], rb[0])
        for x, y, name in zip(ra[1:], rb[1:], ('q_repl', 'kf_mp', 'holder_repl')):
            assert np.array_equal(x, y), (S['trial'], name, int((x != y).sum()))


def test_search_by_sim3():
    """SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), src/ORBmatcher.cc:1771-1983, through the glue: both projection loops on the host,
    both window searches and the agreement check on the device."""
    for S in R.sim3_scenarios():
        na, a = R.call_sim3(R.RM(), S)
        nb, b = R.call_sim3(G(), S)
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))


def test_fuse_keyframe_sim3():
    """Fuse(KeyFrame*, Sim3f&, vpPoints, th, vpReplacePoint), src/ORBmatcher.cc:1648-1769, through the glue: one device search per camera of
    the keyframe, then vpReplacePoint / AddObservation / AddMapPoint in point order."""
    for S in R.fuse_sim3_scenarios():
        na, a = R.call_fuse_sim3(R.RM(), S)
        nb, b = R.call_fuse_sim3(G(), S)
        assert na == nb and na > 20, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))


@pytest.mark.parametrize('ncam', [1, 2, 4])
@pytest.mark.parametrize('coarse', [0, 1])
def test_search_for_triangulation_pinhole(ncam, coarse):
    """SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), src/ORBmatcher.cc:1131-1456, through the glue, with real
    pinhole rigs (1 / 2 / 4 cameras): the glue derives the epipole and the eleven camera pairs from the poses with the reference's
    expressions and hands the device one fundamental matrix per pair; the reference side runs its own Pinhole::epipolarConstrain
    (src/CameraModels/Pinhole.cpp:103-132) per candidate.  Descriptor gate, epipole reject, carried pair state, last minimum, rotation filter."""
    tot = 0
    for S in R.sft_pinhole_scenarios(ncam, coarse):
        na, a = R.call_sft_pinhole(R.RM(), S)
        nb, b = R.call_sft_pinhole(G(), S)
        ub = -1 if coarse else R.sft_row_on_uninitialised_pair(S)
        if ub >= 0 and S['chk']:
            continue                     # (with the rotation filter a different match in that row can shift the histogram for every row)
        if ub >= 0 and a[ub] != b[ub]:   # the one row the reference decides on uninitialised R12 / t12
            na -= int(a[ub] >= 0); nb -= int(b[ub] >= 0); a[ub] = b[ub] = -1
        assert na == nb, (S['trial'], na, nb)
        assert np.array_equal(a, b), (S['trial'], int((a != b).sum()))
        tot += na
    assert tot > 100
