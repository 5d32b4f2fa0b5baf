"""CPU checks of the product's host/device-shared cores against the oracle (same source as the kernels,
executed by one host thread through tests/host_harness)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from openmavis_b200.synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def hc():
    d = os.path.join(ROOT, 'tests', 'host_harness')
    subprocess.check_call(['make', '-C', d], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(d, 'libhost_core.so'))


def test_std_sort_emulation_matches_libstdcxx(hc):
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n = int(rng.integers(0, 600)) if trial % 10 else int(rng.integers(0, 40))
        sizes = rng.integers(2, 2 + int(rng.integers(1, 12)), size=n).astype(np.int32)      # tie-heavy
        ulx = (rng.integers(0, int(rng.integers(1, 40)), size=n) * 17).astype(np.int32)
        if trial % 7 == 0:
            sizes = np.sort(sizes)[::-1].copy()
        ref = O.sort_nodes(sizes, ulx)
        got = np.zeros(n, np.int32)
        hc.hc_std_sort(O.p(sizes), O.p(ulx), n, O.p(got))
        assert np.array_equal(ref, got), trial
        got2 = np.zeros(n, np.int32)
        hc.hc_std_sort_block(O.p(sizes), O.p(ulx), n, O.p(got2))       # block-parallel formulation used by the kernel
        assert np.array_equal(ref, got2), trial


def _distribute_both(hc, xs, ys, sc, minX, maxX, minY, maxY, N):
    ref = O.distribute(xs, ys, sc, minX, maxX, minY, maxY, N)
    got = np.zeros(max(N + 8, 64), np.int32)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); sc = np.ascontiguousarray(sc, np.int32)
    k = hc.hc_distribute(O.p(xs), O.p(ys), O.p(sc), xs.size, minX, maxX, minY, maxY, N, O.p(got), got.size)
    return ref, got[:k]


def test_octree_random_sets(hc):
    rng = np.random.default_rng(1)
    for trial in range(1500):
        w = int(rng.integers(60, 1300)); h = int(rng.integers(max(40, w // 2 + 1), max(42, int(w / 0.7))))
        if round(np.float32(w) / np.float32(h)) < 1:
            continue
        n = int(rng.integers(1, 4000)) if trial % 5 else int(rng.integers(1, 30))
        N = int(rng.integers(1, 500))
        if trial % 3 == 0:    # clustered + duplicates
            cx, cy = rng.integers(0, w), rng.integers(0, h)
            xs = np.clip(rng.normal(cx, w / 8, n), 0, w - 1).astype(np.int32)
            ys = np.clip(rng.normal(cy, h / 8, n), 0, h - 1).astype(np.int32)
        else:
            xs = rng.integers(0, w, n).astype(np.int32); ys = rng.integers(0, h, n).astype(np.int32)
        sc = rng.integers(7, 60, n).astype(np.int32)
        ref, got = _distribute_both(hc, xs, ys, sc, 16, 16 + w, 16, 16 + h, N)
        assert np.array_equal(ref, got), (trial, w, h, n, N)


@pytest.mark.parametrize('w,h,nf', [(752, 480, 1000), (720, 540, 1500), (512, 512, 1500)])
def test_octree_on_real_fast_candidates(hc, w, h, nf):
    img = synth_image(w, h, 77 + w)
    E = O.OracleExtractor(nf, 1.2, 8, 20, 7)
    E(img)
    t = E.tables()
    for l in range(8):
        xs, ys, sc = E.level_candidates(l)
        if xs.size == 0:
            continue
        lw = int(np.rint(np.float32(w) * t['inv_sf'][l])); lh = int(np.rint(np.float32(h) * t['inv_sf'][l]))
        ref, got = _distribute_both(hc, xs, ys, sc, 16, lw - 16, 16, lh - 16, int(t['nfeat'][l]))
        assert np.array_equal(ref, got), l
        assert len(got) <= max(int(t['nfeat'][l]) + 2, 8)


def test_numa_binding_helper_is_harmless_without_topology():
    """openmavis_b200.shard.bind_to_gpu_numa: unknown device / unreadable sysfs -> None and the affinity is left alone."""
    import os
    from openmavis_b200.shard import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa('ffff:ff:1f.0') is None
    assert os.sched_getaffinity(0) == before
