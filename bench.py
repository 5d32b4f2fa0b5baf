#!/usr/bin/env python
"""bench.py — headline benchmark of the ORB-extraction / Hamming-matching hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config 3 of BASELINE.json, the one the metric is quoted on): Hilti-2022 rig frames, 5 x 720x540 cameras,
1500 features/camera, FAST 20/7, 8 levels, scale 1.2, lapping {0,720} on the two front cameras; a "step" extracts a
batch of --frames rig frames per GPU (default 64 = 320 images; the pyramids of one batch are ~460 MB, larger than the
126 MB L2).  Weak scaling: every rank owns its own batch of rig frames, no data-path collective.

One JSON line is printed by rank 0:
  value     rig-frames/s, inputs resident in HBM, device pipeline only (CUDA events, max over ranks)
  e2e       rig-frames/s through the host-buffer C-ABI call orbx_extract_batch (pinned host images in, keypoints +
            descriptors out, copies inside the timed region)
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
            + e2e.copy_roofline: the same pinned buffers moved by plain cudaMemcpyAsync alone (what the host link allows)
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  sustained the device pipeline repeated for >= 2 s with its clock record
  cpu_baseline  the reference's own src/ORBextractor.cc (compiled unmodified, oracle/_ref/libref_orb_fma.so) on the host cores:
            all host threads (value), one std::thread per camera per rig frame as src/Frame.cc:1841-1862 does, the oracle port,
            and a cv2-primitive sanity line
  secondary device-resident extraction on the other BASELINE.json configurations (1, 2, 4, 5)
  matching  secondary metric: brute-force Hamming 2-NN pairs/s vs the measured popc issue rate; guided matchers with CPU timings
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NCAM, NFEAT, INI, MIN = 720, 540, 5, 1500, 20, 7
LAPS = [(0, 720), (0, 720), (0, 0), (0, 0), (0, 0)]
# algorithmic bytes per level-0 pixel: pyramid 7.1, FAST 3.1, blur 6.2  => 16.4 per image pixel
ALG_BYTES_PER_PX = {'pyramid': 7.12, 'fast': 3.095, 'blur': 6.19, 'total': 16.4}


# `config` is the same object in both arms (ours / --impl reference): the driver compares it key by key
CONFIG = {'workload': 'hilti_rig_5x720x540_nf1500',
          'baseline_config': 'BASELINE.json configs[2]: Hilti-2022 5-camera rig 5×720×540, 1500 features/camera, per-rig-frame batched extraction',
          'cameras': NCAM, 'image': [W, H], 'features_per_camera': NFEAT, 'fast_thresholds': [INI, MIN], 'levels': 8, 'scale_factor': 1.2,
          'lapping': LAPS, 'l2': 'GPU arm: the inputs + pyramids of one step (about 1 GB) exceed the 126 MB L2, no flush needed'}

_BATCH_CACHE = {}


def make_batch(frames, rank):
    if (frames, rank) not in _BATCH_CACHE:
        _BATCH_CACHE[(frames, rank)] = _make_batch(frames, rank)
    return _BATCH_CACHE[(frames, rank)]


def _make_batch(frames, rank):
    """frames x 5 synthetic camera images.  10 distinct textured images are generated (numpy, deterministic); the rest of
    the batch are shifted/flipped variants so that every image has different content at the same statistics."""
    from openmavis_b200.synth import synth_image
    base = [synth_image(W, H, 3000 + 10 * rank + i) for i in range(10)]
    rng = np.random.default_rng(rank)
    out = np.empty((frames * NCAM, H, W), np.uint8)
    for i in range(frames * NCAM):
        im = base[i % 10]
        if i >= 10:
            im = np.roll(im, (int(rng.integers(1, H)), int(rng.integers(1, W))), axis=(0, 1))
            if rng.random() < 0.5:
                im = im[:, ::-1]
        out[i] = im
    return out


class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active') and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_lib as O
    return O


def _front_side(frames, rank):
    imgs = make_batch(frames, rank)[: frames * NCAM].reshape(frames, NCAM, H, W)
    return np.ascontiguousarray(imgs[:, :2].reshape(-1, H, W)), np.ascontiguousarray(imgs[:, 2:].reshape(-1, H, W))


def cpu_reference_run(frames, threads, rank=0):
    """The reference's own ORBextractor (src/ORBextractor.cc compiled unmodified with the reference's flags: oracle/_ref/libref_orb_fma.so)
    on `frames` rig frames, `threads` workers each owning one extractor, images in parallel.  -> (rig-frames/s, seconds, keypoints)"""
    O = _oracle()
    front, side = _front_side(frames, rank)           # lapping differs per camera: two groups, same total work
    t_f, k_f = O.ref_extract_batch(front, NFEAT, 1.2, 8, INI, MIN, (0, 720), threads, fma=True)
    t_s, k_s = O.ref_extract_batch(side, NFEAT, 1.2, 8, INI, MIN, (0, 0), threads, fma=True)
    return frames / (t_f + t_s), t_f + t_s, k_f + k_s


def cpu_reference_thread_per_camera(frames, rank=0):
    """Reference-faithful threading (src/Frame.cc:1841-1862): rig frames one after the other, one std::thread per camera per frame."""
    O = _oracle()
    imgs = make_batch(frames, rank)[: frames * NCAM]
    dt, nk = O.ref_extract_rig_frames(imgs, NCAM, NFEAT, 1.2, 8, INI, MIN, LAPS, fma=True)
    return frames / dt, dt, nk


def cpu_oracle_run(frames, threads, rank=0):
    """Times the CPU oracle (port of the reference's ORBextractor) on `frames` rig frames with `threads` host threads."""
    O = _oracle()
    L = O.lib()
    front, side = _front_side(frames, rank)
    n = frames * NCAM
    nkp = np.zeros(n, np.int32); nm = np.zeros(n, np.int32)
    t0 = time.perf_counter()
    tot = L.orc_extract_batch(front.shape[0], O.p(front), W, H, NFEAT, O.C.c_float(1.2), 8, INI, MIN, 0, 720, 0, threads, O.p(nkp), O.p(nm))
    tot += L.orc_extract_batch(side.shape[0], O.p(side), W, H, NFEAT, O.C.c_float(1.2), 8, INI, MIN, 0, 0, 0, threads, O.p(nkp), O.p(nm))
    dt = time.perf_counter() - t0
    return frames / dt, dt, tot


def cv2_primitive_line(n_images=3):
    """Sanity line: the OpenCV primitives the reference calls (cv::resize, copyMakeBorder, FAST, GaussianBlur — SIMD builds in the cv2
    wheel), one OpenCV thread, driven from Python: per-image ms of each stage.  The per-cell FAST loop pays ~900 Python -> C calls per
    image; the whole-level FAST figure is the SIMD lower bound of the same pixels (not the reference's per-cell semantics)."""
    try:
        import cv2
    except Exception as ex:   # noqa: BLE001
        return {'unavailable': 'cv2 not importable: %s' % ex}
    cv2.setNumThreads(1)
    O = _oracle()
    inv = O.OracleExtractor(NFEAT, 1.2, 8, INI, MIN).tables()['inv_sf']
    imgs = make_batch(2, 0)[:n_images]
    t = {'pyramid': 0.0, 'fast_per_cell': 0.0, 'fast_whole_level': 0.0, 'blur': 0.0}
    det_i = cv2.FastFeatureDetector_create(INI, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_m = cv2.FastFeatureDetector_create(MIN, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for img in imgs:
        t0 = time.perf_counter()
        levels = [img]
        for l in range(1, 8):
            dw, dh = int(np.rint(np.float32(W) * inv[l])), int(np.rint(np.float32(H) * inv[l]))
            levels.append(cv2.resize(levels[-1], (dw, dh), interpolation=cv2.INTER_LINEAR))
        for L in levels:
            cv2.copyMakeBorder(L, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        t1 = time.perf_counter()
        for L in levels:                                  # src/ORBextractor.cc:709-794
            hh, ww = L.shape
            maxbx, maxby = ww - 16, hh - 16
            ncols, nrows = int((maxbx - 16) / 35.0), int((maxby - 16) / 35.0)
            wc, hc = int(np.ceil((maxbx - 16) / ncols)), int(np.ceil((maxby - 16) / nrows))
            for i in range(nrows):
                y0 = 16 + i * hc
                if y0 >= maxby - 3:
                    continue
                y1 = min(y0 + hc + 6, maxby)
                for j in range(ncols):
                    x0 = 16 + j * wc
                    if x0 >= maxbx - 6:
                        continue
                    cell = L[y0:y1, x0:min(x0 + wc + 6, maxbx)]
                    if not det_i.detect(cell):
                        det_m.detect(cell)
        t2 = time.perf_counter()
        for L in levels:
            det_i.detect(L)
        t3 = time.perf_counter()
        for L in levels:
            cv2.GaussianBlur(L, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        t4 = time.perf_counter()
        t['pyramid'] += t1 - t0; t['fast_per_cell'] += t2 - t1; t['fast_whole_level'] += t3 - t2; t['blur'] += t4 - t3
    out = {k + '_ms_per_image': 1e3 * v / len(imgs) for k, v in t.items()}
    out.update({'opencv': cv2.__version__, 'threads': 1, 'images': len(imgs)})
    return out


def guided_matchers(ob, device):
    """Secondary numbers for the small matchers (report queries/s and latency): config-4 style SearchByProjection
    (2 x 512x512, 1500 features / camera, 5000 map points), ComputeStereoMatches on a 752x480 pair, ComputeDistinctiveDescriptors on
    2000 map points, ComputeStereoFishEyeMatches on 2 x 1500 keypoints.  Wall-clock per call through the host-buffer C ABI (uploads and downloads included), median of 7."""
    from openmavis_b200.synth import synth_image, synth_descriptors
    rng = np.random.default_rng(7)
    out = {}

    def timed(fn, reps=7):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3

    # CPU side of every number below: the oracle's restatement of the same reference function (one host thread, as the reference
    # runs it) on the same inputs; the grid build is inside the timed call on both sides
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_lib as O
    import test_gpu_match as T

    # ---- SearchByProjection (src/ORBmatcher.cc:23-340)
    ncam, nmp, w = 2, 5000, 512
    rig = ob.RigExtractor(w, w, ncam, 1500, device=device)
    res = rig.extract([synth_image(w, w, 4300 + c) for c in range(ncam)], ob.make_params(ncam, 1500, 20, 7, (0, 511)))
    kps = np.concatenate([r[1] for r in res]); desc = np.concatenate([r[2] for r in res])
    cam_off = np.concatenate([[0], np.cumsum([len(r[1]) for r in res])]).astype(np.int32)
    in_view = (rng.random((nmp, 4)) < np.array([0.9, 0.5, 0, 0])).astype(np.uint8)
    src = np.stack([rng.integers(cam_off[c], cam_off[c + 1], nmp) for c in range(ncam)] + [np.zeros(nmp, np.int64)] * 2, 1)
    px = (kps['x'][src] + rng.uniform(-3, 3, (nmp, 4))).astype(np.float32); py = (kps['y'][src] + rng.uniform(-3, 3, (nmp, 4))).astype(np.float32)
    lvl = np.clip(kps['octave'][src] + rng.integers(0, 2, (nmp, 4)), 0, 7).astype(np.int32)
    md = synth_descriptors(nmp, 8)
    own = rng.random(nmp) < 0.7
    md[own] = desc[src[own, 0]] ^ np.packbits(rng.random((int(own.sum()), 256)) < 0.06, axis=1, bitorder='little')
    frame = dict(cam_off=cam_off, kps=kps, desc=desc, bounds=np.tile(np.array([0, w, 0, w], np.float32), (ncam, 1)), scale_factors=rig.scale)
    points = dict(in_view=in_view, proj_x=px, proj_y=py, level=lvl, view_cos=np.full((nmp, 4), 0.999, np.float32), desc=md)
    m = ob.ORBmatcher(0.8, device=device)
    nm = [0]

    def sbp():
        nm[0] = m.SearchByProjection(frame, points, 3.0)[0]
    ms = timed(sbp)
    frame.update(occupied=np.zeros(int(cam_off[-1]), np.uint8), l2r=None, r2l=None)
    points.update(depth=np.zeros(nmp, np.float32), bad=np.zeros(nmp, np.uint8), has_obs=np.ones(nmp, np.uint8))
    cpu_ms = timed(lambda: T._oracle_sbp(frame, points, 3.0, np.float32(0.8), False, 50.0), reps=3)
    out['search_by_projection'] = {'map_points': nmp, 'cameras': ncam, 'keypoints': int(cam_off[-1]), 'matches': int(nm[0]), 'ms_per_call': ms,
                                   'device_ms': m.last_ms(), 'map_points_per_s': nmp / (ms * 1e-3), 'cpu_ms_per_call': cpu_ms}
    # ---- SearchByProjection(CurrentFrame, LastFrame) (src/ORBmatcher.cc:1985-2413): the per-frame tracking matcher
    nlast = cam_off[-1]
    lsrc = rng.integers(0, nlast, nlast)
    lpx = np.zeros((nlast, 4), np.float32); lpy = np.zeros((nlast, 4), np.float32)
    for c in range(ncam):
        pick = np.where((lsrc >= cam_off[c]) & (lsrc < cam_off[c + 1]), lsrc, rng.integers(cam_off[c], cam_off[c + 1], nlast))
        lpx[:, c] = kps['x'][pick] + rng.uniform(-4, 4, nlast); lpy[:, c] = kps['y'][pick] + rng.uniform(-4, 4, nlast)
    last = dict(valid=np.ones(nlast, np.uint8), proj_x=lpx, proj_y=lpy, octave=kps['octave'][lsrc].astype(np.int32),
                angle=((kps['angle'][lsrc] + 10) % 360).astype(np.float32),
                desc=desc[lsrc] ^ np.packbits(rng.random((nlast, 256)) < 0.05, axis=1, bitorder='little'))

    def sbp_last():
        nm[0] = m.SearchByProjectionLast(frame, last, 7.0)[0]
    ms = timed(sbp_last)
    cpu_ms = timed(lambda: T._oracle_sbp_last(frame, last, 7.0, False, False, True), reps=3)
    out['search_by_projection_last_frame'] = {'points': int(nlast), 'cameras': ncam, 'matches': int(nm[0]), 'ms_per_call': ms,
                                              'device_ms': m.last_ms(), 'points_per_s': nlast / (ms * 1e-3), 'cpu_ms_per_call': cpu_ms}
    # ---- the same matchers batched across frames (orbx_matcher_pool, 16 lanes): F frames per call, F = 1, 8, 64
    pool = ob.MatcherPool(lanes=16, device=device)
    batched = {}
    for F_ in (1, 8, 64):
        ms_a = timed(lambda: pool.SearchByProjectionBatch([frame] * F_, [points] * F_, 3.0, False, 50.0, 0.8), reps=5)
        ms_b = timed(lambda: pool.SearchByProjectionLastBatch([frame] * F_, [last] * F_, 7.0), reps=5)
        batched['F%d' % F_] = {
            'search_by_projection': {'ms_per_batch': ms_a, 'ms_per_frame': ms_a / F_, 'map_points_per_s': F_ * nmp / (ms_a * 1e-3),
                                     'vs_cpu': out['search_by_projection']['cpu_ms_per_call'] * F_ / ms_a},
            'search_by_projection_last_frame': {'ms_per_batch': ms_b, 'ms_per_frame': ms_b / F_, 'points_per_s': F_ * int(nlast) / (ms_b * 1e-3),
                                                'vs_cpu': out['search_by_projection_last_frame']['cpu_ms_per_call'] * F_ / ms_b}}
    # ---- SearchByBoW(KeyFrame, Frame) (src/ORBmatcher.cc:349-666): 64 shared vocabulary nodes
    def fvec(d):
        node = d[:, 0].astype(np.int32) & 63
        order = np.argsort(node, kind='stable')
        nodes, counts = np.unique(node, return_counts=True)
        return nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)
    flips = np.packbits(rng.random((nlast, 256)) < 0.04, axis=1, bitorder='little'); flips[:, 0] &= 0xC0
    dkf = desc[lsrc] ^ flips
    akf = ((kps['angle'][lsrc] + 25) % 360).astype(np.float32)
    fv_kf, fv_f = fvec(dkf), fvec(desc)

    def bow():
        nm[0] = m.SearchByBoW(fv_kf, fv_f, dkf, desc, None, akf, kps['angle'], cam_off)[0]
    ms = timed(bow)
    Lo = O.lib()
    skip0 = np.zeros(nlast, np.uint8); mref = np.zeros(nlast, np.int32)

    def bow_cpu():
        Lo.orc_search_by_bow_frame(len(fv_kf[0]), O.p(fv_kf[0]), O.p(fv_kf[1]), O.p(fv_kf[2]), len(fv_f[0]), O.p(fv_f[0]), O.p(fv_f[1]), O.p(fv_f[2]),
                                   O.p(dkf), O.p(desc), int(nlast), O.p(skip0), O.p(akf), O.p(np.ascontiguousarray(kps['angle'])), ncam, O.p(cam_off),
                                   O.C.c_float(0.8), 1, O.p(mref))
    cpu_ms = timed(bow_cpu, reps=3)
    out['search_by_bow'] = {'keyframe_features': int(nlast), 'frame_features': int(nlast), 'matches': int(nm[0]), 'ms_per_call': ms,
                            'device_ms': m.last_ms(), 'cpu_ms_per_call': cpu_ms}
    for F_ in (1, 8, 64):
        ms_c = timed(lambda: pool.SearchByBoWBatch([fv_kf] * F_, [fv_f] * F_, [dkf] * F_, [desc] * F_, [None] * F_, [akf] * F_, [kps['angle']] * F_,
                                                   [cam_off] * F_, 0.8, True), reps=5)
        batched['F%d' % F_]['search_by_bow'] = {'ms_per_batch': ms_c, 'ms_per_frame': ms_c / F_, 'vs_cpu': cpu_ms * F_ / ms_c}
    batched['what'] = ('F frames per call on a pool of 16 lanes (handle + stream + host worker each); vs_cpu = F x the one-thread CPU time of the '
                       'oracle restatement / the batch time; the Python-side marshalling of the F argument sets is inside the timed call')
    out['batched_across_frames'] = batched
    pool.close()
    # ---- SearchForTriangulation (src/ORBmatcher.cc:1131-1456), BASELINE config 2 style: the two keyframes of a 752x480 stereo setup with 1200
    #      features each, FeatureVector groups = 32 vocabulary nodes; bCoarse = true (no geometry) and the run LocalMapping makes
    #      (bCoarse = false: epipole reject + Pinhole::epipolarConstrain per candidate), both with the rotation filter
    n1 = n2 = 1200
    k2 = np.zeros(n2, ob.KP_DTYPE)
    k2['x'] = rng.uniform(20, 732, n2); k2['y'] = rng.uniform(20, 460, n2); k2['octave'] = rng.integers(0, 8, n2); k2['angle'] = rng.uniform(0, 360, n2)
    d2s = synth_descriptors(n2, 4242)
    ssrc = rng.permutation(n2)[:n1]
    sfl = np.packbits(rng.random((n1, 256)) < 0.05, axis=1, bitorder='little'); sfl[:, 0] &= 0xE0
    d1s = d2s[ssrc] ^ sfl
    d1s[::6] = synth_descriptors(len(d1s[::6]), 4343)
    k1 = k2[ssrc].copy()
    k1['x'] += rng.uniform(-30, 30, n1).astype(np.float32); k1['y'] += rng.uniform(-2.5, 2.5, n1).astype(np.float32)   # near the epipolar line y2 = y1
    k1['angle'] = (k1['angle'] + rng.choice([0, 0, 0, 45, 200], n1)) % 360

    def fvec5(d):
        node = d[:, 0].astype(np.int32) & 31
        order = np.argsort(node, kind='stable')
        nodes, counts = np.unique(node, return_counts=True)
        return nodes.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)
    one = np.array([0, n1], np.int32)
    kf1 = dict(cam_off=one, kps=k1, desc=d1s, has_mp=(rng.random(n1) < 0.3).astype(np.uint8), u_right=None, fv=fvec5(d1s))
    kf2 = dict(cam_off=one.copy(), kps=k2, desc=d2s, has_mp=(rng.random(n2) < 0.3).astype(np.uint8), u_right=None, fv=fvec5(d2s))
    ep = np.array([k2['x'][7] + 2.0, k2['y'][7] - 1.0], np.float32)
    sfs = np.array([np.float32(1.2) ** i for i in range(8)], np.float32)
    sg = (sfs * sfs).astype(np.float32)
    prs = ob.epipolar_pairs()
    for s_ in range(11):
        prs['model'][s_] = ob.CAM_PINHOLE
        prs['F12'][s_] = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32)        # epipolar line y2 = y1 (a rectified pair)
    sft = {}
    for name, coarse in (('coarse', True), ('epipolar', False)):
        def run_sft():
            nm[0] = m.SearchForTriangulationFull(kf1, kf2, ep, sfs, sg, sg, prs, False, coarse, True)[0]
        ms = timed(run_sft)
        dms = m.last_ms()
        cpu_ms = timed(lambda: T._oracle_sft(kf1, kf2, ep, sfs, sg, sg, prs, False, coarse, True), reps=3)
        sft[name] = {'matches': int(nm[0]), 'ms_per_call': ms, 'device_ms': dms, 'cpu_ms_per_call': cpu_ms}
    out['search_for_triangulation'] = dict(sft, features=[n1, n2], nodes=32,
                                           what='BASELINE config 2 style (2 x 1200 features, 32 shared vocabulary nodes): bCoarse = true and the bCoarse = false '
                                                'run with the pinhole epipolar predicate on the device; CPU = the oracle restatement, one host thread')
    # ---- ComputeStereoMatches (src/Frame.cc:1057-1219)
    left = synth_image(752, 480, 2001)
    xs = np.arange(752)
    right = left[:, np.clip(xs + (19 + 16 * np.sin(xs / 90.0)).astype(np.int64), 0, 751)].copy()
    st = ob.RigExtractor(752, 480, 2, 1200, device=device)
    st.extract([left, right], ob.make_params(2, 1200, 20, 7, (0, 0)))
    ns = [0]

    def stereo():
        ns[0] = st.compute_stereo_matches(0, 1, 0.11, 47.9)[0]
    ms = timed(stereo)
    EL = O.OracleExtractor(1200, 1.2, 8, 20, 7); ER = O.OracleExtractor(1200, 1.2, 8, 20, 7)
    EL(left, (0, 0)); ER(right, (0, 0))
    ru = np.zeros(4000, np.float32); rd = np.zeros(4000, np.float32)
    Lo.orc_compute_stereo_matches.argtypes = [O.C.c_void_p, O.C.c_void_p, O.C.c_float, O.C.c_float, O.C.c_void_p, O.C.c_void_p]
    cpu_ms = timed(lambda: Lo.orc_compute_stereo_matches(EL.h, ER.h, 0.11, 47.9, O.p(ru), O.p(rd)), reps=3)
    out['compute_stereo_matches'] = {'image': '2x752x480', 'matches': int(ns[0]), 'ms_per_call': ms, 'cpu_ms_per_call': cpu_ms}
    # ---- ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483)
    sizes = rng.integers(2, 24, 2000).astype(np.int32)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    dd = np.repeat(synth_descriptors(2000, 9), sizes, axis=0) ^ np.packbits(rng.random((int(ptr[-1]), 256)) < 0.1, axis=1, bitorder='little')
    ms = timed(lambda: m.ComputeDistinctiveDescriptors(dd, ptr))
    bref = np.zeros(2000, np.int32)
    cpu_ms = timed(lambda: Lo.orc_distinctive_descriptors(O.p(dd), O.p(ptr), 2000, O.p(bref)), reps=3)
    out['distinctive_descriptors'] = {'map_points': 2000, 'descriptors': int(ptr[-1]), 'ms_per_call': ms, 'device_ms': m.last_ms(),
                                      'map_points_per_s': 2000 / (ms * 1e-3), 'cpu_ms_per_call': cpu_ms}
    # ---- ComputeStereoFishEyeMatches (src/Frame.cc:1409-1459): 2-NN + ratio + KannalaBrandt8::TriangulateMatches
    from openmavis_b200.synth import synth_fisheye_pairs
    nk = 1500
    fs = synth_fisheye_pairs(nk, 77)
    kp_dt = ob.KP_DTYPE
    kl = np.zeros(nk, kp_dt); kr = np.zeros(nk, kp_dt)
    kl['x'], kl['y'], kl['octave'] = fs['xy1'][:, 0], fs['xy1'][:, 1], fs['octave1']
    perm = rng.permutation(nk)
    kr['x'][perm], kr['y'][perm], kr['octave'][perm] = fs['xy2'][:, 0], fs['xy2'][:, 1], fs['octave2']
    dl = synth_descriptors(nk, 78); dr = np.zeros_like(dl)
    dr[perm] = dl ^ np.packbits(rng.random((nk, 256)) < 0.06, axis=1, bitorder='little')
    frig = ob.kb8_stereo(fs['cam1'], fs['cam2'], fs['R12'], fs['t12'])
    sig = np.array([np.float32(1.2) ** (2 * l) for l in range(8)], np.float32)
    nf = [0]

    def fisheye():
        nf[0] = m.ComputeStereoFishEyeMatches(frig, kl, dl, 0, kr, dr, 0, sig, 0.7)['nMatches']
    ms = timed(fisheye)
    out['stereo_fisheye_matches'] = {'keypoints': [nk, nk], 'matches': int(nf[0]), 'ms_per_call': ms, 'queries_per_s': nk / (ms * 1e-3)}
    return out


def secondary_configs(ob, torch, dev, local, stream):
    """Device-resident extraction throughput on the other BASELINE.json configurations (parity for them is covered by tests/): batch of
    synthetic images of the configuration's geometry, 2 warm-up + 5 timed steps, CUDA events on `stream`."""
    from openmavis_b200.synth import synth_image
    out = {}
    cases = [('config1_euroc_mono_752x480_nf1000', 752, 480, 1, 1000, 256, [(0, 1000)]),
             ('config2_euroc_stereo_2x752x480_nf1200', 752, 480, 2, 1200, 128, [(0, 0), (0, 0)]),
             ('config4_tumvi_stereo_2x512x512_nf1500', 512, 512, 2, 1500, 128, [(0, 511), (0, 511)]),
             ('config5_sweep_64x5x1280x1024_nf2000', 1280, 1024, 5, 2000, 64, [(0, 1280), (0, 1280), (0, 0), (0, 0), (0, 0)])]
    for name, w, h, ncam, nf, frames, laps in cases:
        n_img = frames * ncam
        base = [synth_image(w, h, 7000 + 10 * ncam + i) for i in range(4)]
        rng = np.random.default_rng(ncam)
        host = np.empty((n_img, h, w), np.uint8)
        for i in range(n_img):
            im = base[i % 4]
            if i >= 4:
                im = np.roll(im, (int(rng.integers(1, h)), int(rng.integers(1, w))), axis=(0, 1))
            host[i] = im
        d = torch.from_numpy(host).to(dev)
        rig = ob.RigExtractor(w, h, n_img, nf, 1.2, 8, device=local)
        rig.set_stream(stream.cuda_stream)
        prm = ob.make_params(n_img, nf, INI, MIN, laps * frames)
        for _ in range(2):
            rig.extract_device(d.data_ptr(), n_img, prm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(5):
            rig.extract_device(d.data_ptr(), n_img, prm)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res = rig.fetch(n_img)
        out[name] = {'cameras': ncam, 'image': [w, h], 'features_per_camera': nf, 'frames_per_step': frames, 'ms_per_step': ms,
                     'frames_per_s': frames / (ms * 1e-3), 'images_per_s': n_img / (ms * 1e-3),
                     'level0_gpixel_per_s': n_img * w * h / (ms * 1e-3) / 1e9, 'keypoints_per_image': float(np.mean([len(r[1]) for r in res]))}
        rig.close()
        del d
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path — src/ORBextractor.cc compiled unmodified from the reference
    checkout with its own flags (oracle/_ref/libref_orb_fma.so; OpenCV primitives from the cv2-pinned oracle functions), all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = max(1, min(args.ref_frames, 64))
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_run(max(2, threads // NCAM + 1), threads)
    tot_t = 0.0
    for _ in range(args.steps):
        _, dt, _ = cpu_reference_run(frames, threads)
        tot_t += dt
    value = frames * args.steps / tot_t
    line = {
        'impl': 'reference', 'metric': 'rig-frames/sec ORB extract (5×720×540, 1500 kp/cam)', 'value': value, 'unit': 'rig-frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': CONFIG, 'run': {'frames_per_step': frames},
        'cpu_baseline': {'value': value, 'unit': 'rig-frames/s', 'cores': threads, 'kind': 'reference',
                         'sample': '%d rig frames x %d steps, %d host threads (one image per thread at a time); src/ORBextractor.cc compiled '
                                   'unmodified (oracle/_ref/libref_orb_fma.so)' % (frames, args.steps, threads)},
        'e2e': {'value': value, 'unit': 'rig-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=64, help='rig frames per step per GPU')
    ap.add_argument('--ref-frames', type=int, default=32, help='rig frames per step of the CPU reference arm')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='budget of the cpu_baseline sample')
    ap.add_argument('--no-matching', action='store_true')
    ap.add_argument('--no-numa-bind', action='store_true', help='do not pin the rank to the CPU cores next to its GPU')
    ap.add_argument('--knn-n', type=int, default=100000)
    ap.add_argument('--sustain-seconds', type=float, default=2.0, help='length of the sustained-rate run (0 = skip)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the other BASELINE.json configurations')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import openmavis_b200 as ob

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    # one process per GPU: run it (and first-touch its pinned staging buffers) on the NUMA node the GPU hangs off
    affinity0 = os.sched_getaffinity(0)
    numa_node, numa_nodes = None, None
    from openmavis_b200.shard import bind_to_gpu_numa, host_numa_nodes
    numa_nodes = host_numa_nodes()
    if not args.no_numa_bind:
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, a) for a in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            numa_node = bind_to_gpu_numa('%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    frames = args.frames
    n_img = frames * NCAM
    host_imgs = make_batch(frames, rank)
    pinned = torch.from_numpy(host_imgs).pin_memory()
    d_imgs = pinned.to(dev, non_blocking=False)
    params = ob.make_params(n_img, NFEAT, INI, MIN, LAPS * frames)
    rig = ob.RigExtractor(W, H, n_img, NFEAT, 1.2, 8, device=local)
    # a real (non-default) stream: liborbx treats a NULL handle as "use the handle's own stream", and torch events only
    # see the stream they are recorded on — kernels and events must share this one
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    rig.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # L2 flush buffer (inputs are > L2 anyway)

    # ---------------- device-resident throughput (`value`) ----------------
    # Two extractor handles work alternately on the batches (step k on handle k & 1, each on its own stream), as a throughput
    # pipeline does: the latency-bound quadtree of one batch runs under the issue-bound pyramid / FAST of the next.  Timed with CUDA
    # events on `stream`: the second handle's stream starts after e0 and is joined before e1.
    rig_b = ob.RigExtractor(W, H, n_img, NFEAT, 1.2, 8, device=local)
    stream_b = torch.cuda.Stream(device=dev)
    rig_b.set_stream(stream_b.cuda_stream)
    pair = (rig, rig_b)
    for k in range(max(args.warmup, 2)):
        pair[k & 1].extract_device(d_imgs.data_ptr(), n_img, params)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    stream_b.wait_event(e0)
    for k in range(args.steps):
        pair[k & 1].extract_device(d_imgs.data_ptr(), n_img, params)
    stream.wait_stream(stream_b)
    e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    res_b = rig_b.fetch(n_img)
    rig_b.close()
    launches_per_step = rig.last_launches()          # launches of one timed step (all chunks)
    # per-stage durations: a second, untimed-for-`value` pass with CUDA events between the stages (same stream)
    rig.set_profiling(True)
    for _ in range(min(args.steps, 16)):
        rig.extract_device(d_imgs.data_ptr(), n_img, params)
    barrier()
    clocks = sampler.stop(t_wall0, time.perf_counter()) if sampler else None
    stage = rig.stage_ms()
    fast_stats = rig.fast_stats()                    # k_fast counters over the profiled calls
    fast_stats['profiled_calls'] = min(args.steps, 16)
    rig.set_profiling(False)
    value = world * frames * args.steps / (ms_total * 1e-3)
    # ---------------- sustained rate: the same device pipeline for >= sustain-seconds, with its own clock record ----------------
    sustained = None
    if args.sustain_seconds > 0:
        rig_b2 = ob.RigExtractor(W, H, n_img, NFEAT, 1.2, 8, device=local)
        rig_b2.set_stream(stream_b.cuda_stream)
        pair2 = (rig, rig_b2)
        n_sus = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / (ms_total / args.steps))))
        for k in range(2):
            pair2[k & 1].extract_device(d_imgs.data_ptr(), n_img, params)
        barrier()
        samp2 = ClockSampler(local) if rank == 0 else None
        tw0 = time.perf_counter()
        e0.record(stream)
        stream_b.wait_event(e0)
        for k in range(n_sus):
            pair2[k & 1].extract_device(d_imgs.data_ptr(), n_img, params)
        stream.wait_stream(stream_b)
        e1.record(stream)
        barrier()
        tw1 = time.perf_counter()
        ms_sus = max_over_ranks(e0.elapsed_time(e1))
        sustained = {'value': world * frames * n_sus / (ms_sus * 1e-3), 'unit': 'rig-frames/s', 'steps': n_sus, 'seconds': ms_sus * 1e-3,
                     'ms_per_step': ms_sus / n_sus, 'clocks': samp2.stop(tw0, tw1) if samp2 else None}
        rig_b2.close()
    res = rig.fetch(n_img)
    for i in (0, n_img - 1):                         # both handles of the timed pipeline produced the same results
        assert np.array_equal(res[i][1], res_b[i][1]) and np.array_equal(res[i][2], res_b[i][2])
    kp_per_image = float(np.mean([len(r[1]) for r in res]))

    # ---------------- end to end through the host-buffer C-ABI call (`e2e`) ----------------
    # Every step uploads its 320 images from pinned host memory and downloads keypoints + descriptors.  Throughput pipeline of
    # the public API: two handles used alternately (orbx_extract_batch_submit / orbx_extract_wait), so the copies of one batch
    # overlap the kernels of the other; a step's outputs are waited for before its handle is reused.
    import ctypes as C
    ptrs = (C.c_void_p * n_img)(*[pinned.data_ptr() + i * W * H for i in range(n_img)])
    strides = np.full(n_img, W, np.int32)
    cap = rig.cap
    rig2 = ob.RigExtractor(W, H, n_img, NFEAT, 1.2, 8, device=local)
    lanes = []
    for h in (rig, rig2):
        out_kps = torch.empty((n_img, cap, 28), dtype=torch.uint8).pin_memory()
        out_desc = torch.empty((n_img, cap, 32), dtype=torch.uint8).pin_memory()
        out_n = torch.empty((2, n_img), dtype=torch.int32).pin_memory()
        lanes.append((h, out_kps.numpy().view(ob.KP_DTYPE).reshape(n_img, cap), out_desc.numpy(), out_n.numpy()[0], out_n.numpy()[1],
                      (out_kps, out_desc, out_n)))
    rig.set_stream(0)                              # the handles' own streams: the two lanes must not share one
    e2e_steps = max(4, args.steps)

    def e2e_run(k_steps):
        for k in range(k_steps):
            h, kk, dd, nk, nm, _ = lanes[k & 1]
            h.wait()
            h.submit_host_raw(ptrs, strides, params, kk, dd, nk, nm, n_img)
        for h, *_ in lanes:
            h.wait()
    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = world * frames * e2e_steps / (ms_e2e * 1e-3)
    # copy roofline: what the host link alone allows — every rank moves the same bytes (pinned images in, keypoints + descriptors +
    # counts out) with plain asynchronous copies on two streams (both directions at once), no kernels
    cr_in = torch.empty((n_img, H, W), dtype=torch.uint8, device=dev)
    cr_out_d = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in lanes[0][5]]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_run(k_steps):
        for _ in range(k_steps):
            with torch.cuda.stream(s_in):
                cr_in.copy_(pinned, non_blocking=True)
            with torch.cuda.stream(s_out):
                for hbuf, dbuf in zip(lanes[0][5], cr_out_d):
                    hbuf.copy_(dbuf, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    copy_run(2)
    ms_copy = None
    for _ in range(3):                                     # a roofline is what the link CAN do: the best of three trials (max over ranks each)
        barrier()
        t0 = time.perf_counter()
        copy_run(e2e_steps)
        barrier()
        t = max_over_ranks((time.perf_counter() - t0) * 1e3)
        ms_copy = t if ms_copy is None else min(ms_copy, t)
    copy_value = world * frames * e2e_steps / (ms_copy * 1e-3)
    del cr_in, cr_out_d
    # latency-oriented variant: one synchronous orbx_extract_batch call per step
    kps_np, desc_np, nkp_np, nmono_np = lanes[0][1], lanes[0][2], lanes[0][3], lanes[0][4]
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        rig.extract_host_raw(ptrs, strides, params, kps_np, desc_np, nkp_np, nmono_np, n_img)
    barrier()
    ms_sync_call = max_over_ranks((time.perf_counter() - t0) * 1e3) / 5
    # latency of ONE rig frame (5 images) through the same synchronous host-buffer call: what a live tracker sees per frame
    one = (C.c_void_p * NCAM)(*[ptrs[i] for i in range(NCAM)])
    for _ in range(5):
        rig.extract_host_raw(one, strides, params, kps_np, desc_np, nkp_np, nmono_np, NCAM)
    lat = []
    for _ in range(40):
        t0 = time.perf_counter()
        rig.extract_host_raw(one, strides, params, kps_np, desc_np, nkp_np, nmono_np, NCAM)
        lat.append((time.perf_counter() - t0) * 1e3)
    ms_one_frame = float(np.median(lat))
    rig.set_stream(stream.cuda_stream)
    h2d = n_img * W * H + n_img * (20 + 32)
    d2h = n_img * cap * (28 + 32) + 2 * 4 * n_img
    # the e2e results must equal the device-resident ones (both lanes)
    for lane in lanes:
        for i in (0, n_img - 1):
            assert lane[3][i] == len(res[i][1]) and np.array_equal(lane[2][i, :lane[3][i]], res[i][2])
    rig2.close()

    # ---------------- roofline of the dominant kernel ----------------
    peaks, peak_kind = measured_peaks()
    kernels = {k: stage[k] for k in ('pyramid', 'blur', 'fast', 'octree', 'slots', 'orient_desc')}
    dom = max(kernels, key=kernels.get)
    px = n_img * W * H
    alg = {'pyramid': ALG_BYTES_PER_PX['pyramid'] * px, 'blur': ALG_BYTES_PER_PX['blur'] * px, 'fast': ALG_BYTES_PER_PX['fast'] * px,
           'octree': None, 'slots': None, 'orient_desc': None}
    hbm_dom = dom if alg[dom] else max(('pyramid', 'blur', 'fast'), key=lambda k: kernels[k])
    achieved = alg[hbm_dom] / (kernels[hbm_dom] * 1e-3) / 1e9
    traffic = None          # dram bytes of the same stage from the committed ncu --set full capture (profiles/ncu_traffic.json)
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        traffic = tj['bytes'][hbm_dom] * n_img / tj['images']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': hbm_dom, 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'], 'traffic': traffic, 'algorithmic_bytes': alg[hbm_dom], 'peak_source': peak_kind,
                'stage_ms': stage, 'dominant_stage': dom,
                'pipeline_frac': ALG_BYTES_PER_PX['total'] * px / (stage['total'] * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'per_kernel_frac': {k: (alg[k] / (kernels[k] * 1e-3) / 1e9 / peaks['hbm_gbs']) for k in ('pyramid', 'blur', 'fast')}}

    # ---------------- secondary metric: brute-force Hamming 2-NN ----------------
    matching = None
    if not args.no_matching:
        from openmavis_b200.synth import synth_descriptors, synth_query_from_train
        n = args.knn_n
        m = ob.ORBmatcher(0.8, device=local)
        m.set_stream(stream.cuda_stream)
        train = synth_descriptors(n, 5)
        query = synth_query_from_train(train, 6)
        from openmavis_b200 import shard
        lo, cnt = shard.row_shard(n, rank, world)        # query-sharded; train rows start distributed the same way
        hi = lo + cnt
        d_q = torch.from_numpy(query[lo:hi]).to(dev)
        d_local_train = torch.from_numpy(train[lo:hi]).to(dev)
        d_idx = torch.empty((hi - lo, 2), dtype=torch.int32, device=dev)
        d_dist = torch.empty((hi - lo, 2), dtype=torch.int32, device=dev)

        shared = None
        if world > 1:
            ok = 1
            try:
                shared = ob.SharedRows(dist, local, n)
            except Exception as ex:                      # e.g. CUDA IPC not permitted in this container
                ok = 0
                print('rank %d: peer mapping failed (%s); falling back to an NCCL all-gather' % (rank, ex), file=sys.stderr)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if shared is not None:
                    shared.close(barrier=False)
                shared = None
        token = torch.zeros(1, device=dev)
        step_no = [0]

        def knn_step():
            if world > 1 and shared is None:   # fallback: one NCCL all-gather of the train rows, then the local 2-NN
                d_t = shard.gather_train_rows(d_local_train, n, dist)
                m.knnMatch2_device(d_q.data_ptr(), hi - lo, d_t.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
            elif world > 1:
                # The one real exchange step, fused with the compute: every rank publishes its train rows in a peer-mapped buffer and
                # an epoch flag; the pull kernel of each rank waits for the peers' flags on the device and fetches their rows over
                # NVLink on an auxiliary stream while the main stream already matches the own rows (see openmavis_b200.SharedRows):
                # no collective call, no host synchronisation.
                step_no[0] += 1
                buf = step_no[0] & 1
                shared.fill(buf, d_local_train.data_ptr(), stream.cuda_stream)
                shared.publish(step_no[0], stream.cuda_stream)
                m.knnMatch2_device_parts_sync(d_q.data_ptr(), hi - lo, shared.ptrs[buf], shared.rows, rank, shared.flags, step_no[0],
                                              d_idx.data_ptr(), d_dist.data_ptr(), timeout_s=20.0)
            else:
                m.knnMatch2_device(d_q.data_ptr(), hi - lo, d_local_train.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        for _ in range(2):
            knn_step()
        barrier()
        ksteps = 5
        e0.record(stream)
        for _ in range(ksteps):
            knn_step()
        e1.record(stream)
        barrier()
        ms_knn = max_over_ranks(e0.elapsed_time(e1)) / ksteps
        pairs = float(n) * float(n)
        popc = ob.popc_peak(local)
        matching = {'metric': 'Hamming matches/sec (brute-force 2-NN %dx%d)' % (n, n), 'value': pairs / (ms_knn * 1e-3), 'unit': 'pairs/s',
                    'ms': ms_knn, 'popc_gops_achieved': 8 * pairs / (ms_knn * 1e-3) / 1e9 / world, 'popc_gops_peak_measured': popc,
                    'frac_of_popc_peak': 8 * pairs / (ms_knn * 1e-3) / 1e9 / world / popc,
                    'executed_popc_per_pair': 6, 'executed_popc_frac_of_peak': 6 * pairs / (ms_knn * 1e-3) / 1e9 / world / popc,
                    'note': 'algorithmic work = 8 popc.b32 per pair (SURVEY.md 8d); the kernel executes 6 (two carry-save adders per pair), so the '
                            'XU-pipe utilisation is the executed figure', 'n_gpus': world,
                    'collective': None if world == 1 else ('none: peer-mapped train rows pulled over NVLink (cudaIpc) behind device-side epoch flags, overlapped with the 2-NN over the own rows' if shared is not None else 'nccl all_gather of the train rows (peer mapping unavailable)')}
        if rank == 0:
            torch.cuda.set_stream(torch.cuda.default_stream(dev))
            matching['guided'] = guided_matchers(ob, local)
            torch.cuda.set_stream(stream)
        idx_h = d_idx.cpu().numpy(); dist_h = d_dist.cpu().numpy()
        chk = np.unpackbits(query[lo:hi][:256] ^ train[idx_h[:256, 0]], axis=1).sum(axis=1)
        assert np.array_equal(chk, dist_h[:256, 0])
        assert (idx_h[:256, 0] >= 0).all() and (idx_h[:256, 0] < n).all()
        if shared is not None:
            torch.cuda.synchronize()
            shared.close()

    # ---------------- the other BASELINE.json configurations (rank 0, N == 1 only) ----------------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = secondary_configs(ob, torch, dev, local, stream)

    # ---------------- CPU baseline (rank 0, N == 1 only; bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        os.sched_setaffinity(0, affinity0)             # the CPU arm uses every host core again
        threads = os.cpu_count() or 1
        cpu_reference_run(max(2, threads // NCAM + 1), threads)       # warm-up (page in, malloc arenas of the worker threads)
        nfr, done, spent = 32, 0, 0.0
        while spent < args.cpu_seconds:                 # bounded sample: repeat a 32-rig-frame batch for ~cpu_seconds
            _, dt, _ = cpu_reference_run(nfr, threads)
            done += nfr; spent += dt
        v_tpc, dt_tpc, _ = cpu_reference_thread_per_camera(max(4, int(4.0 * done / spent / max(1, threads // NCAM))))
        v_port, dt_port, _ = cpu_oracle_run(32, threads)
        cpu = {'value': done / spent, 'unit': 'rig-frames/s', 'cores': threads, 'kind': 'reference',
               'sample': '%d rig frames (%d images) of the same workload, %.1f s of extraction, %d host threads; src/ORBextractor.cc compiled '
                         'unmodified with the reference flags (oracle/_ref/libref_orb_fma.so), OpenCV primitives from the cv2-pinned oracle' % (
                             done, done * NCAM, spent, threads),
               'thread_per_camera': {'value': v_tpc, 'unit': 'rig-frames/s', 'threads': NCAM, 'seconds': dt_tpc,
                                     'what': 'rig frames one after the other, one std::thread per camera per frame (src/Frame.cc:1841-1862)'},
               'oracle_port': {'value': v_port, 'unit': 'rig-frames/s', 'cores': threads, 'seconds': dt_port},
               'cv2_primitives': cv2_primitive_line()}

    if rank == 0:
        line = {
            'metric': 'rig-frames/sec ORB extract (5×720×540, 1500 kp/cam)', 'value': value, 'unit': 'rig-frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': CONFIG,
            'run': {'frames_per_step_per_gpu': frames, 'images_per_step_per_gpu': n_img, 'keypoints_per_image': kp_per_image,
                    'pipeline': 'two extractor handles alternate over the steps (one stream each)',
                    'host_numa_node': numa_node, 'host_numa_nodes': numa_nodes, 'host_cpus': len(affinity0),
                    'bytes_per_step_mb': (px * 1.0 + px * 3.7 * 2) / 1e6, 'fast_counters': fast_stats},
            'e2e': {'value': e2e_value, 'unit': 'rig-frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / e2e_steps, 'steps': e2e_steps, 'timing': 'host wall clock between barrier + synchronize',
                    'pipeline': 'two handles alternating (orbx_extract_batch_submit / orbx_extract_wait)',
                    'sync_call_ms_per_step': ms_sync_call, 'sync_call_value': world * frames / (ms_sync_call * 1e-3),
                    'one_rig_frame_latency_ms': ms_one_frame,
                    'copy_roofline': {'value': copy_value, 'unit': 'rig-frames/s', 'ms_per_step': ms_copy / e2e_steps,
                                      'h2d_gb_per_s_per_gpu': h2d / (ms_copy / e2e_steps * 1e-3) / 1e9, 'd2h_gb_per_s_per_gpu': d2h / (ms_copy / e2e_steps * 1e-3) / 1e9,
                                      'what': 'the same pinned buffers moved by plain cudaMemcpyAsync on two streams, all ranks at once, no kernels; best of 3 trials'},
                    'frac_of_copy_roofline': e2e_value / copy_value},
            'gpu_launches': launches_per_step * args.steps,
            'clocks': clocks, 'roofline': roofline, 'sustained': sustained, 'cpu_baseline': cpu, 'secondary': secondary, 'matching': matching,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
