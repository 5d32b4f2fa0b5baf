"""openmavis_b200 — B200-native ORB extraction + Hamming matching (the OpenMAVIS / ORB-SLAM3 front-end hot path).

Python host-side mirror of the reference's operator interface on top of the C ABI of liborbx.so
(include/orbx.h).  The classes keep the reference's names and argument meaning:

  ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)      include/ORBextractor.h:28-84
      __call__(image, mask, vLappingArea) -> (n_mono, keypoints, descriptors)   src/ORBextractor.cc:987-1071
  ORBmatcher(nnratio, checkOri)                                             include/ORBmatcher.h:17-123
      DescriptorDistance, SearchByProjection, SearchForTriangulation, ...

There is no CPU fallback: importing works anywhere (so the CPU test-suite can check the ABI surface), but every
compute call raises OrbxError unless liborbx.so is built and a CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liborbx.so')

KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'),
                     ('octave', 'i4'), ('class_id', 'i4')])
PARAMS_DTYPE = np.dtype([('nfeatures', 'i4'), ('ini_th_fast', 'i4'), ('min_th_fast', 'i4'), ('lap0', 'i4'), ('lap1', 'i4')])
NUM_STAGES = 7
STAGE_NAMES = ('pyramid', 'blur', 'fast', 'octree', 'slots', 'orient_desc', 'total')

TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30     # src/ORBmatcher.cc:15-17
GATE_AROUND, GATE_FORWARD, GATE_BACKWARD, GATE_BELOW = 0, 1, 2, 3   # ORBX_GATE_*


class OrbxError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads liborbx.so (built in-tree by openmavis_b200/csrc/Makefile).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbxError('liborbx.so not built: run `make -C openmavis_b200/csrc` (or __graft_entry__.build())')
        L = C.CDLL(LIB_PATH)
        L.orbx_last_error.restype = C.c_char_p
        L.orbx_version.restype = C.c_char_p
        L.orbx_extractor_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                            C.c_int, C.c_int]
        L.orbx_extractor_destroy.argtypes = [C.c_void_p]
        L.orbx_matcher_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.orbx_matcher_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


KB8_STEREO_DTYPE = np.dtype([('cam1', 'f4', 8), ('cam2', 'f4', 8), ('R12', 'f4', 9), ('t12', 'f4', 3), ('precision1', 'f4'),
                             ('precision2', 'f4')])   # orbx_kb8_stereo


EPIPOLAR_PAIR_DTYPE = np.dtype([('model', 'i4'), ('reserved', 'i4'), ('F12', 'f4', (9,)), ('kb8', KB8_STEREO_DTYPE)])
CAM_PINHOLE, CAM_KB8 = 0, 1


def epipolar_pairs():
    """11 x orbx_epipolar_pair (include/orbx.h): [0] the single pair of a monocular keyframe pair, [1] ll, [2] lr, [3] rl, [4] rr,
    [5] l-sl, [6] sl-l, [7] sl-sl, [8] r-sr, [9] sr-r, [10] sr-sr.  Fill ['model'] and ['F12'] (pinhole) or ['kb8'] (kb8_stereo())."""
    return np.zeros(11, EPIPOLAR_PAIR_DTYPE)


def kb8_stereo(cam1, cam2, R12, t12, precision=1e-6):
    """orbx_kb8_stereo: cam = (fx, fy, cx, cy, k0..k3) of KannalaBrandt8::mvParameters, R12 / t12 = Frame::mRlr / mtlr."""
    r = np.zeros(1, KB8_STEREO_DTYPE)
    r['cam1'][0] = np.asarray(cam1, np.float32); r['cam2'][0] = np.asarray(cam2, np.float32)
    r['R12'][0] = np.asarray(R12, np.float32).reshape(9); r['t12'][0] = np.asarray(t12, np.float32).reshape(3)
    r['precision1'] = np.float32(precision); r['precision2'] = np.float32(precision)
    return r


def _check(status):
    if status < 0:
        raise OrbxError('orbx status %d: %s' % (status, lib().orbx_last_error().decode()))
    return status


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count():
    return lib().orbx_device_count()


def features_per_level(nfeatures, scale_factor=1.2, nlevels=8):
    out = np.zeros(nlevels, np.int32)
    _check(lib().orbx_features_per_level(int(nfeatures), C.c_float(scale_factor), int(nlevels), _p(out)))
    return out


def make_params(n, nfeatures, ini_th, min_th, lapping=(0, 0)):
    p = np.zeros(n, PARAMS_DTYPE)
    p['nfeatures'] = nfeatures; p['ini_th_fast'] = ini_th; p['min_th_fast'] = min_th
    lap = np.asarray(lapping, np.int32).reshape(-1, 2)
    p['lap0'] = lap[:, 0]; p['lap1'] = lap[:, 1]
    return p


class RigExtractor:
    """Batched extractor for one image geometry: all cameras of a rig frame, or many rig frames, per call."""

    def __init__(self, width, height, max_images, max_nfeatures, scale_factor=1.2, nlevels=8, device=0, fma_mode=0):
        self.L = lib()
        self.width, self.height, self.max_images, self.nlevels = int(width), int(height), int(max_images), int(nlevels)
        self.max_nfeatures = int(max_nfeatures)
        self.scale_factor = float(scale_factor)
        h = C.c_void_p()
        _check(self.L.orbx_extractor_create(C.byref(h), int(device), self.width, self.height, self.max_images,
                                            self.max_nfeatures, C.c_float(scale_factor), self.nlevels, int(fma_mode)))
        self.h = h
        self.cap = _check(self.L.orbx_extractor_capacity(self.h))
        n = self.nlevels
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = (np.zeros(n, np.float32) for _ in range(4))
        self.level_w, self.level_h = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _check(self.L.orbx_extractor_tables(self.h, _p(self.scale), _p(self.inv_scale), _p(self.sigma2), _p(self.inv_sigma2),
                                            _p(self.level_w), _p(self.level_h)))

    def close(self):
        if getattr(self, 'h', None):
            self.L.orbx_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self.L.orbx_extractor_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_profiling(self, on):
        _check(self.L.orbx_extractor_set_profiling(self.h, int(bool(on))))

    def stage_ms(self):
        ms = np.zeros(NUM_STAGES, np.float32)
        _check(self.L.orbx_extractor_stage_ms(self.h, _p(ms)))
        return dict(zip(STAGE_NAMES, ms.tolist()))

    def fast_stats(self):
        """FAST counters over the profiled calls since set_profiling(True): tested pixels, quick-test survivors, corners, kept, ..."""
        st = np.zeros(6, np.uint64)
        _check(self.L.orbx_extractor_fast_stats(self.h, _p(st)))
        return dict(zip(('tested_pixels', 'quick_survivors', 'corners', 'kept_after_nms', 'retried_cells', 'tiles'), (int(x) for x in st)))

    def last_launches(self):
        return self.L.orbx_extractor_last_launches(self.h)

    def alloc_outputs(self, n):
        return (np.zeros((n, self.cap), KP_DTYPE), np.zeros((n, self.cap, 32), np.uint8), np.zeros(n, np.int32),
                np.zeros(n, np.int32))

    def extract(self, images, params, out=None):
        """images: list/array of HxW uint8 host images.  Returns [(n_mono, keypoints, descriptors)] per image."""
        n = len(images)
        imgs = [np.ascontiguousarray(im) if im.strides[1] != 1 else im for im in images]
        for im in imgs:
            if im.dtype != np.uint8 or im.ndim != 2 or im.shape != (self.height, self.width):
                raise OrbxError('image must be CV_8UC1 of %dx%d' % (self.width, self.height))
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        strides = np.array([im.strides[0] for im in imgs], np.int32)
        params = np.ascontiguousarray(params, PARAMS_DTYPE)
        kps, desc, nkp, nmono = out if out is not None else self.alloc_outputs(n)
        _check(self.L.orbx_extract_batch(self.h, n, ptrs, _p(strides), _p(params), _p(kps), _p(desc), self.cap, _p(nkp), _p(nmono)))
        return [(int(nmono[i]), kps[i, :nkp[i]], desc[i, :nkp[i]]) for i in range(n)]

    def extract_host_raw(self, ptrs, strides, params, kps, desc, nkp, nmono, n):
        """Zero-overhead variant for benchmarks: ctypes pointer array + preallocated outputs."""
        _check(self.L.orbx_extract_batch(self.h, n, ptrs, _p(strides), _p(params), _p(kps), _p(desc), self.cap, _p(nkp), _p(nmono)))

    def submit_host_raw(self, ptrs, strides, params, kps, desc, nkp, nmono, n):
        """Asynchronous extract_host_raw: returns after enqueueing; call wait() before reading the outputs."""
        _check(self.L.orbx_extract_batch_submit(self.h, n, ptrs, _p(strides), _p(params), _p(kps), _p(desc), self.cap, _p(nkp), _p(nmono)))

    def wait(self):
        _check(self.L.orbx_extract_wait(self.h))

    def extract_device(self, d_images_ptr, n, params):
        """Device-resident input (n tightly packed images at device address d_images_ptr); asynchronous."""
        params = np.ascontiguousarray(params, PARAMS_DTYPE)
        _check(self.L.orbx_extract_batch_device(self.h, int(n), C.c_void_p(d_images_ptr), _p(params)))

    def fetch(self, n, out=None):
        kps, desc, nkp, nmono = out if out is not None else self.alloc_outputs(n)
        _check(self.L.orbx_extract_fetch(self.h, _p(kps), _p(desc), self.cap, _p(nkp), _p(nmono)))
        return [(int(nmono[i]), kps[i, :nkp[i]], desc[i, :nkp[i]]) for i in range(n)]

    def device_results(self):
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self.L.orbx_extract_device_results(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def compute_stereo_fisheye_matches(self, matcher, rig, img_left, img_right, counts_left, counts_right, level_sigma2=None, ratio=0.7):
        """Frame::ComputeStereoFishEyeMatches / ComputeMultiFishEyeMatches (src/Frame.cc:1409-1525) on two images of the last call,
        reading keypoints and descriptors where the extraction left them on the device.  counts_* = (n_kp, n_mono) the extraction
        returned for the image.  Returns the dict of ORBmatcher.ComputeStereoFishEyeMatches."""
        d_kps, d_desc, _, _ = self.device_results()
        nl, ml = int(counts_left[0]), int(counts_left[1]); nr, mr = int(counts_right[0]), int(counts_right[1])
        sg = np.ascontiguousarray(self.sigma2 if level_sigma2 is None else level_sigma2, np.float32)
        l2r = np.full(max(nl, 1), -1, np.int32); r2l = np.full(max(nr, 1), -1, np.int32)
        depth = np.full(max(nl, 1), -1, np.float32); p3d = np.zeros((max(nl, 1), 3), np.float32); cnt = np.zeros(2, np.int32)
        kp_b = KP_DTYPE.itemsize
        _check(self.L.orbx_stereo_fisheye_matches_device(
            matcher.h, _p(rig), C.c_void_p(d_kps + img_left * self.cap * kp_b), C.c_void_p(d_desc + img_left * self.cap * 32), nl, ml,
            C.c_void_p(d_kps + img_right * self.cap * kp_b), C.c_void_p(d_desc + img_right * self.cap * 32), nr, mr, _p(sg), sg.size,
            C.c_double(ratio), _p(l2r), _p(r2l), _p(depth), _p(p3d), _p(cnt)))
        return dict(nMatches=int(cnt[0]), descMatches=int(cnt[1]), mvLeftToRightMatch=l2r[:nl], mvRightToLeftMatch=r2l[:nr],
                    mvDepth=depth[:nl], mvStereo3Dpoints=p3d[:nl])

    def compute_stereo_matches(self, img_left, img_right, mb, mbf, right=None):
        """Frame::ComputeStereoMatches (src/Frame.cc:1057-1219) on two images of the last call(s) -> (n, mvuRight, mvDepth).
        right: the RigExtractor that holds the right image (default: this one)."""
        u = np.zeros(self.cap, np.float32); d = np.zeros(self.cap, np.float32); n = C.c_int(0)
        _check(self.L.orbx_compute_stereo_matches2(self.h, int(img_left), (right or self).h, int(img_right), C.c_float(mb),
                                                   C.c_float(mbf), _p(u), _p(d), self.cap, C.byref(n)))
        return n.value, u, d

    def pyramid_level(self, image, level, with_border=False):
        w, h = int(self.level_w[level]), int(self.level_h[level])
        if with_border:
            w, h = w + 38, h + 38
        out = np.zeros((h, w), np.uint8)
        _check(self.L.orbx_extractor_pyramid_level(self.h, int(image), int(level), int(with_border), _p(out), out.strides[0]))
        return out

    def debug_blur_level(self, image, level):
        out = np.zeros((int(self.level_h[level]), int(self.level_w[level])), np.uint8)
        _check(self.L.orbx_debug_blur_level(self.h, int(image), int(level), _p(out), out.strides[0]))
        return out

    def debug_candidates(self, image, level):
        n = C.c_int(0)
        _check(self.L.orbx_debug_candidates(self.h, int(image), int(level), None, None, None, 0, C.byref(n)))
        xs, ys, sc = (np.zeros(max(n.value, 1), np.int32) for _ in range(3))
        _check(self.L.orbx_debug_candidates(self.h, int(image), int(level), _p(xs), _p(ys), _p(sc), n.value, C.byref(n)))
        return xs[:n.value], ys[:n.value], sc[:n.value]

    def debug_level_keypoints(self, image, level):
        n = C.c_int(0)
        kps = np.zeros(self.cap, KP_DTYPE)
        _check(self.L.orbx_debug_level_keypoints(self.h, int(image), int(level), _p(kps), self.cap, C.byref(n)))
        return kps[:n.value]


class ORBextractor:
    """Drop-in mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:28-84) for one camera."""
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, fma_mode=0):
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(np.float32(scaleFactor)), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self.device, self.fma_mode = device, fma_mode
        self._rig = None
        lib()   # fail early if the native library is missing

    def _handle(self, w, h):
        if self._rig is None or (self._rig.width, self._rig.height) != (w, h):
            if self._rig is not None:
                self._rig.close()
            self._rig = RigExtractor(w, h, 1, self.nfeatures, self.scaleFactor, self.nlevels, self.device, self.fma_mode)
        return self._rig

    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints, descriptors); (-1, empty, empty) for an empty image
        (src/ORBextractor.cc:994-995).  `mask` is ignored, as in the reference."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OrbxError('image.type() must be CV_8UC1 (src/ORBextractor.cc:998)')
        rig = self._handle(image.shape[1], image.shape[0])
        params = make_params(1, self.nfeatures, self.iniThFAST, self.minThFAST, vLappingArea)
        return rig.extract([image], params)[0]

    @property
    def mvImagePyramid(self):
        if self._rig is None:
            return []
        return [self._rig.pyramid_level(0, l) for l in range(self.nlevels)]

    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def _tables(self):
        n = self.nlevels
        out = [np.zeros(n, np.float32) for _ in range(4)]
        if self._rig is not None:
            return self._rig.scale, self._rig.inv_scale, self._rig.sigma2, self._rig.inv_sigma2
        sf = np.ones(n, np.float32)
        for i in range(1, n):
            sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(self.scaleFactor)))
        out[0][:] = sf; out[1][:] = np.float32(1.0) / sf; out[2][:] = sf * sf; out[3][:] = np.float32(1.0) / (sf * sf)
        return out

    def GetScaleFactors(self):
        return self._tables()[0].copy()

    def GetInverseScaleFactors(self):
        return self._tables()[1].copy()

    def GetScaleSigmaSquares(self):
        return self._tables()[2].copy()

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3].copy()


class ORBmatcher:
    """Hamming cores of ORB_SLAM3::ORBmatcher (include/ORBmatcher.h) and of the cv::BFMatcher call in Frame."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = TH_LOW, TH_HIGH, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.L = lib()
        self.mfNNratio = float(np.float32(nnratio))
        self.mbCheckOrientation = bool(checkOri)
        self.device = device
        h = C.c_void_p()
        _check(self.L.orbx_matcher_create(C.byref(h), int(device)))
        self.h = h

    def close(self):
        if getattr(self, 'h', None):
            self.L.orbx_matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self.L.orbx_matcher_set_stream(self.h, C.c_void_p(cuda_stream)))

    def last_ms(self):
        ms = C.c_float(0)
        _check(self.L.orbx_matcher_last_ms(self.h, C.byref(ms)))
        return ms.value

    def last_launches(self):
        return self.L.orbx_matcher_last_launches(self.h)

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """src/ORBmatcher.cc:2577-2591.  a, b: (32,) or (n,32) uint8.  Returns int or int array."""
        a2 = np.ascontiguousarray(np.atleast_2d(a), np.uint8); b2 = np.ascontiguousarray(np.atleast_2d(b), np.uint8)
        assert a2.shape == b2.shape and a2.shape[1] == 32
        d = np.zeros(a2.shape[0], np.int32)
        _check(lib().orbx_descriptor_distance(int(device), _p(a2), _p(b2), a2.shape[0], _p(d)))
        return int(d[0]) if np.ndim(a) == 1 else d

    def knnMatch2(self, query, train):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) (src/Frame.cc:1427,1483) -> idx[nq,2], dist[nq,2]."""
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        idx = np.full((q.shape[0], 2), -1, np.int32); dist = np.full((q.shape[0], 2), -1, np.int32)
        _check(self.L.orbx_knn2(self.h, _p(q), q.shape[0], _p(t), t.shape[0], _p(idx), _p(dist)))
        return idx, dist

    def knnMatch2_device(self, d_query, nq, d_train, nt, d_idx, d_dist):
        _check(self.L.orbx_knn2_device(self.h, C.c_void_p(d_query), int(nq), C.c_void_p(d_train), int(nt), C.c_void_p(d_idx),
                                       C.c_void_p(d_dist)))

    def knnMatch2_device_parts(self, d_query, nq, part_ptrs, part_rows, d_idx, d_dist):
        """2-NN against a train set given as parts in this GPU's or in peer GPUs' memory (read in place over NVLink)."""
        n = len(part_ptrs)
        ptrs = (C.c_void_p * n)(*[int(x) for x in part_ptrs])
        rows = np.ascontiguousarray(part_rows, np.int32)
        _check(self.L.orbx_knn2_device_parts(self.h, C.c_void_p(d_query), int(nq), ptrs, _p(rows), n, C.c_void_p(d_idx), C.c_void_p(d_dist)))

    def knnMatch2_device_parts_sync(self, d_query, nq, part_ptrs, part_rows, self_index, flag_ptrs, epoch, d_idx, d_dist, timeout_s=10.0):
        """Sharded 2-NN with the barrier on the device: waits (inside the pull kernel, on an auxiliary stream) until every peer has
        published `epoch` in its flag word, pulls the peers' rows over NVLink while the main stream matches the own part
        (`self_index`) in place, then matches the pulled parts and merges.  No collective, no host synchronisation."""
        n = len(part_ptrs)
        ptrs = (C.c_void_p * n)(*[int(x) for x in part_ptrs])
        flags = (C.c_void_p * n)(*[int(x) for x in flag_ptrs])
        rows = np.ascontiguousarray(part_rows, np.int32)
        self.L.orbx_knn2_device_parts_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                       C.c_uint32, C.c_double, C.c_void_p, C.c_void_p]
        _check(self.L.orbx_knn2_device_parts_sync(self.h, C.c_void_p(d_query), int(nq), ptrs, _p(rows), n, int(self_index), flags,
                                                  int(epoch) & 0xFFFFFFFF, float(timeout_s), C.c_void_p(d_idx), C.c_void_p(d_dist)))

    @staticmethod
    def ratio_test(dist, ratio):
        dist = np.ascontiguousarray(dist, np.int32)
        keep = np.zeros(dist.shape[0], np.uint8)
        _check(lib().orbx_ratio_test(_p(dist), dist.shape[0], C.c_double(ratio), _p(keep)))
        return keep.astype(bool)

    def match_csr(self, query, train, row_ptr, cand):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        row_ptr = np.ascontiguousarray(row_ptr, np.int32); cand = np.ascontiguousarray(cand, np.int32)
        nq = q.shape[0]
        o = [np.zeros(nq, np.int32) for _ in range(5)]
        _check(self.L.orbx_match_csr(self.h, _p(q), nq, _p(t), t.shape[0], _p(row_ptr), _p(cand), *[_p(x) for x in o]))
        return dict(best_idx=o[0], best_dist=o[1], second_dist=o[2], best_pos=o[3], second_pos=o[4])

    def match_csr_dists(self, query, train, row_ptr, cand):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        row_ptr = np.ascontiguousarray(row_ptr, np.int32); cand = np.ascontiguousarray(cand, np.int32)
        d = np.zeros(max(int(row_ptr[-1]), 1), np.int32)
        _check(self.L.orbx_match_csr_dists(self.h, _p(q), q.shape[0], _p(t), t.shape[0], _p(row_ptr), _p(cand), _p(d)))
        return d[:int(row_ptr[-1])]

    def SearchForTriangulation(self, fv1, fv2, desc1, desc2, skip1=None, skip2=None):
        """Hamming core of src/ORBmatcher.cc:1131-1456 (bCoarse semantics).  fv = (nodes, ptr, idx) CSR of a
        DBoW2 FeatureVector.  Returns (nmatches, vMatchedPairs[(idx1, idx2)...] ascending idx1)."""
        n1n, p1, i1 = (np.ascontiguousarray(x, np.int32) for x in fv1)
        n2n, p2, i2 = (np.ascontiguousarray(x, np.int32) for x in fv2)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        s1 = np.zeros(d1.shape[0], np.uint8) if skip1 is None else np.ascontiguousarray(skip1, np.uint8)
        s2 = np.zeros(d2.shape[0], np.uint8) if skip2 is None else np.ascontiguousarray(skip2, np.uint8)
        m12 = np.full(d1.shape[0], -1, np.int32)
        nm = C.c_int(0)
        _check(self.L.orbx_search_for_triangulation(self.h, n1n.size, _p(n1n), _p(p1), _p(i1), n2n.size, _p(n2n), _p(p2), _p(i2),
                                                    _p(d1), d1.shape[0], _p(d2), d2.shape[0], _p(s1), _p(s2), _p(m12),
                                                    C.byref(nm)))
        ii = np.flatnonzero(m12 >= 0)
        return nm.value, list(zip(ii.tolist(), m12[ii].tolist())), m12

    def SearchForTriangulationFull(self, kf1, kf2, ep, scale_factors2, sigma2_1, sigma2_2, pairs=None, only_stereo=False, coarse=False,
                                   check_orientation=False):
        """ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), src/ORBmatcher.cc:1131-1456, the whole
        function (orbx_search_for_triangulation_full).  kf = dict(cam_off, kps, desc, has_mp, u_right (or None), fv = (nodes, ptr, idx));
        pairs = epipolar_pairs() filled for the rig (None only with coarse).  Returns (nmatches, [(idx1, idx2)...], match12)."""
        def side(kf):
            off = np.ascontiguousarray(kf['cam_off'], np.int32)
            k = np.ascontiguousarray(kf['kps'], KP_DTYPE); d = np.ascontiguousarray(kf['desc'], np.uint8).reshape(-1, 32)
            mp = np.zeros(len(k), np.uint8) if kf.get('has_mp') is None else np.ascontiguousarray(kf['has_mp'], np.uint8)
            ur = None if kf.get('u_right') is None else np.ascontiguousarray(kf['u_right'], np.float32)
            nodes, ptr, idx = (np.ascontiguousarray(x, np.int32) for x in kf['fv'])
            return off, k, d, mp, ur, nodes, ptr, idx
        o1, k1, d1, mp1, ur1, n1n, p1, i1 = side(kf1)
        o2, k2, d2, mp2, ur2, n2n, p2, i2 = side(kf2)
        ncam = o1.size - 1
        ep = np.ascontiguousarray(ep, np.float32); sf2 = np.ascontiguousarray(scale_factors2, np.float32)
        s1 = np.ascontiguousarray(sigma2_1, np.float32); s2 = np.ascontiguousarray(sigma2_2, np.float32)
        pr = None if pairs is None else np.ascontiguousarray(pairs, EPIPOLAR_PAIR_DTYPE)
        m12 = np.full(len(k1), -1, np.int32)
        nm = C.c_int(0)
        _check(self.L.orbx_search_for_triangulation_full(self.h, ncam, _p(o1), _p(k1), _p(d1), _p(mp1), _p(ur1), n1n.size, _p(n1n), _p(p1), _p(i1),
                                                         _p(o2), _p(k2), _p(d2), _p(mp2), _p(ur2), n2n.size, _p(n2n), _p(p2), _p(i2), _p(ep), _p(sf2),
                                                         _p(s1), _p(s2), sf2.size, int(only_stereo), int(coarse), int(check_orientation), _p(pr),
                                                         _p(m12), C.byref(nm)))
        i1 = np.flatnonzero(m12 >= 0)
        return nm.value, list(zip(i1.tolist(), m12[i1].tolist())), m12

    def window_best(self, kps, desc, bounds, qx, qy, radius, min_level, max_level, q_desc, valid=None, skip=None, u_right=None,
                    q_ur=None, inv_level_sigma2=None):
        """Search loop of Fuse (src/ORBmatcher.cc:1549-1601, 1719-1738) / SearchBySim3 (:1837-1861) over one KeyFrame camera: window,
        level gate, optional reprojection gate, first minimum.  Returns (best_idx[nq], best_dist[nq])."""
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(bounds, np.float32)
        qx, qy, radius = (np.ascontiguousarray(a, np.float32) for a in (qx, qy, radius))
        nq = qx.size
        mn = np.ascontiguousarray(min_level, np.int32); mx = np.ascontiguousarray(max_level, np.int32)
        qd = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        va = np.ones(nq, np.uint8) if valid is None else np.ascontiguousarray(valid, np.uint8)
        sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        qur = None if q_ur is None else np.ascontiguousarray(q_ur, np.float32)
        isg = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        bi = np.full(nq, -1, np.int32); bd = np.full(nq, 256, np.int32)
        _check(self.L.orbx_window_best(self.h, _p(kps), _p(desc), kps.size, _p(b), _p(sk), _p(ur), _p(isg), 0 if isg is None else isg.size, nq,
                                       _p(va), _p(qx), _p(qy), _p(qur), _p(radius), _p(mn), _p(mx), _p(qd), _p(bi), _p(bd)))
        return bi, bd

    def SearchBySim3(self, kf1, kf2, side1, side2):
        """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1771-1983).  kf = dict(kps, desc, bounds); side = dict(valid, u, v, radius,
        level, desc) — the KeyFrame's map points projected into the other KeyFrame.  Returns (nFound, match12[n1])."""
        k1 = np.ascontiguousarray(kf1['kps'], KP_DTYPE); k2 = np.ascontiguousarray(kf2['kps'], KP_DTYPE)
        d1 = np.ascontiguousarray(kf1['desc'], np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(kf2['desc'], np.uint8).reshape(-1, 32)
        b1 = np.ascontiguousarray(kf1['bounds'], np.float32); b2 = np.ascontiguousarray(kf2['bounds'], np.float32)

        def side(sd):
            return (np.ascontiguousarray(sd['valid'], np.uint8), np.ascontiguousarray(sd['u'], np.float32), np.ascontiguousarray(sd['v'], np.float32),
                    np.ascontiguousarray(sd['radius'], np.float32), np.ascontiguousarray(sd['level'], np.int32),
                    np.ascontiguousarray(sd['desc'], np.uint8).reshape(-1, 32))
        s1, s2 = side(side1), side(side2)
        m12 = np.full(max(k1.size, 1), -1, np.int32); nf = C.c_int(0)
        _check(self.L.orbx_search_by_sim3(self.h, _p(k1), _p(d1), k1.size, _p(b1), _p(k2), _p(d2), k2.size, _p(b2), *[_p(a) for a in s1],
                                          *[_p(a) for a in s2], _p(m12), C.byref(nf)))
        return nf.value, m12[:k1.size]

    def SearchForInitialization(self, kps1, desc1, kps2, desc2, bounds2, prev_matched, windowSize=10):
        """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:895-1004), whole function.  prev_matched [n1,2] is updated in
        place like vbPrevMatched.  Returns (nmatches, vnMatches12[n1])."""
        k1 = np.ascontiguousarray(kps1, KP_DTYPE); k2 = np.ascontiguousarray(kps2, KP_DTYPE)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        b2 = np.ascontiguousarray(bounds2, np.float32)
        if not (isinstance(prev_matched, np.ndarray) and prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous):
            raise TypeError('prev_matched must be a C-contiguous float32 array (updated in place)')
        m12 = np.full(max(k1.size, 1), -1, np.int32); nm = C.c_int(0)
        _check(self.L.orbx_search_for_initialization(self.h, _p(k1), _p(d1), k1.size, _p(k2), _p(d2), k2.size, _p(b2), _p(prev_matched),
                                                     int(windowSize), C.c_float(self.mfNNratio), int(self.mbCheckOrientation), _p(m12),
                                                     C.byref(nm)))
        return nm.value, m12[:k1.size]

    def SearchByBoW(self, fv_kf, fv_f, desc_kf, desc_f, skip_kf=None, angle_kf=None, angle_f=None, cam_off_f=None):
        """ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches) (src/ORBmatcher.cc:349-666), whole function.
        fv = (nodes, ptr, idx) CSR of a DBoW2 FeatureVector; cam_off_f = None (F.Nleft == -1) or the 3 / 5 camera block offsets.
        Returns (nmatches, match_f[n_f]) — match_f[i] = KeyFrame feature whose MapPoint goes to vpMapPointMatches[i]."""
        n1n, p1, i1 = (np.ascontiguousarray(x, np.int32) for x in fv_kf)
        n2n, p2, i2 = (np.ascontiguousarray(x, np.int32) for x in fv_f)
        d1 = np.ascontiguousarray(desc_kf, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc_f, np.uint8).reshape(-1, 32)
        s1 = None if skip_kf is None else np.ascontiguousarray(skip_kf, np.uint8)
        a1 = None if angle_kf is None else np.ascontiguousarray(angle_kf, np.float32)
        a2 = None if angle_f is None else np.ascontiguousarray(angle_f, np.float32)
        co = None if cam_off_f is None else np.ascontiguousarray(cam_off_f, np.int32)
        ncam = 1 if co is None else co.size - 1
        match = np.full(max(d2.shape[0], 1), -1, np.int32); nm = C.c_int(0)
        _check(self.L.orbx_search_by_bow_frame(self.h, n1n.size, _p(n1n), _p(p1), _p(i1), n2n.size, _p(n2n), _p(p2), _p(i2), _p(d1),
                                               d1.shape[0], _p(d2), d2.shape[0], _p(s1), _p(a1), _p(a2), ncam, _p(co),
                                               C.c_float(self.mfNNratio), int(self.mbCheckOrientation), _p(match), C.byref(nm)))
        return nm.value, match[:d2.shape[0]]

    def SearchByBoWKeyFrames(self, fv1, fv2, desc1, desc2, skip1=None, skip2=None, angle1=None, angle2=None):
        """ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:1006-1129), whole function.
        Returns (nmatches, match12[n1]) — match12[idx1] = idx2 whose MapPoint goes to vpMatches12[idx1]."""
        n1n, p1, i1 = (np.ascontiguousarray(x, np.int32) for x in fv1)
        n2n, p2, i2 = (np.ascontiguousarray(x, np.int32) for x in fv2)
        d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        s1 = None if skip1 is None else np.ascontiguousarray(skip1, np.uint8)
        s2 = None if skip2 is None else np.ascontiguousarray(skip2, np.uint8)
        a1 = None if angle1 is None else np.ascontiguousarray(angle1, np.float32)
        a2 = None if angle2 is None else np.ascontiguousarray(angle2, np.float32)
        match = np.full(max(d1.shape[0], 1), -1, np.int32); nm = C.c_int(0)
        _check(self.L.orbx_search_by_bow_keyframes(self.h, n1n.size, _p(n1n), _p(p1), _p(i1), n2n.size, _p(n2n), _p(p2), _p(i2), _p(d1),
                                                   d1.shape[0], _p(d2), d2.shape[0], _p(s1), _p(s2), _p(a1), _p(a2),
                                                   C.c_float(self.mfNNratio), int(self.mbCheckOrientation), _p(match), C.byref(nm)))
        return nm.value, match[:d1.shape[0]]

    def ComputeDistinctiveDescriptors(self, desc, ptr):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483) for many map points: desc[total,32], ptr[ngroups+1]
        -> best index inside each group (-1 for an empty group)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); ptr = np.ascontiguousarray(ptr, np.int32)
        best = np.full(ptr.size - 1, -1, np.int32)
        _check(self.L.orbx_distinctive_descriptors(self.h, _p(desc), _p(ptr), ptr.size - 1, _p(best)))
        return best

    def TriangulateMatches(self, rig, xy1, xy2, sigma1, sigma2):
        """KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:319-388) for n keypoint pairs.
        rig = kb8_stereo(...); returns (depth_or_code[n], p3d[n,3])."""
        xy1 = np.ascontiguousarray(xy1, np.float32).reshape(-1, 2); xy2 = np.ascontiguousarray(xy2, np.float32).reshape(-1, 2)
        n = xy1.shape[0]
        s1 = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma1, np.float32), (n,)))
        s2 = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma2, np.float32), (n,)))
        depth = np.zeros(n, np.float32); p3d = np.zeros((n, 3), np.float32)
        _check(self.L.orbx_kb8_triangulate_matches(self.h, _p(rig), _p(xy1), _p(xy2), _p(s1), _p(s2), n, _p(depth), _p(p3d)))
        return depth, p3d

    def ComputeStereoFishEyeMatches(self, rig, kps_left, desc_left, mono_left, kps_right, desc_right, mono_right, level_sigma2,
                                    ratio=0.7):
        """Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1409-1459; ratio 0.7) / ComputeMultiFishEyeMatches (:1461-1525; ratio
        0.8): 2-NN of the lapping blocks + ratio test + TriangulateMatches on the device.
        Returns dict(nMatches, descMatches, mvLeftToRightMatch, mvRightToLeftMatch, mvDepth, mvStereo3Dpoints)."""
        kl = np.ascontiguousarray(kps_left, KP_DTYPE); kr = np.ascontiguousarray(kps_right, KP_DTYPE)
        dl = np.ascontiguousarray(desc_left, np.uint8).reshape(-1, 32); dr = np.ascontiguousarray(desc_right, np.uint8).reshape(-1, 32)
        sg = np.ascontiguousarray(level_sigma2, np.float32)
        l2r = np.full(kl.size, -1, np.int32); r2l = np.full(kr.size, -1, np.int32)
        depth = np.full(kl.size, -1, np.float32); p3d = np.zeros((kl.size, 3), np.float32); cnt = np.zeros(2, np.int32)
        _check(self.L.orbx_stereo_fisheye_matches(self.h, _p(rig), _p(kl), _p(dl), kl.size, int(mono_left), _p(kr), _p(dr), kr.size,
                                                  int(mono_right), _p(sg), sg.size, C.c_double(ratio), _p(l2r), _p(r2l), _p(depth),
                                                  _p(p3d), _p(cnt)))
        return dict(nMatches=int(cnt[0]), descMatches=int(cnt[1]), mvLeftToRightMatch=l2r, mvRightToLeftMatch=r2l, mvDepth=depth,
                    mvStereo3Dpoints=p3d)

    def grid_candidates(self, kps, bounds, qx, qy, qr, min_level, max_level, cand_cap=None):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        qx, qy, qr = (np.ascontiguousarray(a, np.float32) for a in (qx, qy, qr))
        mn, mx = np.ascontiguousarray(min_level, np.int32), np.ascontiguousarray(max_level, np.int32)
        nq = qx.size
        cap = int(cand_cap) if cand_cap else max(64 * nq, 1024)
        row_ptr = np.zeros(nq + 1, np.int32); cand = np.zeros(cap, np.int32); n = C.c_int(0)
        _check(self.L.orbx_grid_candidates(self.h, _p(kps), kps.size, C.c_float(bounds[0]), C.c_float(bounds[1]),
                                           C.c_float(bounds[2]), C.c_float(bounds[3]), _p(qx), _p(qy), _p(qr), _p(mn), _p(mx), nq,
                                           _p(row_ptr), _p(cand), cap, C.byref(n)))
        return row_ptr, cand[:n.value]

    def SearchByProjection(self, frame, points, th=3.0, bFarPoints=False, thFarPoints=50.0):
        """src/ORBmatcher.cc:23-340 over plain arrays.
        frame : dict(cam_off[ncam+1], kps, desc, bounds[ncam,4], scale_factors, occupied?, l2r?, r2l?, u_right? (mvuRight, ncam == 1))
        points: dict(in_view[n,4], proj_x[n,4], proj_y[n,4], level[n,4], view_cos[n,4], desc[n,32], depth?, bad?, has_obs?,
                     proj_xr? (mTrackProjXR, with frame['u_right']: the rectified-stereo gate of :81-85))
        Returns (nmatches, assigned[N])  — assigned[idx] = index of the map point written to F.mvpMapPoints[idx]."""
        cam_off = np.ascontiguousarray(frame['cam_off'], np.int32)
        ncam = cam_off.size - 1
        N = int(cam_off[-1])
        kps = np.ascontiguousarray(frame['kps'], KP_DTYPE); desc = np.ascontiguousarray(frame['desc'], np.uint8)
        bounds = np.ascontiguousarray(frame['bounds'], np.float32).reshape(ncam, 4)
        sf = np.ascontiguousarray(frame['scale_factors'], np.float32)
        occ = None if frame.get('occupied') is None else np.ascontiguousarray(frame['occupied'], np.uint8)
        l2r = None if frame.get('l2r') is None else np.ascontiguousarray(frame['l2r'], np.int32)
        r2l = None if frame.get('r2l') is None else np.ascontiguousarray(frame['r2l'], np.int32)
        nmp = int(np.asarray(points['in_view']).shape[0])
        iv = np.ascontiguousarray(points['in_view'], np.uint8).reshape(nmp, 4)
        px = np.ascontiguousarray(points['proj_x'], np.float32).reshape(nmp, 4)
        py = np.ascontiguousarray(points['proj_y'], np.float32).reshape(nmp, 4)
        lv = np.ascontiguousarray(points['level'], np.int32).reshape(nmp, 4)
        vc = np.ascontiguousarray(points['view_cos'], np.float32).reshape(nmp, 4)
        md = np.ascontiguousarray(points['desc'], np.uint8).reshape(nmp, 32)
        depth = None if points.get('depth') is None else np.ascontiguousarray(points['depth'], np.float32)
        bad = None if points.get('bad') is None else np.ascontiguousarray(points['bad'], np.uint8)
        ho = None if points.get('has_obs') is None else np.ascontiguousarray(points['has_obs'], np.uint8)
        ur = None if frame.get('u_right') is None else np.ascontiguousarray(frame['u_right'], np.float32)
        pxr = None if points.get('proj_xr') is None else np.ascontiguousarray(points['proj_xr'], np.float32)
        assigned = np.full(max(N, 1), -1, np.int32); nm = C.c_int(0)
        _check(self.L.orbx_search_by_projection(self.h, ncam, _p(cam_off), _p(kps), _p(desc), _p(bounds), _p(sf), sf.size, _p(occ),
                                                _p(l2r), _p(r2l), _p(ur), nmp, _p(iv), _p(px), _p(py), _p(lv), _p(vc), _p(pxr), _p(depth), _p(bad),
                                                _p(ho), _p(md), C.c_float(th), int(bool(bFarPoints)), C.c_float(thFarPoints),
                                                C.c_float(self.mfNNratio), _p(assigned), C.byref(nm)))
        return nm.value, assigned[:N]


    def SearchByProjectionLast(self, frame, last, th, forward=False, backward=False, max_dist=TH_HIGH, level_gate=None):
        """ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (src/ORBmatcher.cc:1985-2413).
        frame: as in SearchByProjection (+ optional u_right[N] for the rectified-stereo layout, ncam == 1)
        last : dict(valid[n], proj_x[n,4], proj_y[n,4], octave[n], angle[n], desc[n,32], has_obs?, proj_ur?)
        forward / backward = bForward / bBackward.  The relocalisation overload SearchByProjection(Frame, KeyFrame, sAlreadyFound, th,
        ORBdist) (:2415-2535) is the same call with max_dist = ORBdist, SearchByProjection(KeyFrame, Sim3, ...) (:668-893) with
        level_gate = GATE_BELOW and checkOri off (see include/orbx.h).  Returns (nmatches, assigned[N])."""
        gate = level_gate if level_gate is not None else (GATE_FORWARD if forward else GATE_BACKWARD if backward else GATE_AROUND)
        cam_off = np.ascontiguousarray(frame['cam_off'], np.int32)
        ncam = cam_off.size - 1
        N = int(cam_off[-1])
        kps = np.ascontiguousarray(frame['kps'], KP_DTYPE); desc = np.ascontiguousarray(frame['desc'], np.uint8)
        bounds = np.ascontiguousarray(frame['bounds'], np.float32).reshape(ncam, 4)
        sf = np.ascontiguousarray(frame['scale_factors'], np.float32)
        occ = None if frame.get('occupied') is None else np.ascontiguousarray(frame['occupied'], np.uint8)
        ur = None if frame.get('u_right') is None else np.ascontiguousarray(frame['u_right'], np.float32)
        n = int(np.asarray(last['valid']).shape[0])
        valid = np.ascontiguousarray(last['valid'], np.uint8)
        px = np.ascontiguousarray(last['proj_x'], np.float32).reshape(n, 4); py = np.ascontiguousarray(last['proj_y'], np.float32).reshape(n, 4)
        octv = np.ascontiguousarray(last['octave'], np.int32); ang = np.ascontiguousarray(last['angle'], np.float32)
        pd = np.ascontiguousarray(last['desc'], np.uint8).reshape(n, 32)
        ho = None if last.get('has_obs') is None else np.ascontiguousarray(last['has_obs'], np.uint8)
        pur = None if last.get('proj_ur') is None else np.ascontiguousarray(last['proj_ur'], np.float32)
        assigned = np.full(max(N, 1), -1, np.int32); nm = C.c_int(0)
        _check(self.L.orbx_search_by_projection_last(self.h, ncam, _p(cam_off), _p(kps), _p(desc), _p(bounds), _p(sf), sf.size, _p(occ),
                                                     _p(ur), n, _p(valid), _p(px), _p(py), _p(pur), _p(octv), _p(ang), _p(ho), _p(pd),
                                                     C.c_float(th), int(gate), int(self.mbCheckOrientation), int(max_dist), _p(assigned), C.byref(nm)))
        self._last_n = N
        return nm.value, assigned[:N]

    def SearchByProjectionLastCleared(self):
        """After SearchByProjectionLast: cleared[N] = 1 where a claim was made and then cleared by the rotation filter (the reference writes
        NULL into CurrentFrame.mvpMapPoints there, src/ORBmatcher.cc:2404; `assigned` is -1 both there and at keypoints nobody claimed)."""
        n = int(getattr(self, '_last_n', 0))
        cleared = np.zeros(max(n, 1), np.uint8)
        _check(self.L.orbx_search_by_projection_last_cleared(self.h, _p(cleared), n))
        return cleared[:n]

class _SbpCall(C.Structure):
    _fields_ = [('ncam', C.c_int), ('cam_off', C.c_void_p), ('kps', C.c_void_p), ('desc', C.c_void_p), ('grid_bounds', C.c_void_p),
                ('scale_factors', C.c_void_p), ('nlevels', C.c_int), ('occupied', C.c_void_p), ('l2r', C.c_void_p), ('r2l', C.c_void_p),
                ('u_right', C.c_void_p), ('nmp', C.c_int), ('in_view', C.c_void_p), ('proj_x', C.c_void_p), ('proj_y', C.c_void_p),
                ('level', C.c_void_p), ('view_cos', C.c_void_p), ('proj_xr', C.c_void_p), ('depth', C.c_void_p), ('bad', C.c_void_p),
                ('has_obs', C.c_void_p), ('mp_desc', C.c_void_p), ('th', C.c_float), ('far_points', C.c_int), ('th_far_points', C.c_float),
                ('nnratio', C.c_float), ('assigned', C.c_void_p), ('nmatches', C.c_int)]


class _SbpLastCall(C.Structure):
    _fields_ = [('ncam', C.c_int), ('cam_off', C.c_void_p), ('kps', C.c_void_p), ('desc', C.c_void_p), ('grid_bounds', C.c_void_p),
                ('scale_factors', C.c_void_p), ('nlevels', C.c_int), ('occupied', C.c_void_p), ('u_right', C.c_void_p), ('npts', C.c_int),
                ('valid', C.c_void_p), ('proj_x', C.c_void_p), ('proj_y', C.c_void_p), ('proj_ur', C.c_void_p), ('last_octave', C.c_void_p),
                ('last_angle', C.c_void_p), ('has_obs', C.c_void_p), ('pt_desc', C.c_void_p), ('th', C.c_float), ('level_gate', C.c_int),
                ('check_orientation', C.c_int), ('max_dist', C.c_int), ('assigned', C.c_void_p), ('nmatches', C.c_int)]


class _BowCall(C.Structure):
    _fields_ = [('nn_kf', C.c_int), ('nodes_kf', C.c_void_p), ('ptr_kf', C.c_void_p), ('idx_kf', C.c_void_p), ('nn_f', C.c_int),
                ('nodes_f', C.c_void_p), ('ptr_f', C.c_void_p), ('idx_f', C.c_void_p), ('desc_kf', C.c_void_p), ('n_kf', C.c_int),
                ('desc_f', C.c_void_p), ('n_f', C.c_int), ('skip_kf', C.c_void_p), ('angle_kf', C.c_void_p), ('angle_f', C.c_void_p),
                ('ncam_f', C.c_int), ('cam_off_f', C.c_void_p), ('nnratio', C.c_float), ('check_orientation', C.c_int),
                ('match_f', C.c_void_p), ('nmatches', C.c_int)]


def _addr(a):
    return None if a is None else a.ctypes.data


class MatcherPool:
    """orbx_matcher_pool (include/orbx.h): the guided matchers of MANY frames at once — `lanes` matcher handles (stream + buffers) and
    host workers; the device runs the frames' single-CTA resolvers side by side.  Each method takes lists of the per-frame arguments of
    the corresponding ORBmatcher method and returns the per-frame results."""

    def __init__(self, lanes=16, device=0):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.orbx_matcher_pool_create(C.byref(h), int(device), int(lanes)))
        self.h = h
        self.L.orbx_matcher_pool_destroy.argtypes = [C.c_void_p]

    def close(self):
        if getattr(self, 'h', None):
            self.L.orbx_matcher_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SearchByProjectionBatch(self, frames, points, th=3.0, bFarPoints=False, thFarPoints=50.0, nnratio=0.8):
        n = len(frames)
        calls = (_SbpCall * n)(); status = (C.c_int * n)(); keep = []; outs = []
        for i, (f, pt) in enumerate(zip(frames, points)):
            cam_off = np.ascontiguousarray(f['cam_off'], np.int32); ncam = cam_off.size - 1; N = int(cam_off[-1])
            nmp = int(np.asarray(pt['in_view']).shape[0])
            a = dict(cam_off=cam_off, kps=np.ascontiguousarray(f['kps'], KP_DTYPE), desc=np.ascontiguousarray(f['desc'], np.uint8),
                     grid_bounds=np.ascontiguousarray(f['bounds'], np.float32), scale_factors=np.ascontiguousarray(f['scale_factors'], np.float32),
                     in_view=np.ascontiguousarray(pt['in_view'], np.uint8), proj_x=np.ascontiguousarray(pt['proj_x'], np.float32),
                     proj_y=np.ascontiguousarray(pt['proj_y'], np.float32), level=np.ascontiguousarray(pt['level'], np.int32),
                     view_cos=np.ascontiguousarray(pt['view_cos'], np.float32), mp_desc=np.ascontiguousarray(pt['desc'], np.uint8))
            for k_, src, dt in (('occupied', f, np.uint8), ('l2r', f, np.int32), ('r2l', f, np.int32), ('u_right', f, np.float32),
                                ('proj_xr', pt, np.float32), ('depth', pt, np.float32), ('bad', pt, np.uint8), ('has_obs', pt, np.uint8)):
                a[k_] = None if src.get(k_) is None else np.ascontiguousarray(src[k_], dt)
            out = np.full(max(N, 1), -1, np.int32)
            keep.append(a); outs.append((out, N))
            c = calls[i]
            c.ncam, c.nlevels, c.nmp = ncam, a['scale_factors'].size, nmp
            for k_ in ('cam_off', 'kps', 'desc', 'grid_bounds', 'scale_factors', 'occupied', 'l2r', 'r2l', 'u_right', 'in_view', 'proj_x', 'proj_y', 'level',
                       'view_cos', 'proj_xr', 'depth', 'bad', 'has_obs', 'mp_desc'):
                setattr(c, k_, _addr(a[k_]))
            c.th, c.far_points, c.th_far_points, c.nnratio, c.assigned = th, int(bool(bFarPoints)), thFarPoints, nnratio, out.ctypes.data
        _check(self.L.orbx_search_by_projection_batch(self.h, n, calls, status))
        for i in range(n):
            _check(status[i])
        return [(calls[i].nmatches, outs[i][0][:outs[i][1]]) for i in range(n)]

    def SearchByProjectionLastBatch(self, frames, lasts, th, level_gate=None, check_orientation=True, max_dist=TH_HIGH):
        n = len(frames)
        calls = (_SbpLastCall * n)(); status = (C.c_int * n)(); keep = []; outs = []
        for i, (f, la) in enumerate(zip(frames, lasts)):
            cam_off = np.ascontiguousarray(f['cam_off'], np.int32); ncam = cam_off.size - 1; N = int(cam_off[-1])
            a = dict(cam_off=cam_off, kps=np.ascontiguousarray(f['kps'], KP_DTYPE), desc=np.ascontiguousarray(f['desc'], np.uint8),
                     grid_bounds=np.ascontiguousarray(f['bounds'], np.float32), scale_factors=np.ascontiguousarray(f['scale_factors'], np.float32),
                     valid=np.ascontiguousarray(la['valid'], np.uint8), proj_x=np.ascontiguousarray(la['proj_x'], np.float32),
                     proj_y=np.ascontiguousarray(la['proj_y'], np.float32), last_octave=np.ascontiguousarray(la['octave'], np.int32),
                     last_angle=np.ascontiguousarray(la['angle'], np.float32), pt_desc=np.ascontiguousarray(la['desc'], np.uint8))
            for k_, src, key, dt in (('occupied', f, 'occupied', np.uint8), ('u_right', f, 'u_right', np.float32), ('proj_ur', la, 'proj_ur', np.float32),
                                     ('has_obs', la, 'has_obs', np.uint8)):
                a[k_] = None if src.get(key) is None else np.ascontiguousarray(src[key], dt)
            out = np.full(max(N, 1), -1, np.int32)
            keep.append(a); outs.append((out, N))
            c = calls[i]
            c.ncam, c.nlevels, c.npts = ncam, a['scale_factors'].size, int(a['valid'].size)
            for k_ in ('cam_off', 'kps', 'desc', 'grid_bounds', 'scale_factors', 'occupied', 'u_right', 'valid', 'proj_x', 'proj_y', 'proj_ur', 'last_octave',
                       'last_angle', 'has_obs', 'pt_desc'):
                setattr(c, k_, _addr(a[k_]))
            c.th, c.level_gate, c.check_orientation, c.max_dist, c.assigned = th, int(GATE_AROUND if level_gate is None else level_gate), int(check_orientation), int(max_dist), out.ctypes.data
        _check(self.L.orbx_search_by_projection_last_batch(self.h, n, calls, status))
        for i in range(n):
            _check(status[i])
        return [(calls[i].nmatches, outs[i][0][:outs[i][1]]) for i in range(n)]

    def SearchByBoWBatch(self, fv_kfs, fv_fs, desc_kfs, desc_fs, skip_kfs, angle_kfs, angle_fs, cam_off_fs, nnratio=0.7, check_orientation=True):
        n = len(fv_kfs)
        calls = (_BowCall * n)(); status = (C.c_int * n)(); keep = []; outs = []
        for i in range(n):
            nk, pk, ik = (np.ascontiguousarray(x, np.int32) for x in fv_kfs[i]); nf, pf, iff = (np.ascontiguousarray(x, np.int32) for x in fv_fs[i])
            dk = np.ascontiguousarray(desc_kfs[i], np.uint8).reshape(-1, 32); df = np.ascontiguousarray(desc_fs[i], np.uint8).reshape(-1, 32)
            sk = np.zeros(dk.shape[0], np.uint8) if skip_kfs[i] is None else np.ascontiguousarray(skip_kfs[i], np.uint8)
            ak = np.ascontiguousarray(angle_kfs[i], np.float32); af = np.ascontiguousarray(angle_fs[i], np.float32)
            co = np.ascontiguousarray(cam_off_fs[i], np.int32)
            out = np.full(max(df.shape[0], 1), -1, np.int32)
            keep.append((nk, pk, ik, nf, pf, iff, dk, df, sk, ak, af, co)); outs.append((out, df.shape[0]))
            c = calls[i]
            c.nn_kf, c.nodes_kf, c.ptr_kf, c.idx_kf = nk.size, nk.ctypes.data, pk.ctypes.data, ik.ctypes.data
            c.nn_f, c.nodes_f, c.ptr_f, c.idx_f = nf.size, nf.ctypes.data, pf.ctypes.data, iff.ctypes.data
            c.desc_kf, c.n_kf, c.desc_f, c.n_f = dk.ctypes.data, dk.shape[0], df.ctypes.data, df.shape[0]
            c.skip_kf, c.angle_kf, c.angle_f, c.ncam_f, c.cam_off_f = sk.ctypes.data, ak.ctypes.data, af.ctypes.data, co.size - 1, co.ctypes.data
            c.nnratio, c.check_orientation, c.match_f = nnratio, int(check_orientation), out.ctypes.data
        _check(self.L.orbx_search_by_bow_frame_batch(self.h, n, calls, status))
        for i in range(n):
            _check(status[i])
        return [(calls[i].nmatches, outs[i][0][:outs[i][1]]) for i in range(n)]


class ORBVocabulary:
    """DBoW2 TemplatedVocabulary<FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) as flat arrays on the device; transform()
    mirrors mpORBvocabulary->transform(features, BowVector, FeatureVector, levelsup) (src/Frame.cc:980-985)."""

    def __init__(self, child_ptr, child, node_desc, word_id, weight, levels, device=0):
        self.L = lib()
        self.child_ptr = np.ascontiguousarray(child_ptr, np.int32); self.child = np.ascontiguousarray(child, np.int32)
        self.node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        self.word_id = np.ascontiguousarray(word_id, np.int32); self.weight = np.ascontiguousarray(weight, np.float64)
        self.levels = int(levels)
        h = C.c_void_p()
        _check(self.L.orbx_vocabulary_create(C.byref(h), int(device), self.node_desc.shape[0], _p(self.child_ptr), _p(self.child),
                                             _p(self.node_desc), _p(self.word_id), _p(self.weight), self.levels))
        self.h = h
        self.L.orbx_vocabulary_destroy.argtypes = [C.c_void_p]

    def close(self):
        if getattr(self, 'h', None):
            self.L.orbx_vocabulary_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, matcher, desc, levelsup=4):
        """-> (BowVector as (word ids, L1-normalised tf-idf values), FeatureVector as (node ids, ptr, feature indices), word_id[n],
        node_id[n]).  The tree descent runs on the device; the two std::map fills follow TemplatedVocabulary.h:1148-1160 in feature order."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = desc.shape[0]
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
        _check(self.L.orbx_vocabulary_transform(matcher.h, self.h, _p(desc), n, int(levelsup), _p(word), _p(node), _p(w)))
        bow = {}
        fv = {}
        for i in range(n):                       # v.addWeight(id, w); fv.addFeature(nid, i_feature)   (w > 0: not a stopped word)
            if w[i] > 0:
                bow[int(word[i])] = bow.get(int(word[i]), 0.0) + float(w[i])
                fv.setdefault(int(node[i]), []).append(i)
        ids = np.array(sorted(bow), np.int32)
        vals = np.array([bow[k] for k in ids], np.float64)
        norm = 0.0
        for x in vals:                           # BowVector::normalize(L1): sum in key order
            norm += abs(float(x))
        if norm > 0.0:
            vals = vals / norm
        nodes = np.array(sorted(fv), np.int32)
        ptr = np.concatenate([[0], np.cumsum([len(fv[k]) for k in nodes])]).astype(np.int32)
        idx = np.array([i for k in nodes for i in fv[k]], np.int32)
        return (ids, vals), (nodes, ptr, idx), word, node


def popc_peak(device=0):
    v = C.c_double(0)
    _check(lib().orbx_popc_peak(int(device), C.byref(v)))
    return v.value


class SharedRows:
    """Train descriptor rows of the sharded brute-force matcher: every rank owns TWO device buffers that the other ranks of the
    node map through CUDA IPC, so a rank's 2-NN kernels read all parts in place (peer loads over NVLink): no all-gather, no
    staging copy.  Protocol per step k (everything stream-ordered, no host synchronisation):
        fill(k & 1, my rows)  ->  one tiny all-reduce (a stream-ordered barrier)  ->  2-NN kernels read ptrs[k & 1] of all ranks.
    The double buffer makes that single barrier sufficient: a rank can only pass barrier k after every rank has finished the
    kernels of step k-1, so nobody still reads the buffer that step k+1 overwrites.
    `dist` is an initialised torch.distributed module (used to exchange the 64-byte handles and for the barrier)."""

    def __init__(self, dist, device, n_total):
        from . import shard
        self.L, self.device, self.dist = lib(), int(device), dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rows = [shard.row_shard(n_total, r, self.world)[1] for r in range(self.world)]
        self.L.orbx_shared_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
        self.L.orbx_shared_open.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        self.L.orbx_shared_close.argtypes = [C.c_int, C.c_void_p]
        self.L.orbx_shared_free.argtypes = [C.c_int, C.c_void_p]
        self.L.orbx_device_copy.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        self.L.orbx_shared_publish.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
        self.own, self.ptrs, self.flags = [], [], []
        # collectives first (allocation + handle exchange), then the purely local peer mappings: a rank whose mapping fails has
        # no collective left to skip, so the caller can agree on a fallback with one more all-reduce
        handles = []
        for i in range(3):                       # two row buffers and the epoch word (a zero-filled 256-byte buffer)
            own, handle = C.c_void_p(), C.create_string_buffer(64)
            _check(self.L.orbx_shared_alloc(self.device, max(self.rows[self.rank], 1) * 32 if i < 2 else 256, C.byref(own), handle))
            self.own.append(own.value)
            got = [None] * self.world
            dist.all_gather_object(got, handle.raw)
            handles.append(got)
        for buf in range(3):
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self.own[buf])
                else:
                    q = C.c_void_p()
                    _check(self.L.orbx_shared_open(self.device, handles[buf][r], C.byref(q)))
                    ptrs.append(q.value)
            self.ptrs.append(ptrs)
        self.flags = self.ptrs.pop()             # flags[r] = rank r's epoch word

    def fill(self, buf, d_local_rows_ptr, cuda_stream):
        """Copies this rank's rows (device pointer) into its shared buffer `buf` on `cuda_stream`."""
        _check(self.L.orbx_device_copy(self.device, C.c_void_p(self.own[buf]), C.c_void_p(d_local_rows_ptr), self.rows[self.rank] * 32,
                                       C.c_void_p(cuda_stream)))

    def publish(self, epoch, cuda_stream):
        """Stores `epoch` in this rank's flag word, ordered after everything queued on `cuda_stream` (the fill)."""
        _check(self.L.orbx_shared_publish(self.device, C.c_void_p(self.own[2]), int(epoch) & 0xFFFFFFFF, C.c_void_p(cuda_stream)))

    def close(self, barrier=True):
        for lst in self.ptrs + ([self.flags] if self.flags else []):
            for r, q in enumerate(lst):
                if r != self.rank and q:
                    self.L.orbx_shared_close(self.device, C.c_void_p(q))
        if barrier:
            self.dist.barrier()                  # nobody maps the buffers any more
        for own in self.own:
            self.L.orbx_shared_free(self.device, C.c_void_p(own))
        self.ptrs, self.own, self.flags = [], [], []
