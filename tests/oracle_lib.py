"""ctypes binding of oracle/liborb_oracle.so — the CPU oracle (checker only, never the product)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'),
                     ('octave', 'i4'), ('class_id', 'i4')])

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int)
f32p = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, 'oracle', 'liborb_oracle.so')
        if not os.path.exists(so):
            import subprocess
            subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
        _lib = C.CDLL(so)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_grid_create.restype = C.c_void_p
    return _lib


def p(a, t=None):
    """pointer to a numpy array's data (array must stay alive)"""
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """Mirrors ORB_SLAM3::ORBextractor (ORBextractor.h:28-84) on the oracle."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self.L = lib()
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.orc_create(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        nf = np.zeros(n, np.int32); um = np.zeros(16, np.int32)
        self.L.orc_tables(self.h, p(sf), p(isf), p(s2), p(is2), p(nf), p(um))
        return dict(sf=sf, inv_sf=isf, sigma2=s2, inv_sigma2=is2, nfeat=nf, umax=um)

    def __call__(self, image, lapping=(0, 0), fma_mode=0):
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2
        img = np.ascontiguousarray(image)
        hh, w = img.shape
        mono = self.L.orc_extract(self.h, p(img), w, hh, img.strides[0], int(lapping[0]), int(lapping[1]), fma_mode)
        n = self.L.orc_num_keypoints(self.h)
        kps = np.zeros(n, KP_DTYPE); desc = np.zeros((n, 32), np.uint8)
        if n:
            self.L.orc_result(self.h, p(kps), p(desc))
        return mono, kps, desc

    def level(self, l):
        w, hh, st = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(hh), C.byref(st))
        buf = np.zeros((hh.value + 38, w.value + 38), np.uint8)
        self.L.orc_level_bordered(self.h, l, p(buf))
        return buf

    def level_blur(self, l):
        w, hh, st = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(hh), C.byref(st))
        buf = np.zeros((hh.value, w.value), np.uint8)
        ok = self.L.orc_level_blur(self.h, l, p(buf))
        return buf if ok else None

    def level_candidates(self, l):
        n = self.L.orc_level_num_candidates(self.h, l)
        xs, ys, sc = (np.zeros(n, np.int32) for _ in range(3))
        if n:
            self.L.orc_level_candidates(self.h, l, p(xs), p(ys), p(sc))
        return xs, ys, sc

    def level_keypoints(self, l):
        n = self.L.orc_level_num_keypoints(self.h, l)
        kps = np.zeros(n, KP_DTYPE)
        if n:
            self.L.orc_level_keypoints(self.h, l, p(kps))
        return kps

    def stage_seconds(self, reset=True):
        t = np.zeros(6, np.float64)
        self.L.orc_stage_seconds(self.h, p(t), int(reset))
        return t


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src); dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear(p(src), src.shape[1], src.shape[0], src.strides[0], p(dst), dw, dh, dw)
    return dst


def border101(img, b):
    h, w = img.shape
    buf = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    buf[b:b + h, b:b + w] = img
    lib().orc_border_reflect101(p(buf), w, h, buf.strides[0], b)
    return buf


def fast(img, threshold):
    img = np.ascontiguousarray(img)
    cap = img.size
    xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
    n = lib().orc_fast(p(img), img.shape[1], img.shape[0], img.strides[0], threshold, p(xs), p(ys), p(sc), cap)
    return xs[:n], ys[:n], sc[:n]


def blur7(img):
    img = np.ascontiguousarray(img); dst = np.zeros_like(img)
    lib().orc_blur7(p(img), img.shape[1], img.shape[0], img.strides[0], p(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().orc_fast_atan2_n(p(y), p(x), p(out), y.size)
    return out


def distribute(xs, ys, scores, minX, maxX, minY, maxY, N):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    scores = np.ascontiguousarray(scores, np.int32)
    sel = np.zeros(max(1, xs.size), np.int32)
    n = lib().orc_distribute(p(xs), p(ys), p(scores), xs.size, minX, maxX, minY, maxY, N, p(sel), sel.size)
    return sel[:n]


def sort_nodes(sizes, ulx):
    sizes = np.ascontiguousarray(sizes, np.int32); ulx = np.ascontiguousarray(ulx, np.int32)
    perm = np.zeros(sizes.size, np.int32)
    lib().orc_sort_nodes(p(sizes), p(ulx), sizes.size, p(perm))
    return perm


def hamming(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return lib().orc_hamming(p(a), p(b))


def knn2(q, t, nthreads=8):
    q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
    idx = np.zeros((q.shape[0], 2), np.int32); dist = np.zeros((q.shape[0], 2), np.int32)
    lib().orc_knn2(p(q), q.shape[0], p(t), t.shape[0], p(idx), p(dist), nthreads)
    return idx, dist


def ratio_test(dist, ratio):
    dist = np.ascontiguousarray(dist, np.int32)
    keep = np.zeros(dist.shape[0], np.uint8)
    lib().orc_ratio_test(p(dist), dist.shape[0], C.c_double(ratio), p(keep))
    return keep.astype(bool)


def libm_sincosf(x):
    x = np.ascontiguousarray(x, np.float32)
    s = np.zeros_like(x); c = np.zeros_like(x)
    lib().orc_libm_sincosf_n(p(x), p(s), p(c), x.size)
    return s, c


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own src/ORBextractor.cc, compiled unmodified against oracle/cvshim (oracle/Makefile).
# ---------------------------------------------------------------------------------------------------------------------
_ref_libs = {}


def ref_lib(fma=False):
    """oracle/_ref/libref_orb.so (-ffp-contract=off) or libref_orb_fma.so (the reference's own flags: contraction allowed).
    Built from /root/reference when that mount exists (the build container); on the GPU box the prebuilt file must have
    travelled with the snapshot — a missing file is an error, not a skip."""
    if fma not in _ref_libs:
        so = os.path.join(ROOT, 'oracle', '_ref', 'libref_orb_fma.so' if fma else 'libref_orb.so')
        if not os.path.exists(so):
            import subprocess
            subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
        if not os.path.exists(so):
            raise RuntimeError('%s is missing and /root/reference is not mounted to build it' % so)
        lib()   # liborb_oracle.so carries the cv2-pinned primitives the shim delegates to
        L = C.CDLL(so)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_extract_batch.restype = C.c_double
        L.ref_extract_rig_frames.restype = C.c_double
        _ref_libs[fma] = L
    return _ref_libs[fma]


class RefExtractor:
    """ORB_SLAM3::ORBextractor itself (reference object code), same call shape as OracleExtractor."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, fma=False):
        self.L = ref_lib(fma)
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.ref_create(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST))

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        nf = np.zeros(n, np.int32); um = np.zeros(16, np.int32)
        self.L.ref_tables(self.h, p(sf), p(isf), p(s2), p(is2), p(nf), p(um))
        return dict(sf=sf, inv_sf=isf, sigma2=s2, inv_sigma2=is2, nfeat=nf, umax=um)

    def __call__(self, image, lapping=(0, 0)):
        if image is None or image.size == 0:
            mono = self.L.ref_extract(self.h, None, 0, 0, 0, int(lapping[0]), int(lapping[1]))
            return mono, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        hh, w = image.shape
        mono = self.L.ref_extract(self.h, p(image), w, hh, image.strides[0], int(lapping[0]), int(lapping[1]))
        n = self.L.ref_num_keypoints(self.h)
        kps = np.zeros(n, KP_DTYPE); desc = np.zeros((n, 32), np.uint8)
        if n:
            self.L.ref_result(self.h, p(kps), p(desc))
        return mono, kps, desc

    def level(self, l):
        w, hh = C.c_int(), C.c_int()
        self.L.ref_level_size(self.h, l, C.byref(w), C.byref(hh))
        buf = np.zeros((hh.value + 38, w.value + 38), np.uint8)
        self.L.ref_level_bordered(self.h, l, p(buf))
        return buf

    def stages(self, image):
        """ComputePyramid + ComputeKeyPointsOctTree; per-level keypoints (level coordinates, angles set)"""
        img = np.ascontiguousarray(image)
        self.L.ref_stages(self.h, p(img), img.shape[1], img.shape[0], img.strides[0])
        out = []
        for l in range(self.nlevels):
            n = self.L.ref_level_num_keypoints(self.h, l)
            k = np.zeros(n, KP_DTYPE)
            if n:
                self.L.ref_level_keypoints(self.h, l, p(k))
            out.append(k)
        return out

    def distribute(self, xs, ys, scores, minX, maxX, minY, maxY, N):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
        scores = np.ascontiguousarray(scores, np.int32)
        cap = max(1, xs.size)
        ox, oy, osc = (np.zeros(cap, np.int32) for _ in range(3))
        n = self.L.ref_distribute(self.h, p(xs), p(ys), p(scores), xs.size, minX, maxX, minY, maxY, N, p(ox), p(oy), p(osc), cap)
        return ox[:n], oy[:n], osc[:n]


def ref_extract_batch(imgs, nfeatures, sf, nlevels, ini, mn, lapping, nthreads, fma=True):
    """seconds, keypoints: the reference extractor over a (n, h, w) stack with `nthreads` workers (CPU baseline)"""
    imgs = np.ascontiguousarray(imgs)
    n, hh, w = imgs.shape
    nkp = C.c_longlong(0)
    dt = ref_lib(fma).ref_extract_batch(n, p(imgs), w, hh, nfeatures, C.c_float(sf), nlevels, ini, mn, int(lapping[0]), int(lapping[1]),
                                        nthreads, C.byref(nkp))
    return dt, nkp.value


def ref_extract_rig_frames(imgs, ncam, nfeatures, sf, nlevels, ini, mn, laps, fma=True):
    """seconds, keypoints: rig frames one after the other, one std::thread per camera per frame (src/Frame.cc:1841-1862)"""
    imgs = np.ascontiguousarray(imgs)
    n, hh, w = imgs.shape
    lp = np.ascontiguousarray(laps, np.int32).reshape(ncam, 2)
    nkp = C.c_longlong(0)
    dt = ref_lib(fma).ref_extract_rig_frames(n // ncam, ncam, p(imgs), w, hh, nfeatures, C.c_float(sf), nlevels, ini, mn, p(lp), C.byref(nkp))
    return dt, nkp.value
