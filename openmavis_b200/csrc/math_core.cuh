// math_core.cuh — floating-point pieces whose BITS matter, shared by device code and the CPU test harness.
//
//  * orbx_fast_atan2  : cv::fastAtan2 (OpenCV core, scalar atan_f32 polynomial), called by IC_Angle
//                       (src/ORBextractor.cc:42).  Every fp32 operation is rounded separately (no FMA).
//  * orbx_sincosf     : glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, s_sinf.c, s_cosf.c — the
//                       double-precision polynomial scheme by Szabolcs Nagy), which is what
//                       `(float)cos(angle)` / `(float)sin(angle)` resolve to in src/ORBextractor.cc:48.
//                       glibc is a third-party dependency of the reference that is not under /root/reference;
//                       the published algorithm is restated here with the table constants of libm.so.6
//                       (2.39-0ubuntu8.5) and verified EXHAUSTIVELY against that libm for every float in
//                       [0, 6.3] (tests/test_host_core.py::test_sincosf_matches_glibc_exhaustively).
//                       Only the two branches reachable for angles in [0, 2*pi] are implemented.
//  * resize taps      : the coefficient tables of cv::resize(INTER_LINEAR) for 8-bit images.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define MC_HD __host__ __device__ __forceinline__
#define MC_FMUL(a, b) __fmul_rn((a), (b))
#define MC_FADD(a, b) __fadd_rn((a), (b))
#define MC_FSUB(a, b) __fsub_rn((a), (b))
#define MC_FDIV(a, b) __fdiv_rn((a), (b))
#define MC_DMUL(a, b) __dmul_rn((a), (b))
#define MC_DADD(a, b) __dadd_rn((a), (b))
#else
#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif
// host builds use -ffp-contract=off (or -fmad=false under nvcc)
#define MC_FMUL(a, b) ((a) * (b))
#define MC_FADD(a, b) ((a) + (b))
#define MC_FSUB(a, b) ((a) - (b))
#define MC_FDIV(a, b) ((a) / (b))
#define MC_DMUL(a, b) ((a) * (b))
#define MC_DADD(a, b) ((a) + (b))
#endif

namespace orbx {

MC_HD float orbx_fast_atan2(float y, float x) {
    const float p1 = 57.2836266f, p3 = -18.6674461f, p5 = 8.91400051f, p7 = -2.53972459f;
    const float eps = 2.22044605e-16f;   // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = MC_FDIV(ay, MC_FADD(ax, eps));
        c2 = MC_FMUL(c, c);
        a = MC_FMUL(MC_FADD(MC_FMUL(MC_FADD(MC_FMUL(MC_FADD(MC_FMUL(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = MC_FDIV(ax, MC_FADD(ay, eps));
        c2 = MC_FMUL(c, c);
        a = MC_FSUB(90.f, MC_FMUL(MC_FADD(MC_FMUL(MC_FADD(MC_FMUL(MC_FADD(MC_FMUL(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = MC_FSUB(180.f, a);
    if (y < 0) a = MC_FSUB(360.f, a);
    return a;
}

// ---- glibc sinf/cosf --------------------------------------------------------------------------------------
MC_HD float orbx_sincos_poly(double x, double x2, int n, int neg) {
    // __sincosf_table[neg]: c0..c4 / s1..s3 ; neg selects the table that yields -cos
    const double c0 = neg ? -0x1p0 : 0x1p0;
    const double c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    const double c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
    const double c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    const double c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = MC_DMUL(x, x2);
        const double t1 = MC_DADD(s2, MC_DMUL(x2, s3));
        const double x7 = MC_DMUL(x3, x2);
        const double s = MC_DADD(x, MC_DMUL(x3, s1));
        return (float)MC_DADD(s, MC_DMUL(x7, t1));
    } else {
        const double x4 = MC_DMUL(x2, x2);
        const double t2 = MC_DADD(c3, MC_DMUL(x2, c4));
        const double t1 = MC_DADD(c0, MC_DMUL(x2, c1));
        const double x6 = MC_DMUL(x4, x2);
        const double c = MC_DADD(t1, MC_DMUL(x4, c2));
        return (float)MC_DADD(c, MC_DMUL(x6, t2));
    }
}

MC_HD uint32_t orbx_abstop12(float x) {
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(x);
#else
    memcpy(&u, &x, 4);
#endif
    return (u >> 20) & 0x7ff;
}

// valid for 0 <= y < 120 (the reference only produces angles in [0, 2*pi])
MC_HD void orbx_sincosf(float y, float *sn, float *cs) {
    double x = (double)y;
    if (orbx_abstop12(y) < orbx_abstop12(0x1.921FB6p-1f)) {           // |y| < pi/4
        const double x2 = MC_DMUL(x, x);
        if (orbx_abstop12(y) < orbx_abstop12(0x1p-12f)) { *sn = y; *cs = 1.0f; return; }
        *sn = orbx_sincos_poly(x, x2, 0, 0);
        *cs = orbx_sincos_poly(x, x2, 1, 0);
        return;
    }
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double r = MC_DMUL(x, hpi_inv);
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = MC_DADD(x, -MC_DMUL((double)n, hpi));
    const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
    const int neg = (n & 2) ? 1 : 0;
    const double xs = MC_DMUL(x, sg), x2 = MC_DMUL(x, x);
    *sn = orbx_sincos_poly(xs, x2, n, neg);
    *cs = orbx_sincos_poly(xs, x2, n ^ 1, neg);
}

// ---- cv::resize(INTER_LINEAR, 8U) coefficient tables (OpenCV imgproc/resize.cpp) ---------
// tap = {first source index, second source index (clamped), a0, a1} with a0/a1 in Q11.
struct alignas(8) ResizeTap { unsigned short s0, s1; short a0, a1; };   // (8-byte aligned: one 64-bit load per tap)

inline void orbx_resize_taps(int src, int dst, ResizeTap *taps) {
    const double inv_scale = (double)dst / src;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        taps[d].s0 = (unsigned short)s;
        taps[d].s1 = (unsigned short)(s + 1 < src ? s + 1 : src - 1);
        taps[d].a0 = (short)lrintf((1.f - f) * 2048.f);
        taps[d].a1 = (short)lrintf(f * 2048.f);
    }
}

MC_HD int orbx_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { i = i < 0 ? -i : 2 * (n - 1) - i; }
    return i;
}

}  // namespace orbx
