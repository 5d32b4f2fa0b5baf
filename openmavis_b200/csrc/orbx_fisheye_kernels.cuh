// orbx_fisheye_kernels.cuh — the fisheye stereo post-filter after the brute-force 2-NN:
// Frame::ComputeStereoFishEyeMatches / ComputeMultiFishEyeMatches (src/Frame.cc:1409-1459, 1461-1525) =
// Lowe ratio test on the two neighbours + KannalaBrandt8::TriangulateMatches
// (src/CameraModels/KannalaBrandt8.cpp:319-388; unproject :96-126, project :48-65, Triangulate :410-427).
//
// Arithmetic: fp32 with every operation rounded separately, in the reference's order, EXCEPT the 4x4 null vector of the
// DLT system (Eigen::JacobiSVD<Matrix4f> in the reference): it is computed here in fp64 by a one-sided Jacobi SVD and
// rounded to fp32 once, i.e. the correctly rounded answer for the fp32 matrix.  atan2f / tanf / cosf / sinf are CUDA's
// (<= 2 ulp from glibc's), so parity with the oracle is tolerance-based (tests state 1e-4 relative).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx.h"

namespace orbx {

// KannalaBrandt8::unproject (:96-126): Newton iterations on theta, ray = (pw * tan(theta) / theta_d, 1)
__device__ __forceinline__ void kb8_unproject(const float *P, float precision, float u, float v, float &rx, float &ry) {
    const float px = __fdiv_rn(__fsub_rn(u, P[2]), P[0]), py = __fdiv_rn(__fsub_rn(v, P[3]), P[1]);
    float scale = 1.0f;
    float theta_d = __fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
    theta_d = fminf(fmaxf(-1.57079632679489661923f, theta_d), 1.57079632679489661923f);
    if ((double)theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            const float t2 = __fmul_rn(theta, theta), t4 = __fmul_rn(t2, t2), t6 = __fmul_rn(t4, t2), t8 = __fmul_rn(t4, t4);
            const float k0 = __fmul_rn(P[4], t2), k1 = __fmul_rn(P[5], t4), k2 = __fmul_rn(P[6], t6), k3 = __fmul_rn(P[7], t8);
            const float num = __fsub_rn(__fmul_rn(theta, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.0f, k0), k1), k2), k3)), theta_d);
            const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.0f, __fmul_rn(3.0f, k0)), __fmul_rn(5.0f, k1)), __fmul_rn(7.0f, k2)),
                                        __fmul_rn(9.0f, k3));
            const float fix = __fdiv_rn(num, den);
            theta = __fsub_rn(theta, fix);
            if (fabsf(fix) < precision) break;
        }
        scale = __fdiv_rn(tanf(theta), theta_d);
    }
    rx = __fmul_rn(px, scale); ry = __fmul_rn(py, scale);
}

// KannalaBrandt8::project(Eigen::Vector3f) (:48-65)
__device__ __forceinline__ void kb8_project(const float *P, float x, float y, float z, float &u, float &v) {
    const float x2y2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float theta = atan2f(__fsqrt_rn(x2y2), z);
    const float psi = atan2f(y, x);
    const float t2 = __fmul_rn(theta, theta), t3 = __fmul_rn(theta, t2), t5 = __fmul_rn(t3, t2), t7 = __fmul_rn(t5, t2), t9 = __fmul_rn(t7, t2);
    const float r = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(theta, __fmul_rn(P[4], t3)), __fmul_rn(P[5], t5)), __fmul_rn(P[6], t7)), __fmul_rn(P[7], t9));
    u = __fadd_rn(__fmul_rn(__fmul_rn(P[0], r), cosf(psi)), P[2]);
    v = __fadd_rn(__fmul_rn(__fmul_rn(P[1], r), sinf(psi)), P[3]);
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// right singular vector of the smallest singular value of a 4x4 matrix (row-major a[16]): one-sided (Hestenes) Jacobi in fp64
__device__ inline void null_vector4(const float *a, double *v4) {
    double U[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { U[i][j] = (double)a[4 * i + j]; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double al = 0, be = 0, ga = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { al += U[i][p] * U[i][p]; be += U[i][q] * U[i][q]; ga += U[i][p] * U[i][q]; }
                if (ga == 0.0 || fabs(ga) <= 1e-15 * sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double up = U[i][p], uq = U[i][q];
                    U[i][p] = c * up - s * uq; U[i][q] = s * up + c * uq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    int best = 0; double bn = 1e300;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double n = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) n += U[i][j] * U[i][j];
        if (n < bn) { bn = n; best = j; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) v4[i] = best == 0 ? V[i][0] : best == 1 ? V[i][1] : best == 2 ? V[i][2] : V[i][3];
}

// KannalaBrandt8::TriangulateMatches (:319-388).  Returns z1 (> 0) or the reference's codes -1 .. -5; p3 = x3D on success.
__device__ inline float kb8_triangulate_matches(const orbx_kb8_stereo &S, float u1, float v1, float u2, float v2, float sigma1,
                                                float sigma2, float *p3) {
    float r1x, r1y, r2x, r2y;
    kb8_unproject(S.cam1, S.precision1, u1, v1, r1x, r1y);
    kb8_unproject(S.cam2, S.precision2, u2, v2, r2x, r2y);
    const float *R = S.R12;
    // parallax (:332-338)
    const float qx = dot3(R[0], R[1], R[2], r2x, r2y, 1.0f), qy = dot3(R[3], R[4], R[5], r2x, r2y, 1.0f), qz = dot3(R[6], R[7], R[8], r2x, r2y, 1.0f);
    const float n1 = __fsqrt_rn(dot3(r1x, r1y, 1.0f, r1x, r1y, 1.0f)), n2 = __fsqrt_rn(dot3(qx, qy, qz, qx, qy, qz));
    const float cosp = __fdiv_rn(dot3(r1x, r1y, 1.0f, qx, qy, qz), __fmul_rn(n1, n2));
    if ((double)cosp > 0.9998) return -1.0f;
    // Tcw2 = [R21 | -R21 t12], R21 = R12^T (:352-355)
    float T2[12];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T2[4 * i + 0] = R[i]; T2[4 * i + 1] = R[3 + i]; T2[4 * i + 2] = R[6 + i];
        T2[4 * i + 3] = dot3(-R[i], -R[3 + i], -R[6 + i], S.t12[0], S.t12[1], S.t12[2]);
    }
    // DLT rows (:416-420) with Tcw1 = [I | 0]
    float A[16] = {-1.0f, 0.0f, r1x, 0.0f, 0.0f, -1.0f, r1y, 0.0f};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        A[8 + j] = __fsub_rn(__fmul_rn(r2x, T2[8 + j]), T2[j]);
        A[12 + j] = __fsub_rn(__fmul_rn(r2y, T2[8 + j]), T2[4 + j]);
    }
    double v4[4];
    null_vector4(A, v4);
    const float hx = (float)v4[0], hy = (float)v4[1], hz = (float)v4[2], hw = (float)v4[3];
    const float X = __fdiv_rn(hx, hw), Y = __fdiv_rn(hy, hw), Z = __fdiv_rn(hz, hw);
    if (!(Z > 0.0f)) return -2.0f;
    const float z2 = __fadd_rn(dot3(T2[8], T2[9], T2[10], X, Y, Z), T2[11]);
    if (!(z2 > 0.0f)) return -3.0f;
    float pu, pv;
    kb8_project(S.cam1, X, Y, Z, pu, pv);
    const float e1x = __fsub_rn(pu, u1), e1y = __fsub_rn(pv, v1);
    if ((double)__fadd_rn(__fmul_rn(e1x, e1x), __fmul_rn(e1y, e1y)) > 5.991 * (double)sigma1) return -4.0f;
    const float X2 = __fadd_rn(dot3(T2[0], T2[1], T2[2], X, Y, Z), T2[3]), Y2 = __fadd_rn(dot3(T2[4], T2[5], T2[6], X, Y, Z), T2[7]);
    kb8_project(S.cam2, X2, Y2, z2, pu, pv);
    const float e2x = __fsub_rn(pu, u2), e2y = __fsub_rn(pv, v2);
    if ((double)__fadd_rn(__fmul_rn(e2x, e2x), __fmul_rn(e2y, e2y)) > 5.991 * (double)sigma2) return -5.0f;
    p3[0] = X; p3[1] = Y; p3[2] = Z;
    return Z;
}

// n independent pairs (the bare camera-model call)
__global__ void __launch_bounds__(128) k_kb8_triangulate(const __grid_constant__ orbx_kb8_stereo S, const float *__restrict__ xy1,
                                                         const float *__restrict__ xy2, const float *__restrict__ sigma1,
                                                         const float *__restrict__ sigma2, int n, float *__restrict__ depth,
                                                         float *__restrict__ p3d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[3] = {0.0f, 0.0f, 0.0f};
    depth[i] = kb8_triangulate_matches(S, xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1], sigma1[i], sigma2[i], p);
    p3d[3 * i] = p[0]; p3d[3 * i + 1] = p[1]; p3d[3 * i + 2] = p[2];
}

struct FisheyeArgs {
    const orbx_kp *kpL, *kpR;        // all keypoints of the two cameras (mvKeys, mvKeysRight)
    int nL, nR, monoL, monoR;        // the lapping blocks are [monoL, nL) and [monoR, nR)
    const int *idx, *dist;           // [nL - monoL][2] brute-force 2-NN of the left lapping block in the right one
    float sigma2[16];                // mvLevelSigma2
    double ratio;                    // 0.7 (stereo) / 0.8 (multi)
    int *l2r, *r2l;                  // mvLeftToRightMatch [nL], mvRightToLeftMatch [nR]   (pre-filled with -1)
    float *depth, *p3d;              // mvDepth [nL] (pre-filled with -1), mvStereo3Dpoints [nL][3] (pre-filled with 0)
    int *counters;                   // [2]: nMatches, descMatches
};

// one thread per left lapping keypoint (src/Frame.cc:1432-1457).  The reference visits the queries in ascending order, so for a
// right keypoint matched by several left ones the LAST (largest) query index stays in mvRightToLeftMatch: atomicMax.
__global__ void __launch_bounds__(128) k_fisheye_matches(const __grid_constant__ orbx_kb8_stereo S, const __grid_constant__ FisheyeArgs A) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.nL - A.monoL) return;
    const int i0 = A.idx[2 * q], i1 = A.idx[2 * q + 1];
    if (i0 < 0 || i1 < 0) return;                                                   // (*it).size() >= 2
    if (!((double)(float)A.dist[2 * q] < (double)(float)A.dist[2 * q + 1] * A.ratio)) return;
    atomicAdd(&A.counters[1], 1);
    const int iL = q + A.monoL, iR = i0 + A.monoR;
    const orbx_kp kl = A.kpL[iL], kr = A.kpR[iR];
    float p[3];
    const float d = kb8_triangulate_matches(S, kl.x, kl.y, kr.x, kr.y, A.sigma2[kl.octave & 15], A.sigma2[kr.octave & 15], p);
    if (!(d > 0.0001f)) return;
    A.l2r[iL] = iR;
    atomicMax(&A.r2l[iR], iL);
    A.p3d[3 * iL] = p[0]; A.p3d[3 * iL + 1] = p[1]; A.p3d[3 * iL + 2] = p[2];
    A.depth[iL] = d;
    atomicAdd(&A.counters[0], 1);
}

}  // namespace orbx
