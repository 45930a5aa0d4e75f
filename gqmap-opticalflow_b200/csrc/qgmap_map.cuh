// qgmap_map.cuh -- MAP extraction (get_map_mex replacement) and the monitoring kernels (profile_logP, AEPE).
// All fp64: these run every 300 iterations (gqmap_gpu_mixture.m:52-68) and must agree with the reference's
// fp64 results to rounding; "MAP argmax indices bit-exact given identical beliefs" is a parity gate.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "qgmap_device.cuh"

// neg_mixture, legacy/findMixMax.m:31-38 (operation order: square, negate, divide by 2*o^2, exp, *a, / (sqrt2pi*o)).
__device__ __forceinline__ double qg_neg_mixture(double x, const double *a, const double *u, const double *o, int L) {
    const double sqrt2pi = 2.5066282746310002;
    double v = 0.0;
    for (int l = 0; l < L; ++l) {
        double d = x - u[l];
        v = v + a[l] * exp(-(d * d) / (2.0 * (o[l] * o[l]))) / (sqrt2pi * o[l]);
    }
    return -v;
}

__device__ __forceinline__ double qg_sign(double v) { return (double)((v > 0.0) - (v < 0.0)); }

// findmin of legacy/findMixMax.m:15-30 with MATLAB R2018b fminbnd (TolX=1e-4, MaxFunEvals=MaxIter=500) as compiled
// into get_map_mex.mexw64.  Returns the per-layer mixture mode.
__device__ inline double qg_findmin(const double *a, const double *u, const double *o, int L)
{
    const double sqrt2pi = 2.5066282746310002;
    double spk = INFINITY;
    int sid = 0;
    double ax = u[0], bx = u[0];
    for (int l1 = 0; l1 < L; ++l1) {
        double vl = 0.0;
        for (int l2 = 0; l2 < L; ++l2) {
            double d = u[l1] - u[l2];
            vl = vl - a[l2] * exp(-(d * d) / (2.0 * (o[l2] * o[l2]))) / (sqrt2pi * o[l2]);
        }
        if (vl < spk) { spk = vl; sid = l1; }          // strict <: first minimum wins
        ax = fmin(ax, u[l1]);
        bx = fmax(bx, u[l1]);
    }
    // ---- fminbnd (Brent: golden section + parabolic interpolation) ----
    const double tol = 1e-4, seps = 1.4901161193847656e-08, c = 0.3819660112501051;
    double A = ax, B = bx;
    double v = A + c * (B - A), w = v, xf = v, d = 0.0, e = 0.0, x = xf;
    double fx = qg_neg_mixture(x, a, u, o, L);
    int funccount = 1, iter = 0;
    double fv = fx, fw = fx, xm = 0.5 * (A + B), tol1 = seps * fabs(xf) + tol / 3.0, tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (B - A))) {
        bool gs = true;
        if (fabs(e) > tol1) {
            gs = false;
            double r = (xf - w) * (fx - fv);
            double q = (xf - v) * (fx - fw);
            double pp = (xf - v) * q - (xf - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp;
            q = fabs(q);
            r = e; e = d;
            if ((fabs(pp) < fabs(0.5 * q * r)) && (pp > q * (A - xf)) && (pp < q * (B - xf))) {
                d = pp / q; x = xf + d;
                if (((x - A) < tol2) || ((B - x) < tol2)) {
                    double si = qg_sign(xm - xf) + ((xm - xf) == 0.0 ? 1.0 : 0.0);
                    d = tol1 * si;
                }
            } else gs = true;
        }
        if (gs) {
            e = (xf >= xm) ? (A - xf) : (B - xf);
            d = c * e;
        }
        double si = qg_sign(d) + (d == 0.0 ? 1.0 : 0.0);
        x = xf + si * fmax(fabs(d), tol1);
        double fu = qg_neg_mixture(x, a, u, o, L);
        funccount++; iter++;
        if (fu <= fx) {
            if (x >= xf) A = xf; else B = xf;
            v = w; fv = fw; w = xf; fw = fx; xf = x; fx = fu;
        } else {
            if (x < xf) A = x; else B = x;
            if ((fu <= fw) || (w == xf)) { v = w; fv = fw; w = x; fw = fu; }
            else if ((fu <= fv) || (v == xf) || (v == w)) { v = x; fv = fu; }
        }
        xm = 0.5 * (A + B); tol1 = seps * fabs(xf) + tol / 3.0; tol2 = 2.0 * tol1;
        if (funccount >= 500 || iter >= 500) break;
    }
    return (fx < spk) ? xf : u[sid];
}

// One thread per (pixel, layer).  Source beliefs either fp64 column-major host layout (stateless get_map_mex call) or
// the handle's fp32 row-major planes.  Output: fp64 column-major M x N x 2.
template <typename TIn>
__global__ void qgmap_find_map_kernel(const double *__restrict__ alpha, const TIn *__restrict__ mu_u,
                                      const TIn *__restrict__ sig_u, const TIn *__restrict__ mu_v,
                                      const TIn *__restrict__ sig_v, long long comp_stride, int M, int N, int L,
                                      int in_row_major, int in_pitch, int row_off, int r0, int r1, double *__restrict__ map)
{
    const long long MN = (long long)M * N;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * MN) return;
    const int layer = (int)(t / MN);
    const long long pix = t - (long long)layer * MN;           // column-major pixel index: m + M*n
    const int m = (int)(pix % M), n = (int)(pix / M);
    if (m < r0 || m >= r1) return;                             // a row band extracts the rows it owns
    const long long off = in_row_major ? ((long long)(m - row_off) * in_pitch + n) : pix;
    const TIn *mu = layer ? mu_v : mu_u, *sg = layer ? sig_v : sig_u;
    double a[QG_LMAX], u[QG_LMAX], o[QG_LMAX];
    for (int l = 0; l < L; ++l) {
        a[l] = alpha[l];
        u[l] = (double)mu[off + comp_stride * l];
        o[l] = (double)sg[off + comp_stride * l];
    }
    map[t] = (L == 1) ? u[0] : qg_findmin(a, u, o, L);          // gqmap_gpu_mixture.m:54-58
}

// ---- fp64 node_pot for monitoring (gqmap_gpu_mixture.m:156-179), images fp64 row-major -----------------------------
__device__ inline double qg_node_pot_f64(const double *__restrict__ I1, int pitchI, const double *__restrict__ VV,
                                         int pitchV, int Mo, int No, double lambdad, double epsn, double x1, double x2,
                                         int i, int j /* 1-based */)
{
    double Xq = fmin(fmax((double)j + x1, 1.0), (double)No);
    double Yq = fmin(fmax((double)i + x2, 1.0), (double)Mo);
    double ix = (Xq <= 1.0) ? 1.0 : ((Xq <= (double)(No - 1)) ? floor(Xq) : (double)(No - 1));
    double iy = (Yq <= 1.0) ? 1.0 : ((Yq <= (double)(Mo - 1)) ? floor(Yq) : (double)(Mo - 1));
    double so = Xq - ix, to = Yq - iy;
    double wt[4] = {((2.0 - to) * to - 1.0) * to, (3.0 * to - 5.0) * to * to + 2.0, ((4.0 - 3.0 * to) * to + 1.0) * to,
                    (to - 1.0) * to * to};
    double ws[4] = {((2.0 - so) * so - 1.0) * so, (3.0 * so - 5.0) * so * so + 2.0, ((4.0 - 3.0 * so) * so + 1.0) * so,
                    (so - 1.0) * so * so};
    const double *p0 = VV + (long long)((int)iy - 1) * pitchV + ((int)ix - 1);   // padded 0-based origin
    double Vq = 0.0;
    for (int c = 0; c < 4; ++c)                       // same association order as :169-175
        for (int r = 0; r < 4; ++r) Vq = Vq + p0[(long long)r * pitchV + c] * ws[c] * wt[r];
    Vq = Vq / 4.0;
    double d = I1[(long long)(i - 1) * pitchI + (j - 1)] - Vq;
    return -lambdad * sqrt(epsn + d * d);
}

struct QgMonParams {
    const double *I1; int pitchI;
    const double *VV; int pitchV;
    int Mo, No, M, N, super;
    double lambdad, lambdas, epsn;
    int r0, r1;                    // belief rows [r0,r1) this call sums over (a row band's share); whole grid: 0, M
};

__device__ __forceinline__ double qg_block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    __syncthreads();
    return s;
}

// profile_logP (gqmap_gpu_mixture.m:148-154, super S:152-169): per-block partial sums -> partials[blockIdx.x].
// uv: fp64 column-major M x N x 2.
__global__ void qgmap_logp_kernel(const QgMonParams q, const double *__restrict__ uv, double *__restrict__ partials)
{
    __shared__ double sh[32];
    const long long MN = (long long)q.M * q.N;
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < MN; t += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(t % q.M), n = (int)(t / q.M);
        if (m < 1 || m > q.M - 2 || n < 1 || n > q.N - 2 || m < q.r0 || m >= q.r1) continue;
        const double us = uv[t], vs = uv[t + MN];
        double lp = 0.0;
        if (q.super) {
            for (int i = 4 * (m + 1) - 3; i <= 4 * (m + 1); ++i)
                for (int j = 4 * (n + 1) - 3; j <= 4 * (n + 1); ++j)
                    lp += qg_node_pot_f64(q.I1, q.pitchI, q.VV, q.pitchV, q.Mo, q.No, q.lambdad, q.epsn, us, vs, i, j);
        } else {
            lp += qg_node_pot_f64(q.I1, q.pitchI, q.VV, q.pitchV, q.Mo, q.No, q.lambdad, q.epsn, us, vs, m + 1, n + 1);
        }
        for (int c = 0; c < 2; ++c) {
            const double x = uv[t + MN * c];
            double d = x - uv[(m + 1) % q.M + (long long)q.M * n + MN * c];            // circshift(uv,-1)
            lp += -q.lambdas * sqrt(q.epsn + d * d);
            d = x - uv[m + (long long)q.M * ((n + 1) % q.N) + MN * c];                 // circshift(uv,-1,2)
            lp += -q.lambdas * sqrt(q.epsn + d * d);
        }
        acc += lp;
    }
    double s = qg_block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// AEPE (gqmap_gpu_mixture.m:63-64; super S:58-63: repelem(map,4,4), crop 5:end-4).  map fp64 col-major M x N x 2,
// tflow fp64 col-major Mo x No x 2, unknown uint8 col-major Mo x No (may be NULL).
__global__ void qgmap_aepe_kernel(const QgMonParams q, const double *__restrict__ map, const double *__restrict__ tflow,
                                  const unsigned char *__restrict__ unknown, double *__restrict__ partials)
{
    __shared__ double sh[32];
    const long long MoNo = (long long)q.Mo * q.No, MN = (long long)q.M * q.N;
    const int b = q.super ? 4 : 1, sc = q.super ? 4 : 1;
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < MoNo; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % q.Mo), j = (int)(t / q.Mo);
        if (i < b || i >= q.Mo - b || j < b || j >= q.No - b || i / sc < q.r0 || i / sc >= q.r1) continue;
        const long long mi = i / sc + (long long)q.M * (j / sc);
        double fu = map[mi], fv = map[mi + MN];
        if (unknown && unknown[t]) { fu = 0.0; fv = 0.0; }
        const double du = tflow[t] - fu, dv = tflow[t + MoNo] - fv;
        acc += sqrt(du * du + dv * dv);
    }
    double s = qg_block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

