// qgmap_device.cuh -- device-side building blocks of the QGMAP iteration (sm_100a, fp32 arithmetic).
//
// Reference arithmetic being restated (all fp64 there): gqmap_gpu_mixture.m:87-146 (node/edge_grad_spectral),
// :156-182 (node_pot / edge_pot).  Reformulated B200-first:
//   * the five score-function accumulators of :99-103 are linear in six quadrature MOMENTS of the potential
//     (sum f, f*XI, f*XJ, f*XI^2, f*XJ^2, f*XI*XJ); the K x K tensor grid lets the inner (XI) loop keep three
//     running sums with compile-time table operands, folded into the six moments once per XJ row;
//   * sample coordinates are split into floor + fraction BEFORE the pixel index is added, so bicubic weights keep
//     full fp32 precision at any image size (SURVEY hard part 1);
//   * 1-p^2 is formed as (1-p)(1+p) and sqrt(1-p^2) as sqrt(1-p)*sqrt(1+p).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QG_LMAX 10
#define QG_KMAX 32
#define QG_TW 32            // lanes per tile row (lane 0 = halo column)
#define QG_TH 8             // output rows per tile (+1 halo row of threads)
#define QG_NRED 4           // block-reduced scalars: energy, dalpha, sum|G_muu|, sum|G_sigu|

struct __align__(32) QgTap8 { float4 r0, r1; };   // two consecutive image rows x four consecutive columns

// sm_100 256-bit read-only global load (SASS LDG.E.ENL2.256.CONSTANT)
__device__ __forceinline__ QgTap8 qg_ld256(const QgTap8 *p) {
    QgTap8 v;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v.r0.x), "=f"(v.r0.y), "=f"(v.r0.z), "=f"(v.r0.w), "=f"(v.r1.x), "=f"(v.r1.y), "=f"(v.r1.z), "=f"(v.r1.w)
        : "l"(p));
    return v;
}

struct QgTables {           // GaussHermite_2 nodes/weights, premultiplied (gqmap_gpu_mixture.m:8-10)
    float X[QG_KMAX];       // X(c)
    float W[QG_KMAX];       // W(c)
    float WX[QG_KMAX];      // W(c)*X(c)
    float WXX[QG_KMAX];     // W(c)*X(c)^2
};

// Device-resident control block: everything the loop of gqmap_gpu_mixture.m:26-76 carries between iterations,
// so that a CUDA graph of identical kernel nodes can run many iterations with no host involvement.
struct QgCtrl {
    int it;                 // 1-based iteration counter (:25,:74)
    int stop;               // the reference's `break` (:75) fired
    int its;                // options.its
    unsigned int ticket;    // last-block-done counter
    double T;               // current temperature (annealed in the super variant, S:72)
    double w[QG_LMAX];      // softmax logits (:18,:83)
    double alpha[QG_LMAX];  // mixture weights
    double dalpha[QG_LMAX]; // last reduced d(alpha) (:36)
    double sums[QG_LMAX * QG_NRED];   // band mode: this rank's per-component partial sums awaiting all-reduce
};

struct QgIterParams {
    const float *I1;  int pitchI;      // Mo x No row-major
    const QgTap8 *VV8; int pitchV;     // padded second frame (getVV), packed for the gather: VV8[y*pitchV + x] holds
                                       // VV(y, x..x+3) and VV(y+1, x..x+3) in one 32-byte sector, so the 16 taps of a bicubic
                                       // sample are TWO 256-bit loads (LDG.E.256), each lane touching exactly one sector
    float *buf[2];                     // ping-pong state: 9 fields x L planes of rows_local x P floats
    long long plane;                   // floats per plane (rows_local * P)
    int P;                             // row pitch of state planes (floats)
    int M, N, L;                       // belief grid (global)
    int Mo, No;                        // image size
    int g0;                            // global row index of local row 0 (band storage)
    int out_r0, out_r1;                // global rows [out_r0,out_r1) this handle updates
    int K;                             // quadrature order (runtime copy)
    int band;                          // 1: defer control update to the finalize kernel (after all-reduce)
    float lambdad, lambdas, epsn;
    float minu, maxu, minv, maxv, sig_min, sig_max, corr_tor;
    double step0, step_tau;
    double alpha_scale, drate, T_floor, tor;
    int alpha_start, alpha_mode, anneal_every;
    QgCtrl *ctrl;
    double *partials;                  // [nblocks][QG_NRED]
    double *hist_energy, *hist_dmu, *hist_dsig;   // its entries each, index it-1
    float *dbg;                        // DUMP: 11 fields x L planes
    QgTables tab;
};

enum { F_MUU = 0, F_MUV, F_SIGU, F_SIGV, F_PN, F_ROU0, F_ROU1, F_ROU2, F_ROU3, F_COUNT };   // rou q = e + 2*c

__device__ __forceinline__ float qg_sqrt(float x) {          // MUFU.SQRT, ~1 ulp
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float qg_rcp(float x) {           // MUFU.RCP, ~1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct QgMoments { float E, MI, MJ, MII, MJJ, MB; };
struct QgGrad { float da, du1, du2, do1, do2, dp, Ei; };

// Spectral parametrisation of one 2-D Gaussian (a,u1,u2,o1,o2,p): gqmap_gpu_mixture.m:90-93.
struct QgSpectral {
    float s, t, pr, q;      // symmetric sqrt of [[1,p],[p,1]]; pr = 1-p^2; q = sqrt(pr)
    float c1, c2;           // zi - p*zj = c1*XI + c2*XJ ;  zj - p*zi = c2*XI + c1*XJ   (c1 = s - p*t, c2 = t - p*s)
    __device__ __forceinline__ void set(float p) {
        const float omp = 1.0f - p, opp = 1.0f + p;
        const float sp = qg_sqrt(opp), sm = qg_sqrt(omp);
        s = 0.5f * (sp + sm);
        t = 0.5f * (sp - sm);
        pr = omp * opp;
        q = sp * sm;
        // s - p*t and t - p*s cancel catastrophically as |p| -> 1 (the clamp is 1-1e-5) and are then multiplied by 1/pr
        // = 5e4: form them from s-t = sqrt(1-p), s+t = sqrt(1+p) instead, which is exact algebra and cancellation-free.
        if (p >= 0.0f) { c1 = fmaf(t, omp, sm);   c2 = fmaf(s, omp, -sm); }
        else           { c1 = fmaf(-t, opp, sp);  c2 = fmaf(-s, opp, sp); }
    }
};

// Epilogue shared by node (:107-115) and edge (:137-145): moments -> (da,du1,du2,do1,do2,dp,Ei).
// kT: node -3T, edge +T.   m holds moments of the POTENTIAL (already scaled by -lambda).
__device__ __forceinline__ QgGrad qg_epilogue(const QgMoments &m, const QgSpectral &sp, float a, float o1, float o2,
                                              float p, float kT)
{
    const float invpi = 0.31830988618379067f, sqrt2 = 1.4142135623730951f, const1 = 2.8378770664093453f; // 1+log(2pi)
    QgGrad g;
    float A = m.MII + m.MJJ, D = m.MII - m.MJJ;
    float S1 = sp.c1 * m.MI + sp.c2 * m.MJ;
    float S2 = sp.c2 * m.MI + sp.c1 * m.MJ;
    float Sp = p * (m.E - A) + 2.0f * m.MB;
    float Dq = D * qg_rcp(sp.q);
    float T1 = (A - m.E) + Dq, T2 = (A - m.E) - Dq;
    float ipr = qg_rcp(sp.pr), io1 = qg_rcp(o1), io2 = qg_rcp(o2);
    g.du1 = a * S1 * (sqrt2 * invpi) * io1 * ipr;
    g.du2 = a * S2 * (sqrt2 * invpi) * io2 * ipr;
    float H = 0.0f;
    if (kT != 0.0f) H = const1 + logf(sp.q * o1 * o2);
    g.da = m.E * invpi + kT * H;
    g.do1 = a * (T1 * invpi + kT) * io1;
    g.do2 = a * (T2 * invpi + kT) * io2;
    g.dp = a * (Sp * invpi - kT * p) * ipr;
    g.Ei = a * g.da;
    return g;
}

// ---- bicubic taps ---------------------------------------------------------------------------------------------
// 2 x Keys(a=-0.5) weights, gqmap_gpu_mixture.m:164,170,172,174 (the /4 of :176 is applied by the caller).
// w0 = -s(s-1)^2, w3 = s^2(s-1); the other two follow from the partition of unity (sum = 2) and the first moment
// (-w0 + w2 + 2 w3 = 2s): 8 FMA-pipe operations per axis.
__device__ __forceinline__ void qg_cubic_w(float s, float &w0, float &w1, float &w2, float &w3) {
    const float tm = s - 1.0f, st = s * tm;
    w0 = -st * tm;
    w3 = st * s;
    w2 = fmaf(-2.0f, w3, fmaf(2.0f, s, w0));
    w1 = (2.0f - st) - w2;
}

// floor(x) as int + exact fraction without the XU pipe (FRND/F2I are quarter-rate): add 1.5*2^23 rounding towards -inf,
// so floor(x) lands in the mantissa; read it back as an integer.  Valid for |x| < 2^22.
__device__ __forceinline__ int qg_floor_split(float x, float &frac) {
    const float magic = 12582912.0f;
    const float t = __fadd_rd(x, magic);
    frac = x - (t - magic);                          // both subtractions exact in fp32
    return __float_as_int(t) - 0x4B400000;
}

// floor/fraction split + reference clamping (:157-162) for one axis.  pix = 0-based pixel index, x = displacement,
// last = size-2 (largest valid 0-based cell origin).  Returns cell origin; frac in [0,1].
__device__ __forceinline__ int qg_cell(int pix, float x, int last, float &frac) {
    int c = pix + qg_floor_split(x, frac);
    if (c < 0) { c = 0; frac = 0.0f; }
    else if (c > last) { c = last; frac = 1.0f; }
    return c;
}

__device__ __forceinline__ float qg_dot4(const float4 v, float a0, float a1, float a2, float a3) {
    return fmaf(v.w, a3, fmaf(v.z, a2, fmaf(v.y, a1, v.x * a0)));
}

// sqrt(eps + (I1 - bicubic(VV))^2) at displacement (x1 horizontal, x2 vertical) from pixel (m,n) (0-based).
// node_pot = -lambdad * this  (gqmap_gpu_mixture.m:156-179).
// The 4x4 tap block of the cell a thread sampled last.  Once the beliefs have converged (sigma << 1 px) all K*K quadrature
// points of a component fall into one or two cells, so the 16 taps are re-used from registers and the two 256-bit loads are
// predicated off: the L1 gather -- the kernel's co-bottleneck while the beliefs are still scattered -- disappears.
struct QgTapCache {
    int ix, iy;
    QgTap8 v01, v23;
    __device__ __forceinline__ QgTapCache() : ix(-0x40000000), iy(-0x40000000) {}
};

__device__ __forceinline__ float qg_node_sample(const QgTap8 *__restrict__ VV8, int pitchV, int m, int n, int lastx,
                                                int lasty, float x1, float x2, float I1v, float epsn, QgTapCache &tc)
{
    float so, to;
    const int ix = qg_cell(n, x1, lastx, so);
    const int iy = qg_cell(m, x2, lasty, to);
    if (ix != tc.ix || iy != tc.iy) {
        const QgTap8 *r0 = VV8 + (long long)iy * pitchV + ix;    // padded coords: taps rows iy..iy+3, cols ix..ix+3
        tc.v01 = qg_ld256(r0);
        tc.v23 = qg_ld256(r0 + 2 * pitchV);
        tc.ix = ix; tc.iy = iy;
    }
    float a0, a1, a2, a3, b0, b1, b2, b3;
    qg_cubic_w(so, a0, a1, a2, a3);
    qg_cubic_w(to, b0, b1, b2, b3);
    const float h0 = qg_dot4(tc.v01.r0, a0, a1, a2, a3), h1 = qg_dot4(tc.v01.r1, a0, a1, a2, a3);
    const float h2 = qg_dot4(tc.v23.r0, a0, a1, a2, a3), h3 = qg_dot4(tc.v23.r1, a0, a1, a2, a3);
    const float v = fmaf(h3, b3, fmaf(h2, b2, fmaf(h1, b1, h0 * b0)));
    const float d = fmaf(-0.25f, v, I1v);
    return qg_sqrt(fmaf(d, d, epsn));
}

// Super-pixel node sample: sum over the block's 4x4 pixels (gqmap_gpuSuper_mix_entropy.m:99-104), block origin
// pixel (m4,n4) 0-based.  When no pixel of the block is clamped all 16 samples share one set of bicubic weights and
// a 7x7 footprint (49 loads, separable 112+64 FMAs instead of 16 x (16 loads + 20 FMAs)).
__device__ __forceinline__ float qg_super_sample(const QgTap8 *__restrict__ VV8, int pitchV, int m4, int n4, int lastx,
                                                 int lasty, float x1, float x2, const float (&I1b)[16], float epsn)
{
    float so, to;
    const int ix = n4 + qg_floor_split(x1, so), iy = m4 + qg_floor_split(x2, to);
    float acc = 0.0f;
    if (ix >= 0 && ix + 3 <= lastx && iy >= 0 && iy + 3 <= lasty) {
        float a0, a1, a2, a3, b[4];
        qg_cubic_w(so, a0, a1, a2, a3);
        qg_cubic_w(to, b[0], b[1], b[2], b[3]);
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.0f;
        const QgTap8 *rp = VV8 + (long long)iy * pitchV + ix;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {                                  // window rows 2rr, 2rr+1 (row 7 is loaded, unused)
            const QgTap8 lo = qg_ld256(rp), hi = qg_ld256(rp + 3);        // cols ix..ix+3 and ix+3..ix+6
            rp += 2 * pitchV;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = 2 * rr + half;
                if (row < 7) {
                    const float4 l4 = half ? lo.r1 : lo.r0, h4 = half ? hi.r1 : hi.r0;
                    const float v[7] = {l4.x, l4.y, l4.z, l4.w, h4.y, h4.z, h4.w};
#pragma unroll
                    for (int dj = 0; dj < 4; ++dj) {
                        float h = fmaf(v[dj + 3], a3, fmaf(v[dj + 2], a2, fmaf(v[dj + 1], a1, v[dj] * a0)));
#pragma unroll
                        for (int di = 0; di < 4; ++di) {
                            int r = row - di;                       // tap index of this window row for output row di
                            if (r >= 0 && r < 4) o[di * 4 + dj] = fmaf(h, b[r], o[di * 4 + dj]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float d = fmaf(-0.25f, o[i], I1b[i]);
            acc += qg_sqrt(fmaf(d, d, epsn));
        }
    } else {
#pragma unroll 1
        for (int di = 0; di < 4; ++di)
#pragma unroll 1
            for (int dj = 0; dj < 4; ++dj)
                { QgTapCache tc; acc += qg_node_sample(VV8, pitchV, m4 + di, n4 + dj, lastx, lasty, x1, x2, I1b[di * 4 + dj], epsn, tc); }
    }
    return acc;
}

// ---- quadrature loops -------------------------------------------------------------------------------------------
// Tensor-grid accumulation: the inner loop runs over XI = X(c) (k = r + K*c in the reference's column-major table,
// :9), the outer over XJ = X(r).  F(x1,x2) returns the un-scaled potential magnitude; scale = -lambda.
// r0/rstep: the XJ rows this lane evaluates (a group of lanes may share one belief, see qgmap_iter_kernel_g4).
template <int KT, class F>
__device__ __forceinline__ QgMoments qg_quadrature(const QgTables &tab, int Krt, float u1, float u2, float o1, float o2,
                                                   const QgSpectral &sp, float scale, F pot, int r0 = 0, int rstep = 1)
{
    const float sqrt2 = 1.4142135623730951f;
    const int K = KT > 0 ? KT : Krt;
    const float a1s = sqrt2 * o1 * sp.s, a1t = sqrt2 * o1 * sp.t;     // x1 = u1 + a1s*XI + a1t*XJ
    const float a2s = sqrt2 * o2 * sp.s, a2t = sqrt2 * o2 * sp.t;     // x2 = u2 + a2t*XI + a2s*XJ
    QgMoments m = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int r = r0; r < K; r += rstep) {
        const float xj = tab.X[r];
        const float b1 = fmaf(a1t, xj, u1), b2 = fmaf(a2s, xj, u2);
        float S0 = 0.f, S1 = 0.f, S2 = 0.f;
        if (KT > 0) {
#pragma unroll
            for (int c = 0; c < (KT > 0 ? KT : 1); ++c) {
                float f = pot(fmaf(a1s, tab.X[c], b1), fmaf(a2t, tab.X[c], b2));
                S0 = fmaf(tab.W[c], f, S0);
                S1 = fmaf(tab.WX[c], f, S1);
                S2 = fmaf(tab.WXX[c], f, S2);
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < K; ++c) {
                float xc = tab.X[c];
                float f = pot(fmaf(a1s, xc, b1), fmaf(a2t, xc, b2));
                S0 = fmaf(tab.W[c], f, S0);
                S1 = fmaf(tab.WX[c], f, S1);
                S2 = fmaf(tab.WXX[c], f, S2);
            }
        }
        const float wr = tab.W[r], wxr = tab.WX[r], wxxr = tab.WXX[r];
        m.E = fmaf(wr, S0, m.E);
        m.MI = fmaf(wr, S1, m.MI);
        m.MII = fmaf(wr, S2, m.MII);
        m.MJ = fmaf(wxr, S0, m.MJ);
        m.MB = fmaf(wxr, S1, m.MB);
        m.MJJ = fmaf(wxxr, S0, m.MJJ);
    }
    m.E *= scale; m.MI *= scale; m.MJ *= scale; m.MII *= scale; m.MJJ *= scale; m.MB *= scale;
    return m;
}

// Edge quadrature (:118-146): the potential depends on x1-x2 only, which is affine in (XI,XJ):
// x1-x2 = (u1-u2) + A*XI + B*XJ.  One FMA + one MUFU per point.
template <int KT>
__device__ __forceinline__ QgGrad qg_edge(const QgTables &tab, int Krt, float a, float u1, float u2, float o1, float o2,
                                          float p, float lambdas, float epsn, float T)
{
    const float sqrt2 = 1.4142135623730951f;
    const int K = KT > 0 ? KT : Krt;
    QgSpectral sp;
    sp.set(p);
    const float A = sqrt2 * (o1 * sp.s - o2 * sp.t), B = sqrt2 * (o1 * sp.t - o2 * sp.s), d0 = u1 - u2;
    QgMoments m = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int r = 0; r < K; ++r) {
        const float dr = fmaf(B, tab.X[r], d0);
        float S0 = 0.f, S1 = 0.f, S2 = 0.f;
        if (KT > 0) {
#pragma unroll
            for (int c = 0; c < (KT > 0 ? KT : 1); ++c) {
                float d = fmaf(A, tab.X[c], dr);
                float f = qg_sqrt(fmaf(d, d, epsn));
                S0 = fmaf(tab.W[c], f, S0);
                S1 = fmaf(tab.WX[c], f, S1);
                S2 = fmaf(tab.WXX[c], f, S2);
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < K; ++c) {
                float d = fmaf(A, tab.X[c], dr);
                float f = qg_sqrt(fmaf(d, d, epsn));
                S0 = fmaf(tab.W[c], f, S0);
                S1 = fmaf(tab.WX[c], f, S1);
                S2 = fmaf(tab.WXX[c], f, S2);
            }
        }
        const float wr = tab.W[r], wxr = tab.WX[r], wxxr = tab.WXX[r];
        m.E = fmaf(wr, S0, m.E);
        m.MI = fmaf(wr, S1, m.MI);
        m.MII = fmaf(wr, S2, m.MII);
        m.MJ = fmaf(wxr, S0, m.MJ);
        m.MB = fmaf(wxr, S1, m.MB);
        m.MJJ = fmaf(wxxr, S0, m.MJJ);
    }
    const float sc = -lambdas;
    m.E *= sc; m.MI *= sc; m.MJ *= sc; m.MII *= sc; m.MJJ *= sc; m.MB *= sc;
    return qg_epilogue(m, sp, a, o1, o2, p, T);
}
