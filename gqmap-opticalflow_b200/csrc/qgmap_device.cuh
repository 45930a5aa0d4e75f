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
#include <cuda_fp16.h>
#include <stdint.h>

#define QG_LMAX 10
#define QG_KMAX 32
#define QG_TW 32            // lanes per tile row (lane 0 = halo column)
// Tile shape per kernel instantiation: a CTA is QG_TW lanes x (TH + W0) warps, compiled for MINB resident CTAs per SM.  W0 = 1: warp 0
// is a halo warp (the row above the tile; it evaluates only the down edge whose endpoint-2 gradient the first output row needs and
// then idles at the block reduction).  W0 = 0: no halo warp -- the first output row evaluates that edge itself, after the exchange
// barrier, through the same call site as its own down edge (qgmap_iter.cuh).  Measured on B200 (profiles/r01_tile_ab.txt,
// profiles/r02_paths_ab.txt): K <= 5 is latency-bound and gains from 4 CTAs of 8 warps at 64 registers; without the halo warp all eight
// own rows (K=5: +5.9% at 4K, +3.4% converged); K=3 keeps the halo warp (-2.5% without: its edges are half of a thread's work);
// K >= 7: 3 CTAs of 8 warps (+1.5%); run-time K keeps 3 CTAs of 9 warps.  The super-pixel variant has few, heavy beliefs: smaller CTAs, so a straggler warp (border blocks
// take a slower clamped path) holds back fewer warps at the block reduction.
#ifndef QG_HALO_WARP
#define QG_HALO_WARP 0      // 1: halo warp in every instantiation (A/B builds)
#endif
#ifndef QG_SMALLK_MINB
#define QG_SMALLK_MINB 4
#endif
#ifndef QG_BIGK_TH
#define QG_BIGK_TH 8        // rows (= warps) of the K >= 7 tile without a halo warp, 3 resident CTAs
#endif
#ifndef QG_K5_TH
#define QG_K5_TH 8          // rows (= warps) of the K=5 tile without a halo warp; QG_K5_MINB resident CTAs
#define QG_K5_MINB QG_SMALLK_MINB
#endif
#ifndef QG_SUPER_TH
#define QG_SUPER_TH 4
#define QG_SUPER_MINB 4
#endif
#ifndef QG_SUPER_MINB_NOHALO
#define QG_SUPER_MINB_NOHALO 5      // four-lane kernel without a halo warp: 5 CTAs x 4 warps at 96 registers
#endif
__host__ __device__ constexpr int qg_template_k(int K) { return (K == 3 || K == 5 || K == 7 || K == 9 || K == 11) ? K : 0; }   // 0: run-time K
// K below: the template K (qg_template_k).  Halo warp kept for K = 3 (-2.5% without), run-time K (-4.7%) and the one-lane super-pixel form.
__host__ __device__ constexpr int qg_tile_halo(int K, bool super) { return (QG_HALO_WARP || super || K == 3 || K == 0) ? 1 : 0; }
__host__ __device__ constexpr int qg_tile_rows(int K, bool super) {      // rows per tile
    return super ? QG_SUPER_TH : ((K == 5 && !qg_tile_halo(K, super)) ? QG_K5_TH :
                                  ((K >= 7 && !qg_tile_halo(K, super)) ? QG_BIGK_TH : 8 - ((K == 3 || K == 5) ? qg_tile_halo(K, super) : 0)));
}
template <int KT, bool SUPER> struct QgTile {
    static constexpr int W0 = qg_tile_halo(KT, SUPER);
    static constexpr int TH = qg_tile_rows(KT, SUPER);
    static constexpr int MINB = SUPER ? QG_SUPER_MINB : ((KT == 5 && !W0) ? QG_K5_MINB : ((KT > 0 && KT <= 5) ? QG_SMALLK_MINB : 3));
};
#define QG_TH_MAX 8
#define QG_NRED 4           // block-reduced scalars: energy, dalpha, sum|G_muu|, sum|G_sigu|

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 -- two FMAs per issue slot; the iteration kernel is issue-bound,
// not FMA-pipe-bound, so pairing work halves the slots that work costs).  Negations fold into SASS operand modifiers and a
// scalar operand is broadcast to both halves (R.F32 / UR.F32 / immediate), so no extra moves are needed. -----------------
typedef unsigned long long qg_u64;
__device__ __forceinline__ qg_u64 qg_bits(float2 v) { return *reinterpret_cast<qg_u64 *>(&v); }
__device__ __forceinline__ float2 qg_f2(qg_u64 v) { return *reinterpret_cast<float2 *>(&v); }
__device__ __forceinline__ float2 qg_fma2(float2 a, float2 b, float2 c) {
    qg_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(qg_bits(a)), "l"(qg_bits(b)), "l"(qg_bits(c)));
    return qg_f2(d);
}
__device__ __forceinline__ float2 qg_mul2(float2 a, float2 b) {
    qg_u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(qg_bits(a)), "l"(qg_bits(b)));
    return qg_f2(d);
}
__device__ __forceinline__ float2 qg_add2(float2 a, float2 b) {
    qg_u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(qg_bits(a)), "l"(qg_bits(b)));
    return qg_f2(d);
}
__device__ __forceinline__ float2 qg_sub2(float2 a, float2 b) {
    qg_u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(qg_bits(a)), "l"(qg_bits(b)));
    return qg_f2(d);
}
__device__ __forceinline__ float2 qg_add2_rm(float2 a, float2 b) {       // round towards -inf
    qg_u64 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(qg_bits(a)), "l"(qg_bits(b)));
    return qg_f2(d);
}
__device__ __forceinline__ float2 qg_bc(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 qg_neg2(float2 v) { return make_float2(-v.x, -v.y); }

// Gather entry: two consecutive image rows x four consecutive columns, interleaved BY ROW inside each column
// (p[c] = (VV(y, x+c), VV(y+1, x+c))), so that one FFMA2 with a broadcast column weight advances both rows.
struct __align__(32) QgTap8 { float2 p[4]; };

// State loads: each value is read once per iteration, so it must not displace the gather entries in L1 (+1.3% at K=5).
__device__ __forceinline__ float qg_lds(const float *ptr) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(ptr));
    return v;
}
// sm_100 256-bit read-only global load (SASS LDG.E.ENL2.256.CONSTANT)
__device__ __forceinline__ QgTap8 qg_ld256(const QgTap8 *ptr) {
    QgTap8 v;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v.p[0].x), "=f"(v.p[0].y), "=f"(v.p[1].x), "=f"(v.p[1].y), "=f"(v.p[2].x), "=f"(v.p[2].y), "=f"(v.p[3].x), "=f"(v.p[3].y)
        : "l"(ptr));
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
    int comm_error;         // band p2p mode: a peer did not publish its iteration in time (the run was stopped)
};

struct QgIterParams {
    const float *I1;  int pitchI;      // Mo x No row-major
    const struct QgTap16h *VVh;        // the same frame as 4 x 4 blocks of fp16 (one sector per sample) -- only when every value of
                                       // the padded frame is exactly representable in fp16 (integer grey levels); else null
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
    int band;                          // 1: leave the band's partial sums in ctrl->sums (a later kernel all-reduces them and advances
                                       // the control block); 2: exchange rows and sums with the peers inside the iteration kernel
    float lambdad, lambdas, epsn;
    float minu, maxu, minv, maxv, sig_min, sig_max, corr_tor, sig_step;
    double step0, step_tau;
    double alpha_scale, drate, T_floor, tor;
    int alpha_start, alpha_mode, anneal_every;
    QgCtrl *ctrl;
    double *partials;                  // [nblocks][QG_NRED]
    double *gpartials;                 // row-walking kernel: [strip rows x L][QG_NRED], second reduction level
    unsigned int *tickets;             // row-walking kernel: [strip rows x L] finished-strip counters
    int strip_rows;                    // row-walking kernel: rows one warp walks
    float wide_reach;                  // half-extent (px) of the sample cloud from which a belief counts as wide (QGMAP_WIDE_REACH)
    int pub_row[2];                    // band == 2: global row whose updated beliefs also go to the band above [0] / below [1]; -1 = none
    const struct QgPeer *peer;         // band == 2: peer-memory exchange fused into the row-walking kernel (qgmap_peer.cuh)
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
// 2 x Keys(a=-0.5) weights, gqmap_gpu_mixture.m:164,170,172,174 (the /4 of :176 is applied by the caller), for the pair
// s = (so, to) at once: w0 = -s(1-s)^2, w3 = -s^2(1-s); the other two follow from the partition of unity (sum = 2) and the
// first moment (-w0 + w2 + 2 w3 = 2s).  Returns n0 = -w0 and n3 = -w3 (the signs fold into operand modifiers). 9 packed ops.
__device__ __forceinline__ void qg_cubic_w2(float2 s, float2 &n0, float2 &w1, float2 &w2, float2 &n3) {
    const float2 om = qg_sub2(qg_bc(1.0f), s);
    const float2 q = qg_mul2(s, om);
    n0 = qg_mul2(q, om);
    n3 = qg_mul2(q, s);
    w2 = qg_fma2(n3, qg_bc(2.0f), qg_sub2(qg_add2(s, s), n0));
    w1 = qg_sub2(qg_add2(q, qg_bc(2.0f)), w2);
}
__device__ __forceinline__ void qg_cubic_w(float s, float &w0, float &w1, float &w2, float &w3) {
    const float tm = s - 1.0f, st = s * tm;
    w0 = -st * tm;
    w3 = st * s;
    w2 = fmaf(-2.0f, w3, fmaf(2.0f, s, w0));
    w1 = (2.0f - st) - w2;
}

// floor(x) as int + exact fraction without the XU pipe (FRND/F2I are quarter-rate): add 1.5*2^23 rounding towards -inf,
// so floor(x) lands in the mantissa; read it back as an integer.  Valid for |x| < 2^22.  Both axes at once.
__device__ __forceinline__ float2 qg_floor_split2(float2 x, int &ix, int &iy) {
    const float2 magic = qg_bc(12582912.0f);
    const float2 t = qg_add2_rm(x, magic);
    ix = __float_as_int(t.x) - 0x4B400000;
    iy = __float_as_int(t.y) - 0x4B400000;
    return qg_sub2(x, qg_sub2(t, magic));            // exact in fp32
}
__device__ __forceinline__ int qg_floor_split(float x, float &frac) {
    const float magic = 12582912.0f;
    const float t = __fadd_rd(x, magic);
    frac = x - (t - magic);
    return __float_as_int(t) - 0x4B400000;
}

// The 4x4 tap block of the cell a thread sampled last.  Once the beliefs have converged (sigma << 1 px) all K*K quadrature
// points of a component fall into one or two cells, so the 16 taps are re-used from registers and the two 256-bit loads are
// predicated off: the L1 gather -- the kernel's co-bottleneck while the beliefs are still scattered -- disappears.
struct QgTapCache {
    int ix, iy;
    QgTap8 v01, v23;
    __device__ __forceinline__ QgTapCache() : ix(-0x40000000), iy(-0x40000000) {}
};

// sqrt(eps + (I1 - bicubic(VV))^2) at displacement x = (x1 horizontal, x2 vertical) from pixel (m,n) (0-based).
// node_pot = -lambdad * this  (gqmap_gpu_mixture.m:156-179).  x- and y-axis arithmetic runs as one fp32x2 stream.
// CLAMP_FIRST: the reference clamping (:157-162: Xq = min(max(j + x1, 1), N), cell index capped at N-1) is applied to the
// displacement itself, in coordinates relative to the pixel, where the bounds are small integers and exact in fp32: four FMNMX
// instead of the twelve compare/select instructions of clamping cell index and fraction separately.  A sample clamped at the far
// edge lands on cell last+1 with fraction 0 instead of cell last with fraction 1 -- the same tap with weight exactly 2, the other
// weights exactly 0, and the gather layout is zero-filled beyond the padded image: the value is bit-identical (the Energy of a
// 4000-iteration run does not change in any digit).  Measured (profiles/r01_tile_ab.txt, experiment 9): +3..4% for K >= 7
// (issue-bound, 72 registers); -1..-3% for K <= 5 and -7% for the super-pixel kernel, whose register budgets (64 / 96) turn the
// four extra live bounds into spills -- so it is a per-instantiation choice.
template <bool CLAMP_FIRST>
__device__ __forceinline__ float qg_node_sample(const QgTap8 *__restrict__ VV8, int pitchV, int m, int n, int lastx,
                                                int lasty, float2 x, float I1v, float epsn, QgTapCache &tc)
{
    int ix, iy;
    if (CLAMP_FIRST) {
        x.x = fminf(fmaxf(x.x, (float)(-n)), (float)(lastx + 1 - n));
        x.y = fminf(fmaxf(x.y, (float)(-m)), (float)(lasty + 1 - m));
    }
    float2 fr = qg_floor_split2(x, ix, iy);
    ix += n; iy += m;
    if (!CLAMP_FIRST) {
        if (ix < 0) { ix = 0; fr.x = 0.0f; } else if (ix > lastx) { ix = lastx; fr.x = 1.0f; }     // reference clamping :157-162
        if (iy < 0) { iy = 0; fr.y = 0.0f; } else if (iy > lasty) { iy = lasty; fr.y = 1.0f; }
    }
    if (ix != tc.ix || iy != tc.iy) {
        const QgTap8 *r0 = VV8 + (long long)iy * pitchV + ix;    // padded coords: taps rows iy..iy+3, cols ix..ix+3
        tc.v01 = qg_ld256(r0);
        tc.v23 = qg_ld256(r0 + 2 * pitchV);
        tc.ix = ix; tc.iy = iy;
    }
    float2 n0, w1, w2, n3;                                       // .x: column (so) weights, .y: row (to) weights
    qg_cubic_w2(fr, n0, w1, w2, n3);
    // horizontal pass, two rows per instruction: (h0,h1) and (h2,h3); column weight broadcast
    float2 h01 = qg_mul2(tc.v01.p[0], qg_bc(-n0.x));
    float2 h23 = qg_mul2(tc.v23.p[0], qg_bc(-n0.x));
    h01 = qg_fma2(tc.v01.p[1], qg_bc(w1.x), h01);  h23 = qg_fma2(tc.v23.p[1], qg_bc(w1.x), h23);
    h01 = qg_fma2(tc.v01.p[2], qg_bc(w2.x), h01);  h23 = qg_fma2(tc.v23.p[2], qg_bc(w2.x), h23);
    h01 = qg_fma2(tc.v01.p[3], qg_bc(-n3.x), h01); h23 = qg_fma2(tc.v23.p[3], qg_bc(-n3.x), h23);
    const float v = fmaf(h23.y, -n3.y, fmaf(h23.x, w2.y, fmaf(h01.y, w1.y, h01.x * -n0.y)));
    const float d = fmaf(-0.25f, v, I1v);
    return qg_sqrt(fmaf(d, d, epsn));
}

// Super-pixel node sample: sum over the block's 4x4 pixels (gqmap_gpuSuper_mix_entropy.m:99-104), block origin
// pixel (m4,n4) 0-based.  When no pixel of the block is clamped all 16 samples share one set of bicubic weights and
// a 7x7 footprint (49 loads, separable 112+64 FMAs instead of 16 x (16 loads + 20 FMAs)).
__device__ __forceinline__ float qg_super_sample(const QgTap8 *__restrict__ VV8, int pitchV, int m4, int n4, int lastx,
                                                 int lasty, float2 x, const float (&I1b)[16], float epsn)
{
    int ix, iy;
    const float2 fr = qg_floor_split2(x, ix, iy);
    ix += n4; iy += m4;
    float acc = 0.0f;
    if (ix >= 0 && ix + 3 <= lastx && iy >= 0 && iy + 3 <= lasty) {
        float2 n0, w1, w2, n3;
        qg_cubic_w2(fr, n0, w1, w2, n3);
        const float a[4] = {-n0.x, w1.x, w2.x, -n3.x}, b[4] = {-n0.y, w1.y, w2.y, -n3.y};
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.0f;
        const QgTap8 *rp = VV8 + (long long)iy * pitchV + ix;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {                                  // window rows 2rr, 2rr+1 (row 7 is loaded, unused)
            const QgTap8 lo = qg_ld256(rp), hi = qg_ld256(rp + 3);        // cols ix..ix+3 and ix+3..ix+6
            rp += 2 * pitchV;
            const float2 v[7] = {lo.p[0], lo.p[1], lo.p[2], lo.p[3], hi.p[1], hi.p[2], hi.p[3]};
#pragma unroll
            for (int dj = 0; dj < 4; ++dj) {
                float2 h = qg_mul2(v[dj], qg_bc(a[0]));                   // rows 2rr and 2rr+1 together
                h = qg_fma2(v[dj + 1], qg_bc(a[1]), h);
                h = qg_fma2(v[dj + 2], qg_bc(a[2]), h);
                h = qg_fma2(v[dj + 3], qg_bc(a[3]), h);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = 2 * rr + half;
                    const float hv = half ? h.y : h.x;
#pragma unroll
                    for (int di = 0; di < 4; ++di) {
                        const int r = row - di;                           // tap index of this window row for output row di
                        if (row < 7 && r >= 0 && r < 4) o[di * 4 + dj] = fmaf(hv, b[r], o[di * 4 + dj]);
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
            for (int dj = 0; dj < 4; ++dj) {
                QgTapCache tc;
                acc += qg_node_sample<false>(VV8, pitchV, m4 + di, n4 + dj, lastx, lasty, x, I1b[di * 4 + dj], epsn, tc);
            }
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
    // (x1,x2) = (u1,u2) + aI*XI + aJ*XJ  with aI = sqrt2*(o1*s, o2*t), aJ = sqrt2*(o1*t, o2*s)
    const float2 aI = make_float2(sqrt2 * o1 * sp.s, sqrt2 * o2 * sp.t), aJ = make_float2(sqrt2 * o1 * sp.t, sqrt2 * o2 * sp.s);
    const float2 u12 = make_float2(u1, u2);
    QgMoments m = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int r = r0; r < K; r += rstep) {
        const float2 b12 = qg_fma2(aJ, qg_bc(tab.X[r]), u12);
        float S0 = 0.f, S1 = 0.f, S2 = 0.f;
        if (KT > 0) {
#pragma unroll
            for (int c = 0; c < (KT > 0 ? KT : 1); ++c) {
                float f = pot(qg_fma2(aI, qg_bc(tab.X[c]), b12));
                S0 = fmaf(tab.W[c], f, S0);
                S1 = fmaf(tab.WX[c], f, S1);
                S2 = fmaf(tab.WXX[c], f, S2);
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < K; ++c) {
                float f = pot(qg_fma2(aI, qg_bc(tab.X[c]), b12));
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

// Flat split of the K*K quadrature points over `nlanes` lanes (k = lane, lane+nlanes, ...): better balance than whole rows
// when K is small against the lane count (K=5 over 4 lanes: 7,6,6,6 points instead of 10,5,5,5).  Used by the super-pixel
// variant, whose ~300-instruction node sample dwarfs the per-point table look-ups.
template <int KT, class F>
__device__ __forceinline__ QgMoments qg_quadrature_flat(const QgTables &tab, int Krt, float u1, float u2, float o1, float o2,
                                                        const QgSpectral &sp, float scale, F pot, int lane, int nlanes)
{
    const float sqrt2 = 1.4142135623730951f;
    const int K = KT > 0 ? KT : Krt;
    const float2 aI = make_float2(sqrt2 * o1 * sp.s, sqrt2 * o2 * sp.t), aJ = make_float2(sqrt2 * o1 * sp.t, sqrt2 * o2 * sp.s);
    const float2 u12 = make_float2(u1, u2);
    QgMoments m = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int k = lane; k < K * K; k += nlanes) {
        const int c = k / K, r = k - c * K;                        // reference order k = r + K*c (:9)
        const float xc = tab.X[c], xr = tab.X[r];
        const float f = pot(qg_fma2(aI, qg_bc(xc), qg_fma2(aJ, qg_bc(xr), u12)));
        const float wf = tab.W[c] * tab.W[r] * f;
        m.E += wf;
        m.MI = fmaf(wf, xc, m.MI);
        m.MJ = fmaf(wf, xr, m.MJ);
        m.MII = fmaf(wf * xc, xc, m.MII);
        m.MJJ = fmaf(wf * xr, xr, m.MJJ);
        m.MB = fmaf(wf * xc, xr, m.MB);
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

// ---- packed epilogue: both flow layers of one edge direction at once (.x = u layer, .y = v layer); gqmap_gpu_mixture.m:137-145.
struct QgSpectral2 {
    float2 s, t, pr, q, c1, c2;
    __device__ __forceinline__ void set(float2 p) {
        const float2 one = qg_bc(1.0f);
        const float2 omp = qg_sub2(one, p), opp = qg_add2(one, p);
        const float2 sp = make_float2(qg_sqrt(opp.x), qg_sqrt(opp.y)), sm = make_float2(qg_sqrt(omp.x), qg_sqrt(omp.y));
        s = qg_mul2(qg_add2(sp, sm), qg_bc(0.5f));
        t = qg_mul2(qg_sub2(sp, sm), qg_bc(0.5f));
        pr = qg_mul2(omp, opp);
        q = qg_mul2(sp, sm);
        // c1 = s - p*t, c2 = t - p*s without cancellation (see QgSpectral::set): p >= 0: c1 = t*omp + sm, c2 = s*omp - sm;
        // p < 0: c1 = -t*opp + sp, c2 = -s*opp + sp.  Select the operands per layer, then two packed FMAs.
        const bool px = p.x >= 0.0f, py = p.y >= 0.0f;
        const float2 A = make_float2(px ? omp.x : opp.x, py ? omp.y : opp.y);
        const float2 B = make_float2(px ? sm.x : sp.x, py ? sm.y : sp.y);
        const float2 sg = make_float2(px ? 1.0f : -1.0f, py ? 1.0f : -1.0f);
        c1 = qg_fma2(qg_mul2(sg, t), A, B);
        c2 = qg_fma2(qg_mul2(sg, s), A, qg_neg2(qg_mul2(sg, B)));
    }
};

struct QgGrad2 { float2 da, du1, du2, do1, do2, dp, Ei; };

// Raw moments of the potential and their common scale sc = -lambda/pi.  kT = +T for edges.
// ptxas contracts a packed mul.rn.f32x2 that feeds an add/sub.rn.f32x2 into FFMA2 (also under -fmad=false), and may do so differently in
// two inlined copies of this function; the iteration kernel evaluates down edges at two call sites whose results must agree bit for bit
// (qgmap_iter.cuh), so no packed product feeds a packed sum here: sums are formed before the scaling and every multiply-add is written
// as the FMA it is meant to be.
__device__ __forceinline__ QgGrad2 qg_epilogue2(float2 Er, float2 MIr, float2 MJr, float2 MIIr, float2 MJJr, float2 MBr, float2 sc,
                                                const QgSpectral2 &sp, float a, float2 o1, float2 o2, float2 p, float kT)
{
    const float sqrt2 = 1.4142135623730951f, const1 = 2.8378770664093453f;   // 1+log(2pi)
    QgGrad2 g;
    const float2 Ar = qg_add2(MIIr, MJJr);
    const float2 E = qg_mul2(Er, sc), MI = qg_mul2(MIr, sc), MJ = qg_mul2(MJr, sc);
    const float2 D = qg_mul2(qg_sub2(MIIr, MJJr), sc);
    const float2 S1 = qg_fma2(sp.c1, MI, qg_mul2(sp.c2, MJ));
    const float2 S2 = qg_fma2(sp.c2, MI, qg_mul2(sp.c1, MJ));
    const float2 EmA = qg_mul2(qg_sub2(Er, Ar), sc);
    const float2 Sp = qg_fma2(p, EmA, qg_mul2(qg_mul2(MBr, sc), qg_bc(2.0f)));
    const float2 rq = make_float2(qg_rcp(sp.q.x), qg_rcp(sp.q.y));
    const float2 T1 = qg_fma2(D, rq, qg_neg2(EmA)), T2 = qg_neg2(qg_fma2(D, rq, EmA));
    const float2 ipr = make_float2(qg_rcp(sp.pr.x), qg_rcp(sp.pr.y));
    const float2 io1 = make_float2(qg_rcp(o1.x), qg_rcp(o1.y)), io2 = make_float2(qg_rcp(o2.x), qg_rcp(o2.y));
    const float2 a2 = qg_bc(a), kT2 = qg_bc(kT);
    const float2 as = qg_mul2(qg_mul2(a2, qg_bc(sqrt2)), ipr);
    g.du1 = qg_mul2(qg_mul2(as, S1), io1);
    g.du2 = qg_mul2(qg_mul2(as, S2), io2);
    float2 H = qg_bc(0.0f);
    if (kT != 0.0f) {
        const float2 arg = qg_mul2(qg_mul2(sp.q, o1), o2);
        H = make_float2(__fadd_rn(const1, logf(arg.x)), __fadd_rn(const1, logf(arg.y)));
    }
    g.da = qg_fma2(kT2, H, E);
    g.do1 = qg_mul2(qg_mul2(a2, qg_add2(T1, kT2)), io1);
    g.do2 = qg_mul2(qg_mul2(a2, qg_add2(T2, kT2)), io2);
    g.dp = qg_mul2(qg_mul2(a2, qg_fma2(qg_neg2(kT2), p, Sp)), ipr);
    g.Ei = qg_mul2(a2, g.da);
    return g;
}

// Two edge quadratures at once (u and v layer of one direction), :118-146: the two layers share every table constant, so their
// samples run as one fp32x2 stream (5 FFMA2 + 2 MUFU per sample pair instead of 10 FFMA + 2 MUFU) into the packed epilogue.
template <int KT>
__device__ __forceinline__ QgGrad2 qg_edge2p(const QgTables &tab, int Krt, float a, float2 u1, float2 u2, float2 o1, float2 o2, float2 p,
                                             float lambdas, float epsn, float T)
{
    const float sqrt2 = 1.4142135623730951f, invpi = 0.31830988618379067f;
    const int K = KT > 0 ? KT : Krt;
    QgSpectral2 sp;
    sp.set(p);
    const float2 r2 = qg_bc(sqrt2);
    const float2 A = qg_mul2(r2, qg_fma2(o1, sp.s, qg_neg2(qg_mul2(o2, sp.t))));      // explicit FMAs: see qg_epilogue2
    const float2 B = qg_mul2(r2, qg_fma2(o1, sp.t, qg_neg2(qg_mul2(o2, sp.s))));
    const float2 d0 = qg_sub2(u1, u2), eps2 = qg_bc(epsn), zero = qg_bc(0.0f);
    float2 E = zero, MI = zero, MJ = zero, MII = zero, MJJ = zero, MB = zero;
#pragma unroll 1
    for (int r = 0; r < K; ++r) {
        const float2 dr = qg_fma2(B, qg_bc(tab.X[r]), d0);
        float2 S0 = zero, S1 = zero, S2 = zero;
        if (KT > 0) {
#pragma unroll
            for (int c = 0; c < (KT > 0 ? KT : 1); ++c) {
                const float2 d = qg_fma2(A, qg_bc(tab.X[c]), dr);
                const float2 q = qg_fma2(d, d, eps2);
                const float2 f = make_float2(qg_sqrt(q.x), qg_sqrt(q.y));
                S0 = qg_fma2(f, qg_bc(tab.W[c]), S0);
                S1 = qg_fma2(f, qg_bc(tab.WX[c]), S1);
                S2 = qg_fma2(f, qg_bc(tab.WXX[c]), S2);
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < K; ++c) {
                const float2 d = qg_fma2(A, qg_bc(tab.X[c]), dr);
                const float2 q = qg_fma2(d, d, eps2);
                const float2 f = make_float2(qg_sqrt(q.x), qg_sqrt(q.y));
                S0 = qg_fma2(f, qg_bc(tab.W[c]), S0);
                S1 = qg_fma2(f, qg_bc(tab.WX[c]), S1);
                S2 = qg_fma2(f, qg_bc(tab.WXX[c]), S2);
            }
        }
        const float wr = tab.W[r], wxr = tab.WX[r], wxxr = tab.WXX[r];
        E = qg_fma2(S0, qg_bc(wr), E);
        MI = qg_fma2(S1, qg_bc(wr), MI);
        MII = qg_fma2(S2, qg_bc(wr), MII);
        MJ = qg_fma2(S0, qg_bc(wxr), MJ);
        MB = qg_fma2(S1, qg_bc(wxr), MB);
        MJJ = qg_fma2(S0, qg_bc(wxxr), MJJ);
    }
    return qg_epilogue2(E, MI, MJ, MII, MJJ, MB, qg_bc(__fmul_rn(-lambdas, invpi)), sp, a, o1, o2, p, T);
}

// Node sample for a belief whose whole quadrature cloud lies inside the image: the clamps of :157-162 cannot fire, so the
// cell is addressed RELATIVE to the belief's own pixel (vv_mn = &VV8[m*pitchV + n]) and neither m nor n is live in the loop.
// The tap cache is keyed on the raw bits of the magic-number floor (equal bits <=> equal cell); the two 256-bit loads are
// predicated PTX, so the reload is never a branch and lanes that keep their cell keep their registers.
struct QgTapCacheRel {
    int kx, ky;
    QgTap8 v01, v23;
    __device__ __forceinline__ QgTapCacheRel() : kx(0), ky(0) {}     // 0 is not a valid key (keys are 0x4B400000 + small int)
};
__device__ __forceinline__ float qg_node_sample_inside(const QgTap8 *__restrict__ vv_mn, int pitchV, int koff, long long rowskip, float2 x,
                                                       float I1v, float epsn, QgTapCacheRel &tc)
{
    const float2 magic = qg_bc(12582912.0f);
    const float2 t = qg_add2_rm(x, magic);                            // floor(x) in the mantissa (qg_floor_split2)
    const float2 fr = qg_sub2(x, qg_sub2(t, magic));
    const int kx = __float_as_int(t.x), ky = __float_as_int(t.y);
    const int reload = (kx != tc.kx) | (ky != tc.ky);
    tc.kx = kx; tc.ky = ky;
    // (ky-C)*pitch + (kx-C) = ky*pitch + kx + koff; the address arithmetic is unconditional (five integer instructions), only
    // the two loads are predicated: lanes that keep their cell keep their registers
    const QgTap8 *a0 = vv_mn + (ky * pitchV + kx + koff);
    const QgTap8 *a1 = reinterpret_cast<const QgTap8 *>(reinterpret_cast<const char *>(a0) + rowskip);
    asm("{\n\t.reg .pred q;\n\t"
        "setp.ne.s32 q, %18, 0;\n\t"
        "@q ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        "@q ld.global.nc.v8.f32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t}"
        : "+f"(tc.v01.p[0].x), "+f"(tc.v01.p[0].y), "+f"(tc.v01.p[1].x), "+f"(tc.v01.p[1].y), "+f"(tc.v01.p[2].x), "+f"(tc.v01.p[2].y),
          "+f"(tc.v01.p[3].x), "+f"(tc.v01.p[3].y), "+f"(tc.v23.p[0].x), "+f"(tc.v23.p[0].y), "+f"(tc.v23.p[1].x), "+f"(tc.v23.p[1].y),
          "+f"(tc.v23.p[2].x), "+f"(tc.v23.p[2].y), "+f"(tc.v23.p[3].x), "+f"(tc.v23.p[3].y)
        : "l"(a0), "l"(a1), "r"(reload));
    float2 n0, w1, w2, n3;
    qg_cubic_w2(fr, n0, w1, w2, n3);
    float2 h01 = qg_mul2(tc.v01.p[0], qg_bc(-n0.x));
    float2 h23 = qg_mul2(tc.v23.p[0], qg_bc(-n0.x));
    h01 = qg_fma2(tc.v01.p[1], qg_bc(w1.x), h01);  h23 = qg_fma2(tc.v23.p[1], qg_bc(w1.x), h23);
    h01 = qg_fma2(tc.v01.p[2], qg_bc(w2.x), h01);  h23 = qg_fma2(tc.v23.p[2], qg_bc(w2.x), h23);
    h01 = qg_fma2(tc.v01.p[3], qg_bc(-n3.x), h01); h23 = qg_fma2(tc.v23.p[3], qg_bc(-n3.x), h23);
    const float v = fmaf(h23.y, -n3.y, fmaf(h23.x, w2.y, fmaf(h01.y, w1.y, h01.x * -n0.y)));
    const float d = fmaf(-0.25f, v, I1v);
    return qg_sqrt(fmaf(d, d, epsn));
}

// ---- wide beliefs: one sector per sample ----------------------------------------------------------------------------------
// While the beliefs are wide (the first thousands of iterations; for ever on frames whose mixture never settles) every sample of
// every lane lands in its own cell and the kernel is bound by the L1 data stage, not by arithmetic: a warp-wide gather of
// scattered 32-byte sectors retires about one sector per clock, and the 16 fp32 taps of a sample are two sectors
// (ncu, 4K, L=3, K=5: l1tex data-stage 84% busy, 64 sector wavefronts per warp-sample = the whole 4.2 ms).
// The reference's frames are grey levels -- double(rgb2gray(...)), optical_flow.m:8-11: integers 0..255 -- and getVV's quadratic
// border extrapolation 3a-3b+c (:198-207) keeps them integers below 2048, which fp16 holds EXACTLY.  For such frames the 4 x 4 tap
// block of cell (y,x) is stored as one 32-byte entry of fp16 values: ONE 256-bit load, one sector per sample and lane, and the
// values converted back to fp32 are bit-identical to the fp32 layout's.  Entry layout: word c (c = 0..3) = rows (y, y+1) of column
// x+c, word 4+c = rows (y+2, y+3): converted pairwise they are the float2 operands of the row-pair FFMA2 stream of qg_node_sample.
struct __align__(32) QgTap16h { unsigned int w[8]; };
__device__ __forceinline__ QgTap16h qg_ld256h(const QgTap16h *ptr) {
    QgTap16h v;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(ptr));
    return v;
}
__device__ __forceinline__ float2 qg_h2f2(unsigned int w) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
// vh_mn = &VVh[m*pitchV + n]; no tap cache (a wide belief changes cell with every sample): the load is unconditional, so the
// compiler is free to issue the loads of the next samples of the unrolled row ahead of the arithmetic of this one.
__device__ __forceinline__ float qg_node_sample_wide(const QgTap16h *__restrict__ vh_mn, int pitchV, int koff, float2 x, float I1v, float epsn)
{
    const float2 magic = qg_bc(12582912.0f);
    const float2 t = qg_add2_rm(x, magic);
    const float2 fr = qg_sub2(x, qg_sub2(t, magic));
    const QgTap16h e = qg_ld256h(vh_mn + (__float_as_int(t.y) * pitchV + __float_as_int(t.x) + koff));
    float2 n0, w1, w2, n3;
    qg_cubic_w2(fr, n0, w1, w2, n3);
    float2 h01 = qg_mul2(qg_h2f2(e.w[0]), qg_bc(-n0.x));
    float2 h23 = qg_mul2(qg_h2f2(e.w[4]), qg_bc(-n0.x));
    h01 = qg_fma2(qg_h2f2(e.w[1]), qg_bc(w1.x), h01);  h23 = qg_fma2(qg_h2f2(e.w[5]), qg_bc(w1.x), h23);
    h01 = qg_fma2(qg_h2f2(e.w[2]), qg_bc(w2.x), h01);  h23 = qg_fma2(qg_h2f2(e.w[6]), qg_bc(w2.x), h23);
    h01 = qg_fma2(qg_h2f2(e.w[3]), qg_bc(-n3.x), h01); h23 = qg_fma2(qg_h2f2(e.w[7]), qg_bc(-n3.x), h23);
    const float v = fmaf(h23.y, -n3.y, fmaf(h23.x, w2.y, fmaf(h01.y, w1.y, h01.x * -n0.y)));
    const float d = fmaf(-0.25f, v, I1v);
    return qg_sqrt(fmaf(d, d, epsn));
}
