// qgmap_walk.cuh -- the row-walking form of the fused QGMAP iteration kernel (full-resolution variant, one launch == one pass
// of gqmap_gpu_mixture.m:27-50,69-75).
//
// Why a second form: the tiled kernel of qgmap_iter.cuh spends one warp in eight on the halo row above its tile and parks
// warps at two CTA barriers (ncu, K=5: 2.7 of the ~7.6 warps a scheduler holds sit at a barrier; 19% of the memory-stall samples
// are the 17 dependent state loads at the head of every thread).  Here a CTA is ONE warp and owns a strip of 31 output columns
// (lane 0 = halo column left of it) x `strip_rows` rows, which it walks top to bottom:
//   * the down-edge endpoint-2 gradients of row m (gqmap_gpu_mixture.m:37-40) wait in shared memory for row m+1 of the same
//     thread -- no halo warp, no CTA barrier, no warp ever waits for another; the halo edge above the strip is recomputed once
//     per strip (1/strip_rows of a row's edge work instead of 1/7 of every CTA);
//   * the beliefs of row m+1, loaded for the down edge of row m, become the thread's own beliefs one row later, and the right
//     neighbour's beliefs arrive by warp shuffle: 10 state loads per pixel instead of 17, issued a whole row before their use;
//   * per row the four reduced scalars (:36,:48,:69-70) are warp-summed in fp32 (as in the tiled kernel) and accumulated in
//     fp64 per strip; strips -> strip rows -> launch are reduced by the last finisher of each level in a FIXED order, so the
//     result does not depend on which CTA happens to be last.
// Arithmetic per belief is that of qgmap_device.cuh (moments, fp32x2 streams, tap cache); two instruction-count cuts on top:
//   * the u- and v-layer epilogues of an edge direction run as one fp32x2 stream (qg_epilogue2);
//   * a warp whose K x K sample points provably stay inside the image (bounding box of the quadrature cloud, warp vote) runs a
//     node loop without the twelve clamp instructions per sample of :157-162 -- bit-identical by construction, the clamped
//     loop stays for warps near the image border or with wide beliefs.
#pragma once
#include "qgmap_iter.cuh"
#include "qgmap_peer.cuh"

#ifndef QG_WALK_MINB
#define QG_WALK_MINB 32      // resident one-warp CTAs per SM the kernel is compiled for (32 -> 64 registers)
#endif
#ifndef QG_WALK_MINB_BIGK
#define QG_WALK_MINB_BIGK 28 // K >= 7: 72 registers
#endif

template <int KT> struct QgWalk {
    static constexpr int MINB = (KT > 0 && KT <= 5) ? QG_WALK_MINB : QG_WALK_MINB_BIGK;
};

__device__ __forceinline__ double qg_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Strip -> strip row -> launch reduction of the four scalars; every level is summed by its last finisher in index order
// (lane-strided partial sums, then the xor tree), so the bits do not depend on the order the CTAs retire in.
template <bool DUMP>
__device__ __forceinline__ void qg_strip_finish(const QgIterParams &p, QgCtrl *ctrl, const double *acc /* shared, QG_NRED */, int lane)
{
    const unsigned int gx = gridDim.x, gy = gridDim.y;
    const unsigned int group = blockIdx.z * gy + blockIdx.y, strip = group * gx + blockIdx.x;
    if (lane < QG_NRED) {
        p.partials[(size_t)strip * QG_NRED + lane] = acc[lane];
        if (!DUMP) __threadfence();
    }
    if (DUMP) return;
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&p.tickets[group], 1u) == gx - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    {
        double a[QG_NRED] = {0.0, 0.0, 0.0, 0.0};
        const double *pp = p.partials + (size_t)group * gx * QG_NRED;
        for (unsigned int s = lane; s < gx; s += 32) {
#pragma unroll
            for (int k = 0; k < QG_NRED; ++k) a[k] += __ldcg(pp + (size_t)s * QG_NRED + k);
        }
#pragma unroll
        for (int k = 0; k < QG_NRED; ++k) {
            const double v = qg_warp_sum_d(a[k]);
            if (lane == k) p.gpartials[(size_t)group * QG_NRED + k] = v;
        }
        if (lane < QG_NRED) __threadfence();
        if (lane == 0) p.tickets[group] = 0;
    }
    __syncwarp();
    if (lane == 0) last = (atomicAdd(&ctrl->ticket, 1u) == gy * gridDim.z - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    __shared__ double sh_sum[QG_LMAX * QG_NRED];
    for (int ll = 0; ll < p.L; ++ll) {
        double a[QG_NRED] = {0.0, 0.0, 0.0, 0.0};
        const double *pp = p.gpartials + (size_t)ll * gy * QG_NRED;
        for (unsigned int s = lane; s < gy; s += 32) {
#pragma unroll
            for (int k = 0; k < QG_NRED; ++k) a[k] += __ldcg(pp + (size_t)s * QG_NRED + k);
        }
#pragma unroll
        for (int k = 0; k < QG_NRED; ++k) {
            const double v = qg_warp_sum_d(a[k]);
            if (lane == k) sh_sum[ll * QG_NRED + k] = v;
        }
    }
    __syncwarp();
    if (lane == 0) ctrl->ticket = 0;
    if (p.band == 2) {                       // row band over peer memory: post, wait for the other bands, add in rank order, advance
        qg_peer_finish(p, ctrl, p.peer, sh_sum, lane);
    } else if (lane == 0) {
        if (p.band) {
            for (int k = 0; k < p.L * QG_NRED; ++k) ctrl->sums[k] = sh_sum[k];   // summed across bands / ranks, then qg_advance
        } else {
            qg_advance(p, ctrl, sh_sum);
        }
    }
}

template <int KT, bool DUMP>
__global__ void __launch_bounds__(32, QgWalk<KT>::MINB)
qgmap_walk_kernel(const __grid_constant__ QgIterParams p)
{
    QgCtrl *ctrl = p.ctrl;
    if (!DUMP && ctrl->stop) return;
    const int it = ctrl->it;
    const int j = threadIdx.x;
    const int l = blockIdx.z;
    const int n = (int)blockIdx.x * (QG_TW - 1) + j;              // global column (lane 0 = halo column left of the strip)
    int sy = (int)blockIdx.y;                                     // strip rows in the order 0, last, 1, 2, ... (image / band
    { const int lasty = (int)gridDim.y - 1; sy = sy == 0 ? 0 : (sy == 1 ? lasty : sy - 1); }   // boundary strips first)
    const int m0 = p.out_r0 + sy * p.strip_rows;
    const int m1 = min(m0 + p.strip_rows, p.out_r1);
    const bool valid = n <= p.N - 1;                              // a column of the grid
    const bool incol = (n >= 1) && (n <= p.N - 2);                // interior column
    const bool is_out = (j >= 1) && incol;                        // every row of the strip is an updated row
    const bool need_right = is_out || ((j == 0) && (n + 1 <= p.N - 2));

    const float a = (float)ctrl->alpha[l];
    const float T = (float)ctrl->T;
    const float step = (float)(p.step0 / (1.0 + (double)it / p.step_tau));      // :27

    // State addressing: one moving per-thread pointer into field 0 of component l of the buffer this iteration READS (row m-1,
    // column n while walking); every other plane -- and the buffer it WRITES -- is that pointer plus a warp-uniform offset.
    const long long pl = p.plane;
    const long long fs = (long long)p.L * pl;                                   // field stride (floats)
    const long long od = p.buf[it & 1] - p.buf[(it - 1) & 1];                   // written buffer - read buffer (floats)
    const float *pin = p.buf[(it - 1) & 1] + (long long)l * pl + ((long long)(m0 - 1 - p.g0) * p.P + n);
    const int P = p.P;
#define QG_SRC(f, off) qg_lds(pin + ((long long)(f) * fs + (off)))
#define QG_DST(f) const_cast<float *>(pin)[od + (long long)(f) * fs]

    __shared__ float4 sh_up[QG_TW];                  // endpoint-2 gradients of the down edge of the row above, per column
    __shared__ float sh_st[10][QG_TW];               // parked across the node quadrature: next row's beliefs, folded edge gradients
    __shared__ double sh_acc[QG_NRED];
    if (j < QG_NRED) sh_acc[j] = 0.0;

    // The walk starts one row ABOVE the strip (halo row m0-1): there only the down edge is evaluated -- by the same code as in
    // every other row, so a belief's gradients do not depend on where the strips happen to be cut (bit-identical for any
    // strip_rows and any row-band decomposition) -- and its endpoint-2 gradients are left in sh_up for row m0.
    float2 mu_c = make_float2(0.f, 0.f), sg_c = make_float2(1.f, 1.f);
    if (valid) {
        mu_c = make_float2(QG_SRC(F_MUU, 0), QG_SRC(F_MUV, 0));
        sg_c = make_float2(QG_SRC(F_SIGU, 0), QG_SRC(F_SIGV, 0));
    }
    pin -= P;
    sh_up[j] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
    for (int m = m0 - 1; m < m1; ++m) {
        pin += P;                                                               // row m
        const bool halo = m < m0;                                               // warp-uniform
        float pn = 0.f;
        {
            // beliefs of the row below: neighbour of this row's down edges now, this thread's own beliefs one row later
            float2 mu_d = make_float2(0.f, 0.f), sg_d = make_float2(1.f, 1.f);
            if (valid) {
                mu_d = make_float2(QG_SRC(F_MUU, P), QG_SRC(F_MUV, P));
                sg_d = make_float2(QG_SRC(F_SIGU, P), QG_SRC(F_SIGV, P));
            }
            float2 rho_d = make_float2(0.f, 0.f), rho_r = make_float2(0.f, 0.f);
            if (is_out) rho_d = make_float2(QG_SRC(F_ROU0, 0), QG_SRC(F_ROU2, 0));
            QgGrad2 gd = {}, gr = {};
            if (is_out && m0 < m1)                                              // down edge (m,n)->(m+1,n), layers u,v  (:31-34, e=1)
                gd = qg_edge2p<KT>(p.tab, p.K, a, mu_c, mu_d, sg_c, sg_d, rho_d, p.lambdas, p.epsn, T);
            if (halo) {
                sh_up[j] = make_float4(gd.du2.x, gd.do2.x, gd.du2.y, gd.do2.y);
                mu_c = mu_d; sg_c = sg_d;
                continue;
            }
            sh_st[0][j] = mu_d.x; sh_st[1][j] = mu_d.y; sh_st[2][j] = sg_d.x; sh_st[3][j] = sg_d.y;
            // right neighbour's beliefs from the next lane; the strip's last lane reads column n+1 itself
            float2 mu_r = make_float2(__shfl_down_sync(0xffffffffu, mu_c.x, 1), __shfl_down_sync(0xffffffffu, mu_c.y, 1));
            float2 sg_r = make_float2(__shfl_down_sync(0xffffffffu, sg_c.x, 1), __shfl_down_sync(0xffffffffu, sg_c.y, 1));
            if (j == QG_TW - 1 && is_out) {
                mu_r = make_float2(QG_SRC(F_MUU, 1), QG_SRC(F_MUV, 1));
                sg_r = make_float2(QG_SRC(F_SIGU, 1), QG_SRC(F_SIGV, 1));
            }
            if (is_out) pn = QG_SRC(F_PN, 0);
            if (need_right) {                                                   // right edge (m,n)->(m,n+1)  (e=2)
                rho_r = make_float2(QG_SRC(F_ROU1, 0), QG_SRC(F_ROU3, 0));
                gr = qg_edge2p<KT>(p.tab, p.K, a, mu_c, mu_r, sg_c, sg_r, rho_r, p.lambdas, p.epsn, T);
            }
            // endpoint-2 gradients: of the right edge from the previous lane, of the down edge from the previous row
            const float lf_du_u = __shfl_up_sync(0xffffffffu, gr.du2.x, 1), lf_do_u = __shfl_up_sync(0xffffffffu, gr.do2.x, 1);
            const float lf_du_v = __shfl_up_sync(0xffffffffu, gr.du2.y, 1), lf_do_v = __shfl_up_sync(0xffffffffu, gr.do2.y, 1);
            const float4 up = sh_up[j];
            sh_up[j] = make_float4(gd.du2.x, gd.do2.x, gd.du2.y, gd.do2.y);
            if (is_out) {
                // :37-40 edge part of the assembled gradients, :48 / :36 edge parts of Energy and d(alpha)
                sh_st[4][j] = ((gd.du1.x + gr.du1.x) + up.x) + lf_du_u;
                sh_st[5][j] = ((gd.do1.x + gr.do1.x) + up.y) + lf_do_u;
                sh_st[6][j] = ((gd.du1.y + gr.du1.y) + up.z) + lf_du_v;
                sh_st[7][j] = ((gd.do1.y + gr.do1.y) + up.w) + lf_do_v;
                sh_st[8][j] = (gd.Ei.x + gr.Ei.x) + (gd.Ei.y + gr.Ei.y);
                sh_st[9][j] = (gd.da.x + gr.da.x) + (gd.da.y + gr.da.y);
                if (DUMP) {
                    float *d = p.dbg + (pin - p.buf[(it - 1) & 1]);
                    const long long fstr = (long long)p.L * pl;
                    d[5 * fstr] = gd.dp.x; d[6 * fstr] = gr.dp.x; d[7 * fstr] = gd.dp.y; d[8 * fstr] = gr.dp.y;
                } else {                                                          // :45
                    QG_DST(F_ROU0) = qg_clamp(fmaf(gd.dp.x, step, rho_d.x), -p.corr_tor, p.corr_tor);
                    QG_DST(F_ROU1) = qg_clamp(fmaf(gr.dp.x, step, rho_r.x), -p.corr_tor, p.corr_tor);
                    QG_DST(F_ROU2) = qg_clamp(fmaf(gd.dp.y, step, rho_d.y), -p.corr_tor, p.corr_tor);
                    QG_DST(F_ROU3) = qg_clamp(fmaf(gr.dp.y, step, rho_r.y), -p.corr_tor, p.corr_tor);
                    if (p.band == 2 && (m == p.pub_row[0] || m == p.pub_row[1]))            // band boundary row: also into the
                        qg_publish_row(p, it, m, n, l, &QG_DST(0), fs, F_ROU0, 4);          // neighbours' halo rows (NVLink stores)
                }
            }
        }
        // ---- node term set-up (:87-93) and the bounding box of the K x K sample cloud around the mean:
        //      |x - mu| <= sqrt2 * sigma * (|s| + |t|) * X_max.  Inside the image for the whole warp -> no clamps needed.
        const int lastx = p.No - 2, lasty = p.Mo - 2;
        bool inside = true;
        {
            QgSpectral sp;
            sp.set(pn);
            if (is_out) {
                const int Kq = KT > 0 ? KT : p.K;
                const float reach = 1.4142135623730951f * p.tab.X[Kq - 1] * (fabsf(sp.s) + fabsf(sp.t)) * 1.0001f;
                const float ex = fmaf(reach, sg_c.x, 0.01f), ey = fmaf(reach, sg_c.y, 0.01f);
                inside = (mu_c.x - ex >= (float)(-n)) && (mu_c.x + ex <= (float)(lastx - n)) &&
                         (mu_c.y - ey >= (float)(-m)) && (mu_c.y + ey <= (float)(lasty - m));
            }
        }
        const bool all_inside = __all_sync(0xffffffffu, inside);

        float red[QG_NRED] = {0.f, 0.f, 0.f, 0.f};
        if (is_out) {
            // ---- node quadrature (:29, :94-106) -------------------------------------------------------------------------
            const float I1v = __ldg(p.I1 + (long long)m * p.pitchI + n);
            QgMoments mo;
            {
                QgSpectral sp;
                sp.set(pn);
                if (all_inside) {
                    QgTapCacheRel tc;
                    const QgTap8 *vv_mn = p.VV8 + (long long)m * p.pitchV + n;
                    const int koff = -0x4B400000 * (p.pitchV + 1);
                    const long long rowskip = (long long)p.pitchV * (2 * (long long)sizeof(QgTap8));
                    mo = qg_quadrature<KT>(p.tab, p.K, mu_c.x, mu_c.y, sg_c.x, sg_c.y, sp, -p.lambdad, [&](float2 x) {
                        return qg_node_sample_inside(vv_mn, p.pitchV, koff, rowskip, x, I1v, p.epsn, tc);
                    });
                } else {
                    QgTapCache tc;
                    mo = qg_quadrature<KT>(p.tab, p.K, mu_c.x, mu_c.y, sg_c.x, sg_c.y, sp, -p.lambdad, [&](float2 x) {
                        return qg_node_sample<(KT >= 7)>(p.VV8, p.pitchV, m, n, lastx, lasty, x, I1v, p.epsn, tc);
                    });
                }
            }
            QgSpectral sp;                                                        // recomputed (12 instructions) rather than kept
            sp.set(pn);                                                           // live across the quadrature
            const QgGrad gn = qg_epilogue(mo, sp, a, sg_c.x, sg_c.y, pn, -3.0f * T);

            const float G_muu = gn.du1 + sh_st[4][j], G_sigu = gn.do1 + sh_st[5][j];
            const float G_muv = gn.du2 + sh_st[6][j], G_sigv = gn.do2 + sh_st[7][j];
            const float e_px = gn.Ei + sh_st[8][j];                               // :48
            const float da_px = gn.da + sh_st[9][j];                              // :36
            red[0] = e_px; red[1] = da_px; red[2] = fabsf(G_muu); red[3] = fabsf(G_sigu);
            if (DUMP) {
                float *d = p.dbg + (pin - p.buf[(it - 1) & 1]);
                const long long fstr = (long long)p.L * pl;
                d[0 * fstr] = G_muu; d[1 * fstr] = G_muv; d[2 * fstr] = G_sigu; d[3 * fstr] = G_sigv;
                d[4 * fstr] = gn.dp; d[9 * fstr] = e_px; d[10 * fstr] = da_px;
            } else {                                                              // :41-44, :46
                QG_DST(F_MUU) = qg_clamp(fmaf(G_muu, step, mu_c.x), p.minu, p.maxu);
                QG_DST(F_MUV) = qg_clamp(fmaf(G_muv, step, mu_c.y), p.minv, p.maxv);
                QG_DST(F_SIGU) = qg_clamp(fmaf(G_sigu, step * p.sig_step, sg_c.x), p.sig_min, p.sig_max);
                QG_DST(F_SIGV) = qg_clamp(fmaf(G_sigv, step * p.sig_step, sg_c.y), p.sig_min, p.sig_max);
                QG_DST(F_PN) = qg_clamp(fmaf(gn.dp, step, pn), -p.corr_tor, p.corr_tor);
                if (p.band == 2 && (m == p.pub_row[0] || m == p.pub_row[1])) {
                    qg_publish_row(p, it, m, n, l, &QG_DST(0), fs, F_MUU, 5);
                    __threadfence_system();
                }
            }
        }
        __syncwarp();
        // per row: fp32 warp sums, accumulated in fp64 over the strip's rows (:36,:48,:69-70)
#pragma unroll
        for (int k = 0; k < QG_NRED; ++k) {
            const float v = qg_warp_sum(red[k]);
            if (j == k) sh_acc[k] += (double)v;
        }
        mu_c = make_float2(sh_st[0][j], sh_st[1][j]);
        sg_c = make_float2(sh_st[2][j], sh_st[3][j]);
    }
    __syncwarp();
    qg_strip_finish<DUMP>(p, ctrl, sh_acc, j);
#undef QG_SRC
#undef QG_DST
}
