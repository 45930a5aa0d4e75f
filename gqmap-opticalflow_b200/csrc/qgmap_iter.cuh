// qgmap_iter.cuh -- the fused QGMAP iteration kernel (one launch == one pass of gqmap_gpu_mixture.m:27-50,69-75).
//
// One thread per (belief pixel, mixture component): node quadrature (:87-116), the pixel's down and right edge
// quadratures for both flow layers (:118-146), gradient assembly with the neighbours' endpoint-2 contributions
// (:36-40), ascent step + clamps (:41-46), and the block partial sums for Energy / d(alpha) / mean|grad|
// (:36,:48,:69-70).  The last block to finish reduces the partials in a fixed order and advances the device-resident
// control block (step counter, alpha softmax / projsplx update :50,:78-86, temperature anneal S:72, stop test :75).
//
// Tiling: a CTA is 32 lanes x (TH+1) warps (TH = QgTile<K,SUPER>::TH).  Warp 0 is the halo row above the tile and lane 0 the halo column left
// of it: they evaluate only the edge whose endpoint-2 gradient an output pixel needs (Jacobi semantics: every gradient
// uses the OLD state, written state goes to the other ping-pong buffer).  Down-edge endpoint-2 gradients travel through
// shared memory to the warp below, right-edge ones through a warp shuffle to the next lane.  Output tile = 31 x TH.
#pragma once
#include "qgmap_device.cuh"
#include "qgmap_advance.cuh"
#include "qgmap_peer.cuh"

__device__ __forceinline__ float qg_clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// Programmatic dependent launch (qgmap_handle::pdl): iteration t+1 is launched while iteration t still runs and waits HERE, before it
// reads anything iteration t writes (control block, beliefs), until that grid has completed and flushed; it then lets iteration t+2
// be launched.  Both instructions are no-ops in a launch without the programmatic-serialization attribute.
__device__ __forceinline__ void qg_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Tile rows are issued in the order 0, last, 1, 2, ...: tiles touching the top/bottom image border hold the samples that get
// clamped to the image (the slow path of the super-pixel variant); starting them first keeps them out of the launch's tail.
__device__ __forceinline__ int qg_tile_row() {
    const int by = (int)blockIdx.y, last = (int)gridDim.y - 1;
    return by == 0 ? 0 : (by == 1 ? last : by - 1);
}

__device__ __forceinline__ float qg_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block partial sums (fp32 within a warp, fp64 across warps and blocks) of red[] = {energy, dalpha, |G_muu|, |G_sigu|}; the
// last block to finish reduces all partials in a fixed order and advances the control block (:36,:48,:50,:69-75).
// NW = warps of the CTA, W0 = first warp that owns output rows (1 when warp 0 is the halo row).
template <bool DUMP, int NW, int W0>
__device__ __forceinline__ void qg_block_finish_w(const QgIterParams &p, QgCtrl *ctrl, const float (&red)[QG_NRED], int r, int j)
{
    __shared__ double sh_red[NW][QG_NRED];
    __shared__ int sh_last;
#pragma unroll
    for (int k = 0; k < QG_NRED; ++k) {
        float v = qg_warp_sum(red[k]);
        if (j == 0) sh_red[r][k] = (double)v;
    }
    __syncthreads();
    const int tid = r * QG_TW + j;
    const unsigned int nblk_l = gridDim.x * gridDim.y;
    const unsigned int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (tid < QG_NRED) {
        double s = 0.0;
#pragma unroll
        for (int w = W0; w < NW; ++w) s += sh_red[w][tid];
        p.partials[(size_t)blk * QG_NRED + tid] = s;
    }
    if (DUMP) return;

    // Ticket: the four writers and the ticket thread sit in warp 0, so a warp barrier orders their stores before lane 0's RELEASE
    // increment (cumulative over the barrier) -- no CTA-wide barrier between the stores and the ticket, and no sequentially-consistent
    // fence (the MEMBAR.SC + the extra barrier were 4-7% of the warp samples, profiles/r02_iter_full_L3K5_*_v12_*).
    if (tid < 32) {
        __syncwarp();
        if (tid == 0) {
            unsigned int t;
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(&ctrl->ticket) : "memory");
            sh_last = (t == nblk_l * gridDim.z - 1);
        }
    }
    __syncthreads();
    if (!sh_last) return;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire side: every block's partials are visible after its ticket
    __shared__ double sh_sum[QG_LMAX * QG_NRED];
    const int nthr = QG_TW * NW, warp = tid >> 5, lane = tid & 31, nwarp = nthr / 32;
    for (int ll = 0; ll < p.L; ++ll) {
        double acc[QG_NRED] = {0.0, 0.0, 0.0, 0.0};
        const double *pp = p.partials + (size_t)ll * nblk_l * QG_NRED;
        for (unsigned int b = tid; b < nblk_l; b += nthr) {
#pragma unroll
            for (int k = 0; k < QG_NRED; ++k) acc[k] += __ldcg(pp + (size_t)b * QG_NRED + k);
        }
#pragma unroll
        for (int k = 0; k < QG_NRED; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sh_red[warp][k] = v;
        }
        __syncthreads();
        if (tid < QG_NRED) {
            double s = 0.0;
            for (int w = 0; w < nwarp; ++w) s += sh_red[w][tid];
            sh_sum[ll * QG_NRED + tid] = s;
        }
        __syncthreads();
    }
    if (tid == 0) ctrl->ticket = 0;
    if (p.band == 2) {                       // row band over peer memory: post, wait for the other bands, add in rank order, advance
        if (tid < 32) qg_peer_finish(p, ctrl, p.peer, sh_sum, tid);
    } else if (tid == 0) {
        if (p.band) {
            for (int k = 0; k < p.L * QG_NRED; ++k) ctrl->sums[k] = sh_sum[k];   // summed across bands / ranks, then
        } else {                                                               // the advance kernel runs qg_advance
            qg_advance(p, ctrl, sh_sum);
        }
    }
}

template <bool DUMP, int TH>
__device__ __forceinline__ void qg_block_finish(const QgIterParams &p, QgCtrl *ctrl, const float (&red)[QG_NRED], int r, int j)
{
    qg_block_finish_w<DUMP, TH + 1, 1>(p, ctrl, red, r, j);
}

// Row band over peer memory (QgIterParams::band == 2): the updated beliefs of the band's first / last row also go straight into
// the neighbouring band's halo row of the same ping-pong buffer (NVLink stores; qgmap_peer.cuh).  f0..f0+nf-1: fields to copy.
__device__ __forceinline__ void qg_publish_row(const QgIterParams &p, int it, int m, int n, int l, const float *o, long long fstr,
                                               int f0, int nf)
{
    const QgPeer *q = p.peer;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        float *nb = side ? q->dn[it & 1] : q->up[it & 1];
        if (!nb || m != p.pub_row[side]) continue;
        const long long npl = side ? q->dn_plane : q->up_plane;
        float *t = nb + (side ? q->dn_off : q->up_off) + n + (long long)l * npl;
        const long long nfs = (long long)p.L * npl;
        for (int f = f0; f < f0 + nf; ++f) t[f * nfs] = o[f * fstr];
    }
}

template <int KT, bool SUPER, bool DUMP>
__global__ void __launch_bounds__(QG_TW *(QgTile<KT, SUPER>::TH + QgTile<KT, SUPER>::W0), QgTile<KT, SUPER>::MINB)
qgmap_iter_kernel(const __grid_constant__ QgIterParams p)
{
    QgCtrl *ctrl = p.ctrl;
    qg_pdl_enter();
    if (!DUMP && ctrl->stop) return;
    const int it = ctrl->it;
    const float *__restrict__ in = p.buf[(it - 1) & 1];
    float *__restrict__ out = p.buf[it & 1];

    constexpr int TH = QgTile<KT, SUPER>::TH, W0 = QgTile<KT, SUPER>::W0;    // W0 = 1: warp 0 is the halo row above the tile
    const int j = threadIdx.x, r = threadIdx.y;
    const int l = blockIdx.z;
    const int n = (int)blockIdx.x * (QG_TW - 1) + j;              // global column (lane 0 = halo column n0-1)
    const int m = p.out_r0 + qg_tile_row() * TH + r - W0;         // global row    (with a halo warp: warp 0 = row m0-1)
    const bool incol = (n >= 1) && (n <= p.N - 2);                // interior column
    const bool inrow = (m >= p.out_r0) && (m < p.out_r1);         // row this handle updates (interior by construction)
    const bool is_out = (r >= W0) && (j >= 1) && inrow && incol;
    // halo warp: down edge of the row above the tile; halo lane: right edge of the column left of the tile
    const bool need_down = is_out || (W0 && (r == 0) && (j >= 1) && incol && (m + 1 < p.out_r1));
    const bool need_right = is_out || ((j == 0) && (r >= W0) && inrow && (n + 1 <= p.N - 2));
    const bool up_edge = !W0 && (r == 0) && is_out;               // no halo warp: the first row evaluates the down edge above it too
    const bool active = need_down || need_right;

    const float a = (float)ctrl->alpha[l];
    const float T = (float)ctrl->T;
    const float step = (float)(p.step0 / (1.0 + (double)it / p.step_tau));      // :27

    const long long pl = p.plane;
    const float *base = in + (long long)l * pl;
    const long long idx = (long long)(m - p.g0) * p.P + n;
    const long long fstr = (long long)p.L * pl;                                 // field stride

    float muu = 0.f, muv = 0.f, sigu = 1.f, sigv = 1.f;
    if (active) {
        muu = qg_lds(base + F_MUU * fstr + idx);  muv = qg_lds(base + F_MUV * fstr + idx);
        sigu = qg_lds(base + F_SIGU * fstr + idx); sigv = qg_lds(base + F_SIGV * fstr + idx);
    }

    QgGrad2 gd = {}, gr = {};
    float rou0 = 0.f, rou1 = 0.f, rou2 = 0.f, rou3 = 0.f;
    float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
    float lf_du_u = 0.f, lf_do_u = 0.f, lf_du_v = 0.f, lf_do_v = 0.f;
    __shared__ float4 sh_dn[TH + W0][QG_TW];

    // ---- down edge (m,n)->(m+1,n), layers u and v as one fp32x2 stream  (:31-34, e=1) ----------------------------------
    if (need_down) {
        const long long idn = idx + p.P;
        rou0 = qg_lds(base + F_ROU0 * fstr + idx);
        rou2 = qg_lds(base + F_ROU2 * fstr + idx);
        gd = qg_edge2p<KT>(p.tab, p.K, a, make_float2(muu, muv),
                           make_float2(qg_lds(base + F_MUU * fstr + idn), qg_lds(base + F_MUV * fstr + idn)), make_float2(sigu, sigv),
                           make_float2(qg_lds(base + F_SIGU * fstr + idn), qg_lds(base + F_SIGV * fstr + idn)), make_float2(rou0, rou2),
                           p.lambdas, p.epsn, T);
    }
    // ---- right edge (m,n)->(m,n+1)  (e=2) -----------------------------------------------------------------------
    if (need_right) {
        const long long irt = idx + 1;
        rou1 = qg_lds(base + F_ROU1 * fstr + idx);
        rou3 = qg_lds(base + F_ROU3 * fstr + idx);
        gr = qg_edge2p<KT>(p.tab, p.K, a, make_float2(muu, muv),
                           make_float2(qg_lds(base + F_MUU * fstr + irt), qg_lds(base + F_MUV * fstr + irt)), make_float2(sigu, sigv),
                           make_float2(qg_lds(base + F_SIGU * fstr + irt), qg_lds(base + F_SIGV * fstr + irt)), make_float2(rou1, rou3),
                           p.lambdas, p.epsn, T);
    }
    // ---- endpoint-2 exchange (before the node term: the warps of a CTA then never wait for each other again until the
    //      final block reduction) ----------------------------------------------------------------------------------------
    sh_dn[r][j] = make_float4(gd.du2.x, gd.do2.x, gd.du2.y, gd.do2.y);          // to pixel (m+1,n)
    lf_du_u = __shfl_up_sync(0xffffffffu, gr.du2.x, 1);                         // from pixel (m,n-1)
    lf_do_u = __shfl_up_sync(0xffffffffu, gr.do2.x, 1);
    lf_du_v = __shfl_up_sync(0xffffffffu, gr.du2.y, 1);
    lf_do_v = __shfl_up_sync(0xffffffffu, gr.do2.y, 1);
    __syncthreads();
    // ---- tiles without a halo warp: the first row evaluates the down edge (m-1,n)->(m,n) itself, for its endpoint-2 part -- after the
    //      barrier, so that no other row waits for it.  This is a second inlined copy of qg_edge2p; the file is compiled with
    //      -fmad=false (csrc/Makefile) so that both copies round identically: an edge's value must not depend on which tile row
    //      evaluates it (bit-identical results for every band split; with contraction left to ptxas the two copies differed in do2).
    if (up_edge) {
        const long long iup = idx - p.P;
        const QgGrad2 gh = qg_edge2p<KT>(p.tab, p.K, a,
                                         make_float2(qg_lds(base + F_MUU * fstr + iup), qg_lds(base + F_MUV * fstr + iup)), make_float2(muu, muv),
                                         make_float2(qg_lds(base + F_SIGU * fstr + iup), qg_lds(base + F_SIGV * fstr + iup)), make_float2(sigu, sigv),
                                         make_float2(qg_lds(base + F_ROU0 * fstr + iup), qg_lds(base + F_ROU2 * fstr + iup)),
                                         p.lambdas, p.epsn, T);
        up = make_float4(gh.du2.x, gh.do2.x, gh.du2.y, gh.do2.y);
    }

    // ---- node term set-up (:87-93).  Full resolution: the bounding box of the K x K sample cloud around the mean,
    //      |x - mu| <= sqrt2 * sigma * (|s| + |t|) * X_max; inside the image for every belief of the warp -> the clamps of
    //      :157-162 cannot fire and the warp runs the clamp-free sample loop (relative cell addressing, qg_node_sample_inside).
    const float pn = is_out ? qg_lds(base + F_PN * fstr + idx) : 0.f;
    QgSpectral sp;
    sp.set(pn);
    //      Wide -- the cloud reaches more than a pixel to either side, so nearly every sample lands in a new cell -- for at least half
    //      of the warp's beliefs, on a frame that has the fp16 4 x 4-block layout -> one-sector gather without a tap cache
    //      (qg_node_sample_wide); narrower warps re-use taps often enough for the cache to win (profiles/r02_paths_ab.txt).
    bool all_inside = false, go_wide = false;
    if (!SUPER) {
        bool inside = true, wide = false;
        if (is_out) {
            const int Kq = KT > 0 ? KT : p.K;
            const float reach = 1.4142135623730951f * p.tab.X[Kq - 1] * (fabsf(sp.s) + fabsf(sp.t)) * 1.0001f;
            const float ex = fmaf(reach, sigu, 0.01f), ey = fmaf(reach, sigv, 0.01f);
            inside = (muu - ex >= (float)(-n)) && (muu + ex <= (float)(p.No - 2 - n)) &&
                     (muv - ey >= (float)(-m)) && (muv + ey <= (float)(p.Mo - 2 - m));
            wide = fmaxf(ex, ey) >= p.wide_reach;
        }
        all_inside = __all_sync(0xffffffffu, inside);
        go_wide = all_inside && p.VVh && __popc(__ballot_sync(0xffffffffu, wide)) >= 16;
    }

    float red[QG_NRED] = {0.f, 0.f, 0.f, 0.f};
    if (is_out) {
        // :37-40 edge part of the assembled gradients: sum_e d1 + shifted d2 (down edge of (m-1,n), right edge of (m,n-1)).
        // Folded into 6 scalars now so that little stays live across the node quadrature.
        if (W0 || r > 0) up = sh_dn[r - 1][j];
        const float E_muu = ((gd.du1.x + gr.du1.x) + up.x) + lf_du_u;
        const float E_sigu = ((gd.do1.x + gr.do1.x) + up.y) + lf_do_u;
        const float E_muv = ((gd.du1.y + gr.du1.y) + up.z) + lf_du_v;
        const float E_sigv = ((gd.do1.y + gr.do1.y) + up.w) + lf_do_v;
        const float E_e = (gd.Ei.x + gr.Ei.x) + (gd.Ei.y + gr.Ei.y);             // :48 edge part
        const float E_a = (gd.da.x + gr.da.x) + (gd.da.y + gr.da.y);             // :36 edge part
        float *o = out + (long long)l * pl + idx;
        float *d = DUMP ? p.dbg + (long long)l * pl + idx : nullptr;
        if (DUMP) {
            d[5 * fstr] = gd.dp.x; d[6 * fstr] = gr.dp.x; d[7 * fstr] = gd.dp.y; d[8 * fstr] = gr.dp.y;
        } else {                                                                 // :45
            o[F_ROU0 * fstr] = qg_clamp(fmaf(gd.dp.x, step, rou0), -p.corr_tor, p.corr_tor);
            o[F_ROU1 * fstr] = qg_clamp(fmaf(gr.dp.x, step, rou1), -p.corr_tor, p.corr_tor);
            o[F_ROU2 * fstr] = qg_clamp(fmaf(gd.dp.y, step, rou2), -p.corr_tor, p.corr_tor);
            o[F_ROU3 * fstr] = qg_clamp(fmaf(gr.dp.y, step, rou3), -p.corr_tor, p.corr_tor);
            if (p.band == 2 && (m == p.pub_row[0] || m == p.pub_row[1])) qg_publish_row(p, it, m, n, l, o, fstr, F_ROU0, 4);
        }

        // ---- node quadrature (:29, :94-106) ---------------------------------------------------------------------------
        QgMoments mo;
        if (SUPER) {
            float I1b[16];
            const float *ip = p.I1 + (long long)(4 * m) * p.pitchI + 4 * n;
#pragma unroll
            for (int di = 0; di < 4; ++di) {
                float4 v = __ldg(reinterpret_cast<const float4 *>(ip + (long long)di * p.pitchI));
                I1b[di * 4 + 0] = v.x; I1b[di * 4 + 1] = v.y; I1b[di * 4 + 2] = v.z; I1b[di * 4 + 3] = v.w;
            }
            const int lastx = p.No - 2, lasty = p.Mo - 2, m4 = 4 * m, n4 = 4 * n;
            mo = qg_quadrature<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                return qg_super_sample(p.VV8, p.pitchV, m4, n4, lastx, lasty, x, I1b, p.epsn);
            });
        } else {
            const float I1v = __ldg(p.I1 + (long long)m * p.pitchI + n);
            if (go_wide) {                             // wide beliefs on an fp16-exact frame: one sector per sample, no tap cache
                const QgTap16h *vh_mn = p.VVh + (long long)m * p.pitchV + n;
                const int koff = -0x4B400000 * (p.pitchV + 1);
                mo = qg_quadrature<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                    return qg_node_sample_wide(vh_mn, p.pitchV, koff, x, I1v, p.epsn);
                });
            } else if (all_inside) {
                QgTapCacheRel tc;
                const QgTap8 *vv_mn = p.VV8 + (long long)m * p.pitchV + n;
                const int koff = -0x4B400000 * (p.pitchV + 1);
                const long long rowskip = (long long)p.pitchV * (2 * (long long)sizeof(QgTap8));
                mo = qg_quadrature<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                    return qg_node_sample_inside(vv_mn, p.pitchV, koff, rowskip, x, I1v, p.epsn, tc);
                });
            } else {
                const int lastx = p.No - 2, lasty = p.Mo - 2;
                QgTapCache tc;
                mo = qg_quadrature<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                    return qg_node_sample<(KT >= 7)>(p.VV8, p.pitchV, m, n, lastx, lasty, x, I1v, p.epsn, tc);
                });
            }
        }
        const QgGrad gn = qg_epilogue(mo, sp, a, sigu, sigv, pn, -3.0f * T);

        const float G_muu = gn.du1 + E_muu, G_sigu = gn.do1 + E_sigu;
        const float G_muv = gn.du2 + E_muv, G_sigv = gn.do2 + E_sigv;
        const float e_px = gn.Ei + E_e;                                          // :48
        const float da_px = gn.da + E_a;                                         // :36
        red[0] = e_px; red[1] = da_px; red[2] = fabsf(G_muu); red[3] = fabsf(G_sigu);
        if (DUMP) {
            d[0 * fstr] = G_muu; d[1 * fstr] = G_muv; d[2 * fstr] = G_sigu; d[3 * fstr] = G_sigv;
            d[4 * fstr] = gn.dp; d[9 * fstr] = e_px; d[10 * fstr] = da_px;
        } else {                                                                 // :41-44, :46
            o[F_MUU * fstr] = qg_clamp(fmaf(G_muu, step, muu), p.minu, p.maxu);
            o[F_MUV * fstr] = qg_clamp(fmaf(G_muv, step, muv), p.minv, p.maxv);
            o[F_SIGU * fstr] = qg_clamp(fmaf(G_sigu, step * p.sig_step, sigu), p.sig_min, p.sig_max);
            o[F_SIGV * fstr] = qg_clamp(fmaf(G_sigv, step * p.sig_step, sigv), p.sig_min, p.sig_max);
            o[F_PN * fstr] = qg_clamp(fmaf(gn.dp, step, pn), -p.corr_tor, p.corr_tor);
            if (p.band == 2 && (m == p.pub_row[0] || m == p.pub_row[1])) {
                qg_publish_row(p, it, m, n, l, o, fstr, F_MUU, 5);
                __threadfence_system();
            }
        }
    }

    qg_block_finish_w<DUMP, TH + W0, W0>(p, ctrl, red, r, j);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Four lanes per (belief pixel, component).  Used when the belief grid is too small to fill 148 SMs with one thread per
// belief (the super-pixel variant has 1/16 of the beliefs but 16x the work per node sample; small frames with few
// components): lane g of a group evaluates edge quadrature q = g (q = e + 2c: down-u, right-u, down-v, right-v -- the four
// are perfectly balanced) and the XJ rows r = g, g+4, ... of the node quadrature; the six node moments are combined with
// two xor-shuffles, the edge gradients with one or two (north_star: "warp-shuffle reductions of the gradients").
// A warp covers 8 columns (column group 0 = halo column); tile = 7 x TH outputs.
// No halo WARP here (QG_G4_HALO_WARP = 0): a lane's node share is ~9x its edge quadrature, so the first row's down-edge lanes evaluate
// the down edge of the row above the tile as a second quadrature (+9% on one warp of TH) instead of a warp that idles after its edges
// (1 of 5 warp slots in the super-pixel tile; profiles/r02_iter_super_L3K5_480x640_v12_g4_it3000_ncu_full.txt: 24% of the warp samples
// at the block-reduction barrier).  The one-thread-per-belief kernel keeps its halo warp: there the extra edge is +19% (r01 experiment 7).
#define QG_G 4
#define QG_CW (QG_TW / QG_G)
#ifndef QG_G4_HALO_WARP
#define QG_G4_HALO_WARP 0
#endif
#define QG_G4_W0 (QG_G4_HALO_WARP ? 1 : 0)                          // first warp that owns an output row
template <int KT, bool SUPER> struct QgTileG4 {
    static constexpr int TH = QgTile<KT, SUPER>::TH;
    static constexpr int NW = TH + QG_G4_W0;                       // warps per CTA
    static constexpr int MINB = QG_G4_HALO_WARP ? QgTile<KT, SUPER>::MINB : (SUPER ? QG_SUPER_MINB_NOHALO : QgTile<KT, SUPER>::MINB);
};

template <int KT, bool SUPER, bool DUMP>
__global__ void __launch_bounds__(QG_TW * QgTileG4<KT, SUPER>::NW, QgTileG4<KT, SUPER>::MINB)
qgmap_iter_kernel_g4(const __grid_constant__ QgIterParams p)
{
    QgCtrl *ctrl = p.ctrl;
    qg_pdl_enter();
    if (!DUMP && ctrl->stop) return;
    const int it = ctrl->it;
    const float *__restrict__ in = p.buf[(it - 1) & 1];
    float *__restrict__ out = p.buf[it & 1];

    const int j = threadIdx.x, r = threadIdx.y;
    const int jc = j / QG_G, g = j % QG_G;                         // column group within the warp, lane within the group
    const int e = g & 1, c = g >> 1;                               // this lane's edge: e 0=down 1=right, layer c 0=u 1=v
    const int l = blockIdx.z;
    constexpr int TH = QgTile<KT, SUPER>::TH;
    const int n = (int)blockIdx.x * (QG_CW - 1) + jc;              // global column (group 0 = halo column n0-1)
    const int m = p.out_r0 + qg_tile_row() * TH + r - QG_G4_W0;    // global row    (with a halo warp: warp 0 = halo row m0-1)
    const bool incol = (n >= 1) && (n <= p.N - 2);
    const bool inrow = (m >= p.out_r0) && (m < p.out_r1);
    const bool is_out = (r >= QG_G4_W0) && (jc >= 1) && inrow && incol;
    const bool need_down = is_out || (QG_G4_HALO_WARP && (r == 0) && (jc >= 1) && incol && (m + 1 < p.out_r1));
    const bool need_right = is_out || ((jc == 0) && (r >= QG_G4_W0) && inrow && (n + 1 <= p.N - 2));
    const bool my_edge = e ? need_right : need_down;

    const float a = (float)ctrl->alpha[l];
    const float T = (float)ctrl->T;
    const float step = (float)(p.step0 / (1.0 + (double)it / p.step_tau));      // :27
    const long long pl = p.plane;
    const float *base = in + (long long)l * pl;
    const long long idx = (long long)(m - p.g0) * p.P + n;
    const long long fstr = (long long)p.L * pl;

    // ---- this lane's edge quadrature (:31-34).  Without a halo warp the tile's first row evaluates a second one AFTER the exchange
    //      barrier (producer and consumer are the same thread, so the other rows do not wait for it): the down edge (m-1,n)->(m,n) of
    //      its layer, for the endpoint-2 part.  Both run through ONE call site -- a two-trip loop with the barrier in the middle --
    //      because two inlined copies of qg_edge may contract their FMAs differently, and an edge's value must not depend on which
    //      tile row evaluates it (bit-identical results for every band split).
    __shared__ float sh_dn[QgTileG4<KT, SUPER>::NW][QG_CW][4];
    QgGrad ge = {};
    float rouq = 0.f, up_du = 0.f, up_do = 0.f, lf_du = 0.f, lf_do = 0.f;
    const float *mu = base + (c ? F_MUV : F_MUU) * fstr, *sg = base + (c ? F_SIGV : F_SIGU) * fstr;
    const bool up_edge = !QG_G4_HALO_WARP && (r == 0) && (e == 0) && is_out;
#pragma unroll 1
    for (int pass = 0; pass < (QG_G4_HALO_WARP ? 1 : 2); ++pass) {
        const bool on = pass ? up_edge : my_edge;
        const long long i1 = pass ? idx - p.P : idx;                             // edge origin
        const long long i2 = pass ? idx : idx + (e ? 1 : p.P);                   // edge end point
        QgGrad gq = {};
        float rq = 0.f;
        if (on) {
            rq = __ldg(base + (F_ROU0 + g) * fstr + i1);
            gq = qg_edge<KT>(p.tab, p.K, a, __ldg(mu + i1), __ldg(mu + i2), __ldg(sg + i1), __ldg(sg + i2), rq, p.lambdas, p.epsn, T);
        }
        if (pass == 0) {
            ge = gq; rouq = rq;
            // endpoint-2 exchange: down edges through shared memory to the row below, right edges by shuffle to the next group
            if (e == 0) { sh_dn[r][jc][2 * c] = ge.du2; sh_dn[r][jc][2 * c + 1] = ge.do2; }
            lf_du = __shfl_up_sync(0xffffffffu, ge.du2, QG_G);                   // lanes e==1: right edge of pixel (m,n-1), same layer
            lf_do = __shfl_up_sync(0xffffffffu, ge.do2, QG_G);
            __syncthreads();
        } else { up_du = gq.du2; up_do = gq.do2; }
    }

    // per-lane share of the folded edge gradients (:37-40): layer c of this lane
    float pm = ge.du1, ps = ge.do1;
    if (is_out) {
        if (e == 0) {
            if (QG_G4_HALO_WARP || r > 0) { pm += sh_dn[r - 1][jc][2 * c]; ps += sh_dn[r - 1][jc][2 * c + 1]; }
            else                          { pm += up_du; ps += up_do; }
        } else { pm += lf_du; ps += lf_do; }
    }
    const float E_mu = pm + __shfl_xor_sync(0xffffffffu, pm, 1);             // lanes (g, g^1) share the layer
    const float E_sig = ps + __shfl_xor_sync(0xffffffffu, ps, 1);
    float E_e = ge.Ei + __shfl_xor_sync(0xffffffffu, ge.Ei, 1);
    E_e += __shfl_xor_sync(0xffffffffu, E_e, 2);                              // :48 edge part, all four edges
    float E_a = ge.da + __shfl_xor_sync(0xffffffffu, ge.da, 1);
    E_a += __shfl_xor_sync(0xffffffffu, E_a, 2);                              // :36 edge part

    float red[QG_NRED] = {0.f, 0.f, 0.f, 0.f};
    float *o = out + (long long)l * pl + idx;
    float *d = DUMP ? p.dbg + (long long)l * pl + idx : nullptr;
    if (is_out) {
        if (DUMP) d[(5 + g) * fstr] = ge.dp;
        else o[(F_ROU0 + g) * fstr] = qg_clamp(fmaf(ge.dp, step, rouq), -p.corr_tor, p.corr_tor);     // :45
    }
    // ---- node term (:29, :87-116): XJ rows g, g+4, ... on this lane; all 32 lanes take part in the shuffles ---------
    float muu = 0.f, muv = 0.f, sigu = 1.f, sigv = 1.f, pn = 0.f;
    QgMoments mo = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    QgSpectral sp;
    if (is_out) {
        muu = __ldg(base + F_MUU * fstr + idx);  muv = __ldg(base + F_MUV * fstr + idx);
        sigu = __ldg(base + F_SIGU * fstr + idx); sigv = __ldg(base + F_SIGV * fstr + idx);
        pn = __ldg(base + F_PN * fstr + idx);
    }
    sp.set(pn);
    if (is_out) {
        if (SUPER) {
            float I1b[16];
            const float *ip = p.I1 + (long long)(4 * m) * p.pitchI + 4 * n;
#pragma unroll
            for (int di = 0; di < 4; ++di) {
                float4 v = __ldg(reinterpret_cast<const float4 *>(ip + (long long)di * p.pitchI));
                I1b[di * 4 + 0] = v.x; I1b[di * 4 + 1] = v.y; I1b[di * 4 + 2] = v.z; I1b[di * 4 + 3] = v.w;
            }
            const int lastx = p.No - 2, lasty = p.Mo - 2, m4 = 4 * m, n4 = 4 * n;
            mo = qg_quadrature_flat<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                return qg_super_sample(p.VV8, p.pitchV, m4, n4, lastx, lasty, x, I1b, p.epsn);
            }, g, QG_G);
        } else {
            const float I1v = __ldg(p.I1 + (long long)m * p.pitchI + n);
            const int lastx = p.No - 2, lasty = p.Mo - 2;
            QgTapCache tc;
            mo = qg_quadrature<KT>(p.tab, p.K, muu, muv, sigu, sigv, sp, -p.lambdad, [&](float2 x) {
                return qg_node_sample<(KT >= 7)>(p.VV8, p.pitchV, m, n, lastx, lasty, x, I1v, p.epsn, tc);
            }, g, QG_G);
        }
    }
#pragma unroll
    for (int sft = 1; sft < QG_G; sft <<= 1) {                                 // combine the four lanes' partial moments
        mo.E += __shfl_xor_sync(0xffffffffu, mo.E, sft);   mo.MI += __shfl_xor_sync(0xffffffffu, mo.MI, sft);
        mo.MJ += __shfl_xor_sync(0xffffffffu, mo.MJ, sft); mo.MII += __shfl_xor_sync(0xffffffffu, mo.MII, sft);
        mo.MJJ += __shfl_xor_sync(0xffffffffu, mo.MJJ, sft); mo.MB += __shfl_xor_sync(0xffffffffu, mo.MB, sft);
    }
    if (is_out) {
        const QgGrad gn = qg_epilogue(mo, sp, a, sigu, sigv, pn, -3.0f * T);
        // lanes of layer c hold E_mu / E_sig of that layer: g=0 finishes (mu_u, sig_u, pn, sums), g=2 finishes (mu_v, sig_v)
        const float G_mu = (c ? gn.du2 : gn.du1) + E_mu, G_sig = (c ? gn.do2 : gn.do1) + E_sig;
        if (g == 0) {
            const float e_px = gn.Ei + E_e, da_px = gn.da + E_a;
            red[0] = e_px; red[1] = da_px; red[2] = fabsf(G_mu); red[3] = fabsf(G_sig);
            if (DUMP) { d[0 * fstr] = G_mu; d[2 * fstr] = G_sig; d[4 * fstr] = gn.dp; d[9 * fstr] = e_px; d[10 * fstr] = da_px; }
            else {
                o[F_MUU * fstr] = qg_clamp(fmaf(G_mu, step, muu), p.minu, p.maxu);
                o[F_SIGU * fstr] = qg_clamp(fmaf(G_sig, step * p.sig_step, sigu), p.sig_min, p.sig_max);
                o[F_PN * fstr] = qg_clamp(fmaf(gn.dp, step, pn), -p.corr_tor, p.corr_tor);
            }
        } else if (g == 2) {
            if (DUMP) { d[1 * fstr] = G_mu; d[3 * fstr] = G_sig; }
            else {
                o[F_MUV * fstr] = qg_clamp(fmaf(G_mu, step, muv), p.minv, p.maxv);
                o[F_SIGV * fstr] = qg_clamp(fmaf(G_sig, step * p.sig_step, sigv), p.sig_min, p.sig_max);
            }
        }
    }
    qg_block_finish_w<DUMP, QgTileG4<KT, SUPER>::NW, QG_G4_W0>(p, ctrl, red, r, j);
}
