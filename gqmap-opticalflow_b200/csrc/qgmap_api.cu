// qgmap_api.cu -- C ABI of libqgmap.so (include/qgmap.h): handle management, layout conversion at the boundary,
// iteration launches (CUDA graph of identical kernel nodes), monitoring, the one-call solver.
#include "../../include/qgmap.h"
#include "qgmap_iter.cuh"
#include "qgmap_walk.cuh"
#include "qgmap_internal.h"
#include "qgmap_layout.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

static thread_local std::string g_last_error;
thread_local long long g_solve_launches = 0;
thread_local float g_solve_ms = 0.f;
void qgmap_set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }

// Band mode (multi-process): after the NCCL all-reduce of ctrl->sums one thread advances the control block.
__global__ void qgmap_advance_kernel(const __grid_constant__ QgIterParams p)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && !p.ctrl->stop) qg_advance(p, p.ctrl, p.ctrl->sums);
}

#define QG_FAIL(h, code, ...)                                                        \
    do {                                                                             \
        char _b[512];                                                                \
        snprintf(_b, sizeof _b, __VA_ARGS__);                                        \
        if (h) (h)->err = _b;                                                        \
        g_last_error = _b;                                                           \
        return (code);                                                               \
    } while (0)

#define QG_CUDA(h, expr)                                                             \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            QG_FAIL(h, QGMAP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *qgmap_last_error(const qgmap_handle *h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int qgmap_config_defaults(qgmap_config *c, int variant)
{
    if (!c || (variant != QGMAP_VARIANT_FULL && variant != QGMAP_VARIANT_SUPER)) return QGMAP_ERR_ARG;
    std::memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof *c;
    c->variant = variant;
    c->L = 1; c->K = 3;
    c->lambdad = 1.0; c->epsn = 1e-6;
    const bool sup = variant == QGMAP_VARIANT_SUPER;
    c->lambdas = sup ? 16.0 : 5.0;                 // optical_flowSuper.m:22 / optical_flow.m:19
    c->temperature = sup ? 0.2 : 0.0;              // optical_flowSuper.m:25 / optical_flow.m:22
    c->drate = sup ? 0.75 : 0.5;
    c->minu = c->minv = -1.0; c->maxu = c->maxv = 1.0;
    c->sigma_min = 0.01; c->sigma_max = sup ? 25.0 : 23.0;
    c->corr_tor = 1.0 - 1e-5;
    c->step0 = sup ? 0.001 : 0.1; c->step_tau = sup ? 4000.0 : 8000.0;
    c->alpha_scale = 1e-7; c->T_floor = 0.001; c->tor = 1e-4; c->sigma_step_scale = 1.0;
    c->alpha_start = 500; c->alpha_mode = QGMAP_ALPHA_SOFTMAX;
    c->anneal_every = sup ? 500 : 0;
    c->device = -1; c->row_begin = 0; c->row_end = 0; c->log_every = 300;
    return QGMAP_OK;
}

// --------------------------------------------------------------------------------------------------------------------
template <int KT, bool SUPER, bool DUMP>
static void launch_inst(const qgmap_handle *h) {
    if (!SUPER && h->walk) { qgmap_walk_kernel<KT, DUMP><<<h->grid, 32, 0, h->stream>>>(h->params); return; }
    const dim3 block(QG_TW, h->lanes_per_belief == 4 ? QgTileG4<KT, SUPER>::NW : QgTile<KT, SUPER>::TH + QgTile<KT, SUPER>::W0);
    if (h->pdl && !DUMP && h->nranks <= 1) {             // programmatic dependent launch: the next iteration's launch overlaps this one's tail (qg_pdl_enter)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = h->grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (h->lanes_per_belief == 4) cudaLaunchKernelEx(&cfg, qgmap_iter_kernel_g4<KT, SUPER, DUMP>, h->params);
        else cudaLaunchKernelEx(&cfg, qgmap_iter_kernel<KT, SUPER, DUMP>, h->params);
        return;
    }
    if (h->lanes_per_belief == 4) qgmap_iter_kernel_g4<KT, SUPER, DUMP><<<h->grid, block, 0, h->stream>>>(h->params);
    else qgmap_iter_kernel<KT, SUPER, DUMP><<<h->grid, block, 0, h->stream>>>(h->params);
}
template <bool SUPER, bool DUMP>
static void launch_k(const qgmap_handle *h) {
    switch (h->K) {
        case 3: launch_inst<3, SUPER, DUMP>(h); break;
        case 5: launch_inst<5, SUPER, DUMP>(h); break;
        case 7: launch_inst<7, SUPER, DUMP>(h); break;
        case 9: launch_inst<9, SUPER, DUMP>(h); break;
        case 11: launch_inst<11, SUPER, DUMP>(h); break;
        default: launch_inst<0, SUPER, DUMP>(h); break;
    }
}
void qgmap_launch_iteration(const qgmap_handle *h);
static void launch_iter(const qgmap_handle *h, bool dump) {
    const bool sup = h->cfg.variant == QGMAP_VARIANT_SUPER;
    if (sup) { if (dump) launch_k<true, true>(h); else launch_k<true, false>(h); }
    else     { if (dump) launch_k<false, true>(h); else launch_k<false, false>(h); }
}

static int device_check(qgmap_handle *h, int device, int *dev_out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        QG_FAIL(h, QGMAP_ERR_CUDA, "no CUDA device available (%s); libqgmap has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = device;
    if (dev < 0) QG_CUDA(h, cudaGetDevice(&dev));
    if (dev >= ndev) QG_FAIL(h, QGMAP_ERR_ARG, "device %d out of range (%d devices)", dev, ndev);
    QG_CUDA(h, cudaSetDevice(dev));
    *dev_out = dev;
    return QGMAP_OK;
}

static void free_handle(qgmap_handle *h)
{
    if (!h) return;
    if (h->device >= 0) cudaSetDevice(h->device);
    qgmap_band_release(h);
    qgmap_p2p_release(h);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evb0) cudaEventDestroy(h->evb0);
    if (h->evb1) cudaEventDestroy(h->evb1);
    void *ptrs[] = {h->I1f, h->VVf, h->VVh, h->I1d, h->VVd, h->buf[0], h->buf[1], h->dbg, h->ctrl, h->partials, h->gpartials, h->tickets, h->hist[0],
                    h->hist[1], h->hist[2], h->stage, h->mon_partials, h->d_map, h->d_tflow, h->d_unknown};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->ctrl_host) cudaFreeHost(h->ctrl_host);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int ensure_stage(qgmap_handle *h, size_t doubles)
{
    if (doubles <= h->stage_cap) return QGMAP_OK;
    if (h->stage) { QG_CUDA(h, cudaFree(h->stage)); h->stage = nullptr; h->stage_cap = 0; }
    QG_CUDA(h, cudaMalloc(&h->stage, doubles * sizeof(double)));
    h->stage_cap = doubles;
    return QGMAP_OK;
}

static int ensure_hist(qgmap_handle *h, int its)
{
    if (its <= h->hist_cap) return QGMAP_OK;
    int cap = (int)std::min<long long>(std::max<long long>({(long long)its, 2LL * h->hist_cap, 1024LL}), 1LL << 30);
    for (int k = 0; k < 3; ++k) {
        double *n = nullptr;
        QG_CUDA(h, cudaMalloc(&n, (size_t)cap * sizeof(double)));
        QG_CUDA(h, cudaMemsetAsync(n, 0, (size_t)cap * sizeof(double), h->stream));
        if (h->hist[k]) {
            QG_CUDA(h, cudaMemcpyAsync(n, h->hist[k], (size_t)h->hist_cap * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            QG_CUDA(h, cudaStreamSynchronize(h->stream));
            QG_CUDA(h, cudaFree(h->hist[k]));
        }
        h->hist[k] = n;
    }
    h->hist_cap = cap;
    h->params.hist_energy = h->hist[0]; h->params.hist_dmu = h->hist[1]; h->params.hist_dsig = h->hist[2];
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }     // params are baked into graph nodes
    return QGMAP_OK;
}

// Rows one warp of the row-walking kernel walks.  The halo edge above a strip costs about 0.15 of a row, so long strips are
// cheap per row; but a launch is `strips / (SMs x resident warps)` waves and its last wave runs at whatever occupancy is left.
// Rule: long strips (8 rows) when that still leaves >= 4 waves; otherwise the shortest strip that fits the whole launch into
// ONE wave (every warp slot walks one strip, nothing is quantised), capped at 16 rows; QGMAP_STRIP_ROWS / cfg.strip_rows override.
int qgmap_pick_strip_rows(int strips_x, int out_rows, int L, int sms, int K, int cfg_rows)
{
    if (const char *env = getenv("QGMAP_STRIP_ROWS")) { const int r = atoi(env); if (r >= 1) return std::min(r, 1 << 20); }
    if (cfg_rows >= 1) return cfg_rows;
    const int resident = (K == 3 || K == 5) ? QG_WALK_MINB : QG_WALK_MINB_BIGK;
    const long long slots = (long long)sms * resident;
    const long long rows = (long long)strips_x * out_rows * L;              // warp-rows of one launch
    if (rows >= 4 * 8 * slots) return 8;
    for (int r = 1; r <= 16; ++r)
        if ((long long)strips_x * ((out_rows + r - 1) / r) * L <= slots) return r;
    return 8;
}

extern "C" int qgmap_create(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, qgmap_handle **out)
{
    return qg_guard([&]() -> int {
    if (!cfg || !I1 || !I2 || !out) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "qgmap_create: NULL argument");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(qgmap_config))
        QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "qgmap_config.struct_size %d != %zu (ABI mismatch)", cfg->struct_size, sizeof(qgmap_config));
    const bool sup = cfg->variant == QGMAP_VARIANT_SUPER;
    if (cfg->variant != QGMAP_VARIANT_FULL && !sup) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "bad variant %d", cfg->variant);
    if (cfg->L < 1 || cfg->L > QGMAP_LMAX) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "L=%d out of range 1..%d", cfg->L, QGMAP_LMAX);
    if (cfg->K < 1 || cfg->K > QGMAP_KMAX) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "K=%d out of range 1..%d", cfg->K, QGMAP_KMAX);
    if (Mo < 4 || No < 4) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "image %dx%d too small (bicubic needs >=4x4)", Mo, No);
    if ((long long)(Mo + 2) * ((No + 3) / 4 * 4 + 4) >= (1LL << 31))
        QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "image %dx%d too large (gather entries are indexed with 32 bits)", Mo, No);
    if (sup && (Mo % 4 || No % 4)) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "super-pixel variant needs Mo,No divisible by 4 (got %dx%d)", Mo, No);
    const int M = sup ? Mo / 4 : Mo, N = sup ? No / 4 : No;
    if (M < 3 || N < 3) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "belief grid %dx%d has no interior", M, N);
    int rb = cfg->row_begin, re = cfg->row_end;
    if (rb == 0 && re == 0) re = M;
    if (rb < 0 || re > M || rb >= re) QG_FAIL((qgmap_handle *)nullptr, QGMAP_ERR_ARG, "bad row band [%d,%d) of %d", rb, re, M);

    qgmap_handle *h = new qgmap_handle();
    h->cfg = *cfg;
    h->Mo = Mo; h->No = No; h->M = M; h->N = N; h->L = cfg->L; h->K = cfg->K;
    int rc = device_check(h, cfg->device, &h->device);
    if (rc) { g_last_error = h->err; delete h; return rc; }
#define QG_CUDA_C(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
        char _b[512]; snprintf(_b, sizeof _b, "%s failed: %s", #expr, cudaGetErrorString(_e)); g_last_error = _b; free_handle(h); return QGMAP_ERR_CUDA; } } while (0)
    QG_CUDA_C(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    QG_CUDA_C(cudaEventCreate(&h->ev0));
    QG_CUDA_C(cudaEventCreate(&h->ev1));

    // rows stored locally: owned band plus one halo row on each side (clipped to the grid)
    h->row_begin = rb; h->row_end = re;
    h->g0 = std::max(rb - 1, 0); h->g1 = std::min(re + 1, M);
    h->rows_local = h->g1 - h->g0;
    h->out_r0 = std::max(rb, 1); h->out_r1 = std::min(re, M - 1);
    h->P = (N + 31) / 32 * 32;
    h->plane = (long long)h->rows_local * h->P;
    const size_t state_floats = (size_t)F_COUNT * h->L * h->plane;
    QG_CUDA_C(cudaMalloc(&h->buf[0], state_floats * sizeof(float)));
    QG_CUDA_C(cudaMalloc(&h->buf[1], state_floats * sizeof(float)));
    QG_CUDA_C(cudaMemsetAsync(h->buf[0], 0, state_floats * sizeof(float), h->stream));
    QG_CUDA_C(cudaMemsetAsync(h->buf[1], 0, state_floats * sizeof(float), h->stream));

    // images: fp64 row-major copies for monitoring, fp32 for the iteration kernel
    h->pitchI = (No + 3) / 4 * 4; h->pitchV = (No + 2 + 3) / 4 * 4;
    const size_t nI = (size_t)Mo * h->pitchI, nV = (size_t)(Mo + 2) * h->pitchV;
    QG_CUDA_C(cudaMalloc(&h->I1d, nI * sizeof(double)));
    QG_CUDA_C(cudaMalloc(&h->VVd, nV * sizeof(double)));
    QG_CUDA_C(cudaMalloc(&h->I1f, nI * sizeof(float)));
    h->pitch4 = (No + 3) / 4 * 4;
    QG_CUDA_C(cudaMalloc(&h->VVf, (size_t)(Mo + 2) * h->pitch4 * sizeof(QgTap8)));       // entry rows 0..Mo+1
    QG_CUDA_C(cudaMemsetAsync(h->I1d, 0, nI * sizeof(double), h->stream));
    QG_CUDA_C(cudaMemsetAsync(h->VVd, 0, nV * sizeof(double), h->stream));
    {
        const size_t MoNo = (size_t)Mo * No;
        QG_CUDA_C(cudaMalloc(&h->stage, MoNo * sizeof(double)));
        h->stage_cap = MoNo;
        dim3 tb(32, 8), tg((No + 31) / 32, (Mo + 31) / 32, 1);
        QG_CUDA_C(cudaMemcpyAsync(h->stage, I1, MoNo * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        qgmap_import_kernel<double><<<tg, tb, 0, h->stream>>>(h->stage, Mo, No, 1, h->I1d, h->pitchI, 0, 0, Mo, 0);
        QG_CUDA_C(cudaStreamSynchronize(h->stream));            // stage is reused for I2
        QG_CUDA_C(cudaMemcpyAsync(h->stage, I2, MoNo * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        qgmap_import_kernel<double><<<tg, tb, 0, h->stream>>>(h->stage, Mo, No, 1, h->VVd + h->pitchV + 1, h->pitchV, 0, 0, Mo, 0);
        qgmap_vv_rows_kernel<<<(No + 2 + 127) / 128, 128, 0, h->stream>>>(Mo, No, h->VVd, h->pitchV);
        qgmap_vv_cols_kernel<<<(Mo + 2 + 127) / 128, 128, 0, h->stream>>>(Mo, No, h->VVd, h->pitchV);
        qgmap_cast_kernel<float><<<256, 256, 0, h->stream>>>(h->I1d, h->I1f, (long long)nI);
        qgmap_pack8_kernel<<<dim3((h->pitch4 + 127) / 128, Mo + 2), 128, 0, h->stream>>>(h->VVd, h->pitchV, Mo + 2, No + 2,
                                                                                      reinterpret_cast<float4 *>(h->VVf), h->pitch4, Mo + 2);
        QG_CUDA_C(cudaGetLastError());
        // fp16 4x4-block gather layout for wide beliefs -- kept only if the padded frame is exactly representable (integer grey
        // levels, as the reference's double(rgb2gray(...)) frames are); QGMAP_TAPS=f32 disables it (A/B measurements)
        const char *tenv = getenv("QGMAP_TAPS");
        if (!sup && !(tenv && !strcmp(tenv, "f32"))) {
            int *d_flag = nullptr, flag = 0;
            QG_CUDA_C(cudaMalloc(&h->VVh, (size_t)(Mo + 2) * h->pitch4 * sizeof(QgTap16h)));
            QG_CUDA_C(cudaMalloc(&d_flag, sizeof(int)));
            QG_CUDA_C(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
            qgmap_pack16h_kernel<<<dim3((h->pitch4 + 127) / 128, Mo + 2), 128, 0, h->stream>>>(h->VVd, h->pitchV, Mo + 2, No + 2,
                                                                                            reinterpret_cast<uint4 *>(h->VVh), h->pitch4, Mo + 2, d_flag);
            QG_CUDA_C(cudaGetLastError());
            QG_CUDA_C(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            QG_CUDA_C(cudaStreamSynchronize(h->stream));
            cudaFree(d_flag);
            if (flag) { cudaFree(h->VVh); h->VVh = nullptr; }
        }
    }

    // tiles.  One thread per (belief pixel, component) when that fills the GPU; four lanes per belief (edge quadratures and
    // node rows split over the lanes, combined by warp shuffles) when the belief grid is small.
    const int out_rows = std::max(h->out_r1 - h->out_r0, 0);
    {
        const char *env = getenv("QGMAP_LANES");
        const long long beliefs = (long long)(N - 2) * std::max(out_rows, 1) * h->L;
        // measured on B200 (profiles/): the 4-lane mapping pays for the super-pixel variant on small grids (1.8x at 120x160 beliefs,
        // K=5); full-resolution it costs ~40% at K=3 where the replicated per-lane prologue/epilogue outweighs the shorter loop
        (void)beliefs;
        h->lanes_per_belief = env ? atoi(env) : (sup ? 4 : 1);      // super-pixel: 4 lanes win at every size measured (480x640 .. 4K)
        if (h->lanes_per_belief != 4) h->lanes_per_belief = 1;
    }
    const int tw = h->lanes_per_belief == 4 ? QG_CW - 1 : QG_TW - 1;
    const int th = qg_tile_rows(qg_template_k(h->K), sup);
    h->grid = dim3((N - 2 + tw - 1) / tw, std::max((out_rows + th - 1) / th, 1), h->L);
    // QGMAP_ITER=walk selects the row-walking form of the full-resolution kernel (qgmap_walk.cuh: one warp per strip of 31 columns
    // x strip_rows rows, no CTA barrier; ~15% fewer instructions but it needs 96+ registers, measured 3-7% behind the tiled kernel
    // on B200, profiles/r02_walk_ab.txt); default: the tiled kernel of qgmap_iter.cuh.
    int strip_rows = 0;
    {
        const char *env = getenv("QGMAP_ITER");
        h->walk = !sup && h->lanes_per_belief == 1 && env && !strcmp(env, "walk");
        const char *penv = getenv("QGMAP_PDL");
        h->pdl = penv && atoi(penv) != 0;
        if (h->walk) {
            int ndev_sm = 148;
            cudaDeviceGetAttribute(&ndev_sm, cudaDevAttrMultiProcessorCount, h->device);
            strip_rows = qgmap_pick_strip_rows(h->grid.x, std::max(out_rows, 1), h->L, ndev_sm, h->K, cfg->strip_rows);
            h->grid = dim3(h->grid.x, std::max((out_rows + strip_rows - 1) / strip_rows, 1), h->L);
            const size_t ngroups = (size_t)h->grid.y * h->grid.z;
            QG_CUDA_C(cudaMalloc(&h->gpartials, ngroups * QG_NRED * sizeof(double)));
            QG_CUDA_C(cudaMemsetAsync(h->gpartials, 0, ngroups * QG_NRED * sizeof(double), h->stream));
            QG_CUDA_C(cudaMalloc(&h->tickets, ngroups * sizeof(unsigned int)));
            QG_CUDA_C(cudaMemsetAsync(h->tickets, 0, ngroups * sizeof(unsigned int), h->stream));
        }
    }
    const size_t nblk = (size_t)h->grid.x * h->grid.y * h->grid.z;
    QG_CUDA_C(cudaMalloc(&h->partials, nblk * QG_NRED * sizeof(double)));
    QG_CUDA_C(cudaMemsetAsync(h->partials, 0, nblk * QG_NRED * sizeof(double), h->stream));
    QG_CUDA_C(cudaMalloc(&h->ctrl, sizeof(QgCtrl)));
    QG_CUDA_C(cudaMemsetAsync(h->ctrl, 0, sizeof(QgCtrl), h->stream));
    QG_CUDA_C(cudaMallocHost(&h->ctrl_host, sizeof(QgCtrl)));
    QG_CUDA_C(cudaMalloc(&h->mon_partials, 1024 * sizeof(double)));

    // Gauss-Hermite tables (GaussHermite_2.m:21-32; gqmap_gpu_mixture.m:8-10)
    double X[QGMAP_KMAX], W[QGMAP_KMAX];
    if (qgmap_gauss_hermite(h->K, X, W) != QGMAP_OK) { g_last_error = "gauss_hermite failed"; free_handle(h); return QGMAP_ERR_ARG; }
    QgIterParams &p = h->params;
    std::memset(&p, 0, sizeof p);
    for (int k = 0; k < h->K; ++k) {
        p.tab.X[k] = (float)X[k]; p.tab.W[k] = (float)W[k];
        p.tab.WX[k] = (float)(W[k] * X[k]); p.tab.WXX[k] = (float)(W[k] * X[k] * X[k]);
    }
    p.I1 = h->I1f; p.pitchI = h->pitchI; p.VV8 = h->VVf; p.VVh = h->VVh; p.pitchV = h->pitch4;
    p.buf[0] = h->buf[0]; p.buf[1] = h->buf[1];
    p.plane = h->plane; p.P = h->P; p.M = M; p.N = N; p.L = h->L; p.Mo = Mo; p.No = No;
    p.g0 = h->g0; p.out_r0 = h->out_r0; p.out_r1 = h->out_r1; p.K = h->K; p.band = 0;
    p.lambdad = (float)cfg->lambdad; p.lambdas = (float)cfg->lambdas; p.epsn = (float)cfg->epsn;
    p.minu = (float)cfg->minu; p.maxu = (float)cfg->maxu; p.minv = (float)cfg->minv; p.maxv = (float)cfg->maxv;
    p.sig_min = (float)cfg->sigma_min; p.sig_max = (float)cfg->sigma_max; p.corr_tor = (float)cfg->corr_tor;
    p.step0 = cfg->step0; p.step_tau = cfg->step_tau; p.alpha_scale = cfg->alpha_scale; p.drate = cfg->drate;
    p.sig_step = (float)cfg->sigma_step_scale; p.T_floor = cfg->T_floor; p.tor = cfg->tor; p.alpha_start = cfg->alpha_start; p.alpha_mode = cfg->alpha_mode;
    p.anneal_every = cfg->anneal_every;
    { const char *wenv = getenv("QGMAP_WIDE_REACH"); p.wide_reach = wenv ? (float)atof(wenv) : 0.5f; }
    p.ctrl = h->ctrl; p.partials = h->partials; p.gpartials = h->gpartials; p.tickets = h->tickets; p.strip_rows = strip_rows;
    QG_CUDA_C(cudaStreamSynchronize(h->stream));
    if (ensure_hist(h, 1024) != QGMAP_OK) { free_handle(h); return QGMAP_ERR_CUDA; }
    QG_CUDA_C(cudaStreamSynchronize(h->stream));
#undef QG_CUDA_C
    *out = h;
    return QGMAP_OK;
    });
}

extern "C" int qgmap_destroy(qgmap_handle *h)
{
    if (!h) return QGMAP_ERR_ARG;
    free_handle(h);
    return QGMAP_OK;
}

extern "C" int qgmap_dims(const qgmap_handle *h, int *M, int *N, int *L)
{
    if (!h) return QGMAP_ERR_ARG;
    if (M) *M = h->M;
    if (N) *N = h->N;
    if (L) *L = h->L;
    return QGMAP_OK;
}

extern "C" int qgmap_image_dims(const qgmap_handle *h, int *Mo, int *No)
{
    if (!h) return QGMAP_ERR_ARG;
    if (Mo) *Mo = h->Mo;
    if (No) *No = h->No;
    return QGMAP_OK;
}

// column-major fp64 host array (M x N x planes) -> fp32 planes [field_first .. ) of BOTH ping-pong buffers
template <typename S>      // S = host element type: double (MATLAB double arrays) or float (single)
static int import_planes(qgmap_handle *h, const S *src, int planes, int plane_first)
{
    const size_t n = (size_t)h->M * h->N * planes;
    int rc = ensure_stage(h, n);
    if (rc) return rc;
    S *stage = reinterpret_cast<S *>(h->stage);
    if (h->g0 == 0 && h->g1 == h->M)
        QG_CUDA(h, cudaMemcpyAsync(stage, src, n * sizeof(S), cudaMemcpyHostToDevice, h->stream));
    else      // a band needs only its stored rows [g0,g1) of every column: strided copy (column-major source)
        QG_CUDA(h, cudaMemcpy2DAsync(stage + h->g0, (size_t)h->M * sizeof(S), src + h->g0, (size_t)h->M * sizeof(S),
                                     (size_t)(h->g1 - h->g0) * sizeof(S), (size_t)h->N * planes, cudaMemcpyHostToDevice, h->stream));
    dim3 tb(32, 8), tg((h->N + 31) / 32, (h->rows_local + 31) / 32, planes);
    for (int b = 0; b < 2; ++b)
        qgmap_import_kernel<float, S><<<tg, tb, 0, h->stream>>>(stage, h->M, h->N, planes,
                                                                 h->buf[b] + (size_t)plane_first * h->plane, h->P, h->plane,
                                                                 h->g0, h->g1, h->g0);
    QG_CUDA(h, cudaGetLastError());
    QG_CUDA(h, cudaStreamSynchronize(h->stream));     // stage reused by the next call
    return QGMAP_OK;
}

// fp32 planes of the CURRENT buffer -> column-major host array (fp64 or fp32); only the owned rows [row_begin,row_end) are written
template <typename D>
static int export_planes(qgmap_handle *h, int cur, int planes, int plane_first, D *dst)
{
    const size_t n = (size_t)h->M * h->N * planes;
    int rc = ensure_stage(h, n);
    if (rc) return rc;
    D *stage = reinterpret_cast<D *>(h->stage);
    const bool whole = (h->row_begin == 0 && h->row_end == h->M);
    dim3 tb(32, 8), tg((h->N + 31) / 32, (h->row_end - h->row_begin + 31) / 32, planes);
    qgmap_export_kernel<float, D><<<tg, tb, 0, h->stream>>>(h->buf[cur] + (size_t)plane_first * h->plane, h->P, h->plane, h->g0,
                                                             h->row_begin, h->row_end, stage, h->M, h->N);
    QG_CUDA(h, cudaGetLastError());
    if (whole)
        QG_CUDA(h, cudaMemcpyAsync(dst, stage, n * sizeof(D), cudaMemcpyDeviceToHost, h->stream));
    else      // a band returns only the rows it owns: strided copy into the caller's full-grid array, other rows untouched
        QG_CUDA(h, cudaMemcpy2DAsync(dst + h->row_begin, (size_t)h->M * sizeof(D), stage + h->row_begin, (size_t)h->M * sizeof(D),
                                     (size_t)(h->row_end - h->row_begin) * sizeof(D), (size_t)h->N * planes, cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return QGMAP_OK;
}

static int sync_ctrl(qgmap_handle *h)
{
    QG_CUDA(h, cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QgCtrl), cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return QGMAP_OK;
}

template <typename S>
static int set_state_t(qgmap_handle *h, const S *muu, const S *muv, const S *sigu, const S *sigv, const S *pn, const S *rou,
                       const double *w, const double *alpha, double T, int it)
{
    if (!h) return QGMAP_ERR_ARG;
    if (!muu || !muv || !sigu || !sigv || !pn || !rou || !w) QG_FAIL(h, QGMAP_ERR_ARG, "qgmap_set_state: NULL array");
    if (it < 1) QG_FAIL(h, QGMAP_ERR_ARG, "qgmap_set_state: it must be >= 1");
    QG_CUDA(h, cudaSetDevice(h->device));
    const int L = h->L;
    const S *fields[5] = {muu, muv, sigu, sigv, pn};
    int rc;
    for (int f = 0; f < 5; ++f)
        if ((rc = import_planes(h, fields[f], L, f * L)) != QGMAP_OK) return rc;
    for (int q = 0; q < 4; ++q)     // rou(:,:,:,e,c): q = e + 2c is the slowest MATLAB dimension pair
        if ((rc = import_planes(h, rou + (size_t)q * h->M * h->N * L, L, (F_ROU0 + q) * L)) != QGMAP_OK) return rc;
    QgCtrl c;
    std::memset(&c, 0, sizeof c);
    // it odd -> current buffer 0; both buffers hold the same state after import, so any `it` is consistent
    c.it = it; c.stop = 0; c.its = std::numeric_limits<int>::max(); c.ticket = 0; c.T = T;
    double se = 0.0;
    for (int l = 0; l < L; ++l) { c.w[l] = w[l]; se += std::exp(w[l]); }
    for (int l = 0; l < L; ++l) c.alpha[l] = alpha ? alpha[l] : std::exp(w[l]) / se;      // :18
    *h->ctrl_host = c;
    QG_CUDA(h, cudaMemcpyAsync(h->ctrl, h->ctrl_host, sizeof c, cudaMemcpyHostToDevice, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->has_state = true;
    return qgmap_band_refresh(h);
}

template <typename D>
static int get_state_t(qgmap_handle *h, D *muu, D *muv, D *sigu, D *sigv, D *pn, D *rou, double *w, double *alpha, double *T, int *it)
{
    if (!h) return QGMAP_ERR_ARG;
    if (!h->has_state) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_get_state: no state set");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = sync_ctrl(h);
    if (rc) return rc;
    const QgCtrl &c = *h->ctrl_host;
    const int cur = (c.it - 1) & 1, L = h->L;
    D *fields[5] = {muu, muv, sigu, sigv, pn};
    for (int f = 0; f < 5; ++f)
        if (fields[f] && (rc = export_planes(h, cur, L, f * L, fields[f])) != QGMAP_OK) return rc;
    if (rou)
        for (int q = 0; q < 4; ++q)
            if ((rc = export_planes(h, cur, L, (F_ROU0 + q) * L, rou + (size_t)q * h->M * h->N * L)) != QGMAP_OK) return rc;
    for (int l = 0; l < L; ++l) { if (w) w[l] = c.w[l]; if (alpha) alpha[l] = c.alpha[l]; }
    if (T) *T = c.T;
    if (it) *it = c.it;
    return QGMAP_OK;
}

extern "C" int qgmap_set_state(qgmap_handle *h, const double *muu, const double *muv, const double *sigu,
                               const double *sigv, const double *pn, const double *rou, const double *w,
                               const double *alpha, double T, int it)
{
    return set_state_t<double>(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it);
}
extern "C" int qgmap_get_state(qgmap_handle *h, double *muu, double *muv, double *sigu, double *sigv, double *pn,
                               double *rou, double *w, double *alpha, double *T, int *it)
{
    return get_state_t<double>(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it);
}
// The same with single-precision host arrays (MATLAB `single`): the device keeps the beliefs in fp32, so this is the lossless
// boundary format at half the host<->device bytes (4K, L=3: 0.9 GB instead of 1.8 GB per direction).
extern "C" int qgmap_set_state_f32(qgmap_handle *h, const float *muu, const float *muv, const float *sigu, const float *sigv,
                                   const float *pn, const float *rou, const double *w, const double *alpha, double T, int it)
{
    return qg_guard([&]() -> int { return set_state_t<float>(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it); });
}
extern "C" int qgmap_get_state_f32(qgmap_handle *h, float *muu, float *muv, float *sigu, float *sigv, float *pn, float *rou,
                                   double *w, double *alpha, double *T, int *it)
{
    return qg_guard([&]() -> int { return get_state_t<float>(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it); });
}

// splitmix64 -> xoshiro256**
struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed) { for (auto &v : s) { seed += 0x9E3779B97F4A7C15ULL; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; v = z ^ (z >> 31); } }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// gqmap_gpu_mixture.m:18-24 with the library's own generator (MATLAB's rand stream is not reproducible outside MATLAB)
void qgmap_random_state(const qgmap_config &c, size_t n, int L, uint64_t seed, std::vector<double> &w, std::vector<double> &muu,
                        std::vector<double> &muv, std::vector<double> &sigu, std::vector<double> &sigv)
{
    w.resize(L); muu.resize(n); muv.resize(n); sigu.resize(n); sigv.resize(n);
    Rng r(seed);
    for (auto &v : w) v = r.uniform();                                           // :18
    for (auto &v : muu) v = c.minu + r.uniform() * (c.maxu - c.minu);            // :19
    for (auto &v : muv) v = c.minv + r.uniform() * (c.maxv - c.minv);            // :20
    for (auto &v : sigu) v = r.uniform() + (c.maxu - c.minu);                    // :21
    for (auto &v : sigv) v = r.uniform() + (c.maxv - c.minv);                    // :22
}

extern "C" int qgmap_init_state(qgmap_handle *h, uint64_t seed)
{
    return qg_guard([&]() -> int {
    if (!h) return QGMAP_ERR_ARG;
    const size_t n = (size_t)h->M * h->N * h->L;
    const qgmap_config &c = h->cfg;
    std::vector<double> w, muu, muv, sigu, sigv, pn(n, 0.0), rou(n * 4, 0.0);
    qgmap_random_state(c, n, h->L, seed, w, muu, muv, sigu, sigv);
    return qgmap_set_state(h, muu.data(), muv.data(), sigu.data(), sigv.data(), pn.data(), rou.data(), w.data(), nullptr,
                           c.temperature, 1);
    });
}

static const int kGraphLen = 25;

static int build_graph(qgmap_handle *h)
{
    if (h->graph || (h->nranks > 1 && !qgmap_p2p_fused(h))) return QGMAP_OK;   // NCCL / publish-kernel bands: stream launches
    cudaGraph_t g = nullptr;
    QG_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < kGraphLen; ++i) launch_iter(h, false);
    QG_CUDA(h, cudaStreamEndCapture(h->stream, &g));
    cudaError_t e = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) QG_FAIL(h, QGMAP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    return QGMAP_OK;
}

// common prologue of a step: history capacity, options.its, iteration counter at the start
int qgmap_prepare_step(qgmap_handle *h, int n, int its)
{
    if (!h) return QGMAP_ERR_ARG;
    if (n < 0 || its < 1) QG_FAIL(h, QGMAP_ERR_ARG, "qgmap_step: n=%d its=%d", n, its);
    if (!h->has_state) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_step before qgmap_set_state/qgmap_init_state");
    if (h->pending) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_step_begin: previous step not ended");
    QG_CUDA(h, cudaSetDevice(h->device));
    // history is indexed by iteration: room for what can actually run now (its may be "unbounded").  A state resumed PAST its
    // (set_state(it > its)) still executes one iteration before the reference's `it > its` test fires (:74-75), so the bound
    // is max(its, it0), not its.
    const long long it0 = h->ctrl_host->it;
    int rc = ensure_hist(h, (int)std::min<long long>(std::max<long long>((long long)its, it0), it0 - 1 + std::max(n, 1)));
    if (rc) return rc;
    h->it0 = h->ctrl_host->it;
    h->band_steps_enqueued = 0;
    if (h->ctrl_host->its != its) {
        h->ctrl_host->its = its;
        // the reference tests `it > its` after incrementing: a run resumed past its stops after one more iteration
        QG_CUDA(h, cudaMemcpyAsync(&h->ctrl->its, &h->ctrl_host->its, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    return QGMAP_OK;
}

// n iterations of a single domain or of a peer-memory band on the handle's stream: CUDA graphs of kGraphLen identical kernel nodes
// where one launch is one iteration (single domain; band with the exchange inside the row-walking kernel), stream launches else
int qgmap_enqueue_iterations(qgmap_handle *h, int n, long long *launches)
{
    int left = n, rc;
    if (h->nranks > 1 && !qgmap_p2p_fused(h)) {
        for (; left > 0; --left) { if ((rc = qgmap_p2p_iteration(h, launches)) != QGMAP_OK) return rc; ++*launches; }
        return QGMAP_OK;
    }
    if (left >= kGraphLen && (rc = build_graph(h)) != QGMAP_OK) return rc;
    for (; left >= kGraphLen; left -= kGraphLen) { QG_CUDA(h, cudaGraphLaunch(h->graph, h->stream)); *launches += kGraphLen; }
    for (; left > 0; --left) { launch_iter(h, false); ++*launches; }
    return QGMAP_OK;
}

// enqueue up to n iterations on the handle's stream; returns without waiting
extern "C" int qgmap_step_begin(qgmap_handle *h, int n, int its)
{
    if (h && h->in_group) QG_FAIL(h, QGMAP_ERR_STATE, "handle belongs to a qgmap_group: use qgmap_group_step");
    int rc = qgmap_prepare_step(h, n, its);
    if (rc) return rc;
    long long launches = 0;
    if ((h->nranks <= 1 || qgmap_p2p_fused(h)) && n >= kGraphLen && (rc = build_graph(h)) != QGMAP_OK) return rc;   // host-side, before the timed region
    QG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if (h->nranks > 1 && !h->p2p) {                      // row band over NCCL (qgmap_band.cu)
        for (int left = n; left > 0; --left) { if ((rc = qgmap_band_iteration(h, &launches)) != QGMAP_OK) return rc; ++launches; }
    } else {
        if (h->nranks > 1) qgmap_p2p_begin_step(h);      // row band, exchange by our own kernels over peer memory (qgmap_p2p.cu)
        if ((rc = qgmap_enqueue_iterations(h, n, &launches)) != QGMAP_OK) return rc;
    }
    QG_CUDA(h, cudaGetLastError());
    QG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    QG_CUDA(h, cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QgCtrl), cudaMemcpyDeviceToHost, h->stream));
    h->last_launches = launches;
    h->pending = true;
    return QGMAP_OK;
}

// wait for the iterations enqueued by qgmap_step_begin and fetch their per-iteration history
extern "C" int qgmap_step_end(qgmap_handle *h, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped)
{
    if (!h) return QGMAP_ERR_ARG;
    if (!h->pending) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_step_end without qgmap_step_begin");
    QG_CUDA(h, cudaSetDevice(h->device));
    h->pending = false;
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    cudaGetLastError();
    if (h->ctrl_host->comm_error)
        QG_FAIL(h, QGMAP_ERR_COMM, "band exchange timed out at iteration %d: a peer rank did not publish its boundary rows", h->ctrl_host->it);
    const int done = h->ctrl_host->it - h->it0;
    if (n_done) *n_done = done;
    if (stopped) *stopped = h->ctrl_host->stop;
    double *dst[3] = {energy, ptdmu, ptdsigma};
    bool any = false;
    for (int k = 0; k < 3; ++k)
        if (dst[k] && done > 0) {
            QG_CUDA(h, cudaMemcpyAsync(dst[k], h->hist[k] + (h->it0 - 1), (size_t)done * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            any = true;
        }
    if (any) QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return QGMAP_OK;
}

// epilogue of a step whose iterations were enqueued by a group: fetch the control block, then as qgmap_step_end
int qgmap_finish_step(qgmap_handle *h, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped)
{
    QG_CUDA(h, cudaSetDevice(h->device));
    QG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    QG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    QG_CUDA(h, cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QgCtrl), cudaMemcpyDeviceToHost, h->stream));
    h->pending = true;
    return qgmap_step_end(h, energy, ptdmu, ptdsigma, n_done, stopped);
}

extern "C" int qgmap_step(qgmap_handle *h, int n, int its, double *energy, double *ptdmu, double *ptdsigma,
                          int *n_done, int *stopped)
{
    int rc = qgmap_step_begin(h, n, its);
    if (rc) return rc;
    return qgmap_step_end(h, energy, ptdmu, ptdsigma, n_done, stopped);
}

// Independent frame pairs (one handle each, all on the same device): run n iterations on every handle concurrently, each
// on its own stream, and time the whole batch on the device (event on the first handle's stream before, after all).
extern "C" int qgmap_batch_step(qgmap_handle **hs, int nh, int n, int its, float *device_ms, long long *launches)
{
    qgmap_handle *nh0 = nullptr;
    if (!hs || nh < 1) QG_FAIL(nh0, QGMAP_ERR_ARG, "qgmap_batch_step: empty batch");
    for (int i = 0; i < nh; ++i)
        if (!hs[i] || hs[i]->device != hs[0]->device) QG_FAIL(nh0, QGMAP_ERR_ARG, "qgmap_batch_step: handles must share a device");
    qgmap_handle *h0 = hs[0];
    QG_CUDA(h0, cudaSetDevice(h0->device));
    if (!h0->evb0) { QG_CUDA(h0, cudaEventCreate(&h0->evb0)); QG_CUDA(h0, cudaEventCreate(&h0->evb1)); }
    for (int i = 0; i < nh; ++i) QG_CUDA(hs[i], cudaStreamSynchronize(hs[i]->stream));
    QG_CUDA(h0, cudaEventRecord(h0->evb0, h0->stream));
    for (int i = 1; i < nh; ++i) QG_CUDA(hs[i], cudaStreamWaitEvent(hs[i]->stream, h0->evb0, 0));
    int rc;
    long long nl = 0;
    for (int i = 0; i < nh; ++i) {
        if ((rc = qgmap_step_begin(hs[i], n, its)) != QGMAP_OK) { g_last_error = hs[i]->err; return rc; }
        nl += hs[i]->last_launches;
    }
    for (int i = 1; i < nh; ++i) QG_CUDA(h0, cudaStreamWaitEvent(h0->stream, hs[i]->ev1, 0));
    QG_CUDA(h0, cudaEventRecord(h0->evb1, h0->stream));
    for (int i = 0; i < nh; ++i)
        if ((rc = qgmap_step_end(hs[i], nullptr, nullptr, nullptr, nullptr, nullptr)) != QGMAP_OK) { g_last_error = hs[i]->err; return rc; }
    QG_CUDA(h0, cudaEventSynchronize(h0->evb1));
    float ms = 0.f;
    QG_CUDA(h0, cudaEventElapsedTime(&ms, h0->evb0, h0->evb1));
    if (device_ms) *device_ms = ms;
    if (launches) *launches = nl;
    return QGMAP_OK;
}

// Measured FP32 FMA throughput of the device (register-resident FFMA chains): the roofline denominator bench.py reports
// beside the nominal 148 SM x 128 lanes x 2 x clock figure.
__global__ void qgmap_ffma_probe_kernel(float *out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int qgmap_fp32_peak(int device, double *tflops)
{
    qgmap_handle *nh = nullptr;
    if (!tflops) return QGMAP_ERR_ARG;
    int dev;
    int rc = device_check(nh, device, &dev);
    if (rc) return rc;
    cudaDeviceProp prop;
    QG_CUDA(nh, cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float *d = nullptr;
    QG_CUDA(nh, cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        qgmap_ffma_probe_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 128.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(d);
    if (e != cudaSuccess) QG_FAIL(nh, QGMAP_ERR_CUDA, "qgmap_fp32_peak: %s", cudaGetErrorString(e));
    *tflops = best;
    return QGMAP_OK;
}

extern "C" int qgmap_last_step_ms(const qgmap_handle *h, float *ms)
{
    if (!h || !ms) return QGMAP_ERR_ARG;
    *ms = h->last_ms;
    return QGMAP_OK;
}
extern "C" int qgmap_last_launches(const qgmap_handle *h, long long *launches)
{
    if (!h || !launches) return QGMAP_ERR_ARG;
    *launches = h->last_launches;
    return QGMAP_OK;
}

// ---- monitoring ------------------------------------------------------------------------------------------------------
static QgMonArgs mon_params(const qgmap_handle *h)
{
    QgMonArgs q;
    q.I1 = h->I1d; q.pitchI = h->pitchI; q.VV = h->VVd; q.pitchV = h->pitchV;
    q.Mo = h->Mo; q.No = h->No; q.M = h->M; q.N = h->N; q.super = h->cfg.variant == QGMAP_VARIANT_SUPER;
    q.lambdad = h->cfg.lambdad; q.lambdas = h->cfg.lambdas; q.epsn = h->cfg.epsn;
    q.r0 = 0; q.r1 = h->M;
    return q;
}

static int ensure_map(qgmap_handle *h)
{
    if (!h->d_map) QG_CUDA(h, cudaMalloc(&h->d_map, (size_t)h->M * h->N * 2 * sizeof(double)));
    return QGMAP_OK;
}

// MAP of the current beliefs into h->d_map (device, column-major M x N x 2): the rows the handle owns, or (stored = true) every
// row it stores -- a band's halo rows carry the neighbours' current beliefs, so the band can evaluate its own share of
// profile_logP without seeing anybody else's map
static int map_device(qgmap_handle *h, bool stored = false)
{
    int rc = ensure_map(h);
    if (rc) return rc;
    const int cur = (h->ctrl_host->it - 1) & 1;
    const float *b = h->buf[cur];
    const long long fs = (long long)h->L * h->plane;
    const long long tot = 2LL * h->M * h->N;
    qgmap_launch_find_map_f32(&h->ctrl->alpha[0], b + F_MUU * fs, b + F_SIGU * fs, b + F_MUV * fs, b + F_SIGV * fs, h->plane,
                              h->M, h->N, h->L, h->P, h->g0, stored ? h->g0 : h->row_begin, stored ? h->g1 : h->row_end, h->d_map, tot, h->stream);
    QG_CUDA(h, cudaGetLastError());
    return QGMAP_OK;
}

extern "C" int qgmap_get_map(qgmap_handle *h, double *map)
{
    if (!h || !map) return QGMAP_ERR_ARG;
    if (!h->has_state) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_get_map: no state set");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = sync_ctrl(h);
    if (rc) return rc;
    if ((rc = map_device(h)) != QGMAP_OK) return rc;
    if (h->row_begin == 0 && h->row_end == h->M)
        QG_CUDA(h, cudaMemcpyAsync(map, h->d_map, (size_t)h->M * h->N * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    else      // a band fills only the rows it owns (like qgmap_get_state)
        QG_CUDA(h, cudaMemcpy2DAsync(map + h->row_begin, (size_t)h->M * sizeof(double), h->d_map + h->row_begin, (size_t)h->M * sizeof(double),
                                     (size_t)(h->row_end - h->row_begin) * sizeof(double), (size_t)h->N * 2, cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return QGMAP_OK;
}

static int reduce_partials(qgmap_handle *h, int nblk, double *out)
{
    std::vector<double> hp(nblk);
    QG_CUDA(h, cudaMemcpyAsync(hp.data(), h->mon_partials, nblk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (double v : hp) s += v;
    *out = s;
    return QGMAP_OK;
}

static int logp_device(qgmap_handle *h, const double *d_map, double *lp)
{
    const int nblk = 592;    // 4 x 148 SMs
    qgmap_launch_logp(mon_params(h), d_map, h->mon_partials, nblk, h->stream);
    QG_CUDA(h, cudaGetLastError());
    return reduce_partials(h, nblk, lp);
}

extern "C" int qgmap_logp(qgmap_handle *h, const double *map, double *lp)
{
    if (!h || !map || !lp) return QGMAP_ERR_ARG;
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_map(h);
    if (rc) return rc;
    QG_CUDA(h, cudaMemcpyAsync(h->d_map, map, (size_t)h->M * h->N * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return logp_device(h, h->d_map, lp);
}

static int upload_truth(qgmap_handle *h, const double *tflow, const uint8_t *unknown)
{
    const size_t MoNo = (size_t)h->Mo * h->No;
    if (!h->d_tflow) QG_CUDA(h, cudaMalloc(&h->d_tflow, MoNo * 2 * sizeof(double)));
    QG_CUDA(h, cudaMemcpyAsync(h->d_tflow, tflow, MoNo * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->has_unknown = unknown != nullptr;
    if (unknown) {
        if (!h->d_unknown) QG_CUDA(h, cudaMalloc(&h->d_unknown, MoNo));
        QG_CUDA(h, cudaMemcpyAsync(h->d_unknown, unknown, MoNo, cudaMemcpyHostToDevice, h->stream));
    }
    return QGMAP_OK;
}

static int aepe_device(qgmap_handle *h, const double *d_map, double *aepe)
{
    const int nblk = 592;
    qgmap_launch_aepe(mon_params(h), d_map, h->d_tflow, h->has_unknown ? h->d_unknown : nullptr, h->mon_partials, nblk, h->stream);
    QG_CUDA(h, cudaGetLastError());
    double s;
    int rc = reduce_partials(h, nblk, &s);
    if (rc) return rc;
    const int b = h->cfg.variant == QGMAP_VARIANT_SUPER ? 4 : 1;
    *aepe = s / ((double)(h->Mo - 2 * b) * (double)(h->No - 2 * b));
    return QGMAP_OK;
}

extern "C" int qgmap_aepe(qgmap_handle *h, const double *map, const double *tflow, const uint8_t *unknown, double *aepe)
{
    if (!h || !map || !tflow || !aepe) return QGMAP_ERR_ARG;
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_map(h);
    if (rc) return rc;
    if ((rc = upload_truth(h, tflow, unknown)) != QGMAP_OK) return rc;
    QG_CUDA(h, cudaMemcpyAsync(h->d_map, map, (size_t)h->M * h->N * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return aepe_device(h, h->d_map, aepe);
}

extern "C" int qgmap_set_truth(qgmap_handle *h, const double *tflow, const uint8_t *unknown)
{
    if (!h || !tflow) return QGMAP_ERR_ARG;
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = upload_truth(h, tflow, unknown);
    if (rc) return rc;
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    return QGMAP_OK;
}

// Monitoring block :52-68 for the rows this handle OWNS, entirely on its device: MAP of the stored rows, the handle's share of
// profile_logP (:148-154; the edges that start in an owned row) and of the AEPE numerator (:63-64).  A row-band run adds the shares
// in band order and divides the AEPE sum by (Mo-2b)(No-2b); for a whole-grid handle the shares are the totals.
extern "C" int qgmap_monitor_partial(qgmap_handle *h, double *logp_share, double *aepe_sum_share)
{
    if (!h) return QGMAP_ERR_ARG;
    if (!h->has_state) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_monitor_partial: no state set");
    if (aepe_sum_share && !h->d_tflow) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_monitor_partial: AEPE asked for before qgmap_set_truth");
    QG_CUDA(h, cudaSetDevice(h->device));
    int rc = sync_ctrl(h);
    if (rc) return rc;
    if ((rc = map_device(h, true)) != QGMAP_OK) return rc;
    QgMonArgs q = mon_params(h);
    q.r0 = h->row_begin; q.r1 = h->row_end;
    const int nblk = 592;
    if (logp_share) {
        qgmap_launch_logp(q, h->d_map, h->mon_partials, nblk, h->stream);
        QG_CUDA(h, cudaGetLastError());
        if ((rc = reduce_partials(h, nblk, logp_share)) != QGMAP_OK) return rc;
    }
    if (aepe_sum_share) {
        qgmap_launch_aepe(q, h->d_map, h->d_tflow, h->has_unknown ? h->d_unknown : nullptr, h->mon_partials, nblk, h->stream);
        QG_CUDA(h, cudaGetLastError());
        if ((rc = reduce_partials(h, nblk, aepe_sum_share)) != QGMAP_OK) return rc;
    }
    return QGMAP_OK;
}

extern "C" int qgmap_debug_gradients(qgmap_handle *h, double *G_muu, double *G_muv, double *G_sigu, double *G_sigv,
                                     double *dpn, double *drou, double *e_px, double *da_px)
{
    return qg_guard([&]() -> int {
    if (!h) return QGMAP_ERR_ARG;
    if (!h->has_state) QG_FAIL(h, QGMAP_ERR_STATE, "qgmap_debug_gradients: no state set");
    QG_CUDA(h, cudaSetDevice(h->device));
    const size_t nf = (size_t)11 * h->L * h->plane;
    if (!h->dbg) QG_CUDA(h, cudaMalloc(&h->dbg, nf * sizeof(float)));
    QG_CUDA(h, cudaMemsetAsync(h->dbg, 0, nf * sizeof(float), h->stream));
    h->params.dbg = h->dbg;
    launch_iter(h, true);
    QG_CUDA(h, cudaGetLastError());
    const int L = h->L;
    const size_t n = (size_t)h->M * h->N * L;
    int rc = ensure_stage(h, n);
    if (rc) return rc;
    double *outs[11] = {G_muu, G_muv, G_sigu, G_sigv, dpn, drou, drou ? drou + n : nullptr, drou ? drou + 2 * n : nullptr,
                        drou ? drou + 3 * n : nullptr, e_px, da_px};
    dim3 tb(32, 8), tg((h->N + 31) / 32, (h->row_end - h->row_begin + 31) / 32, L);
    for (int f = 0; f < 11; ++f) {
        if (!outs[f]) continue;
        QG_CUDA(h, cudaMemsetAsync(h->stage, 0, n * sizeof(double), h->stream));
        qgmap_export_kernel<float><<<tg, tb, 0, h->stream>>>(h->dbg + (size_t)f * L * h->plane, h->P, h->plane, h->g0,
                                                              h->row_begin, h->row_end, h->stage, h->M, h->N);
        QG_CUDA(h, cudaMemcpyAsync(outs[f], h->stage, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        QG_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return QGMAP_OK;
    });
}

// ---- stateless get_map_mex replacement ---------------------------------------------------------------------------------
extern "C" int qgmap_find_map(const double *alpha, const double *mu_u, const double *sig_u, const double *mu_v,
                              const double *sig_v, int M, int N, int L, double *map, int device)
{
    qgmap_handle *nh = nullptr;
    if (!alpha || !mu_u || !sig_u || !mu_v || !sig_v || !map) QG_FAIL(nh, QGMAP_ERR_ARG, "qgmap_find_map: NULL argument");
    if (M < 1 || N < 1 || L < 1 || L > QGMAP_LMAX) QG_FAIL(nh, QGMAP_ERR_ARG, "qgmap_find_map: bad size M=%d N=%d L=%d (L<=%d)", M, N, L, QGMAP_LMAX);
    int dev;
    int rc = device_check(nh, device, &dev);
    if (rc) return rc;
    const size_t n = (size_t)M * N * L, MN = (size_t)M * N;
    double *d = nullptr;
    QG_CUDA(nh, cudaMalloc(&d, (4 * n + 2 * MN + QGMAP_LMAX) * sizeof(double)));
    double *d_mu_u = d, *d_sig_u = d + n, *d_mu_v = d + 2 * n, *d_sig_v = d + 3 * n, *d_map = d + 4 * n, *d_alpha = d_map + 2 * MN;
    cudaError_t e = cudaSuccess;
    auto cp = [&](double *dst, const double *src, size_t cnt) { if (e == cudaSuccess) e = cudaMemcpy(dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice); };
    cp(d_mu_u, mu_u, n); cp(d_sig_u, sig_u, n); cp(d_mu_v, mu_v, n); cp(d_sig_v, sig_v, n); cp(d_alpha, alpha, L);
    if (e == cudaSuccess) {
        qgmap_launch_find_map_f64(d_alpha, d_mu_u, d_sig_u, d_mu_v, d_sig_v, (long long)MN, M, N, L, d_map, 2LL * MN, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(map, d_map, 2 * MN * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) QG_FAIL(nh, QGMAP_ERR_CUDA, "qgmap_find_map: %s", cudaGetErrorString(e));
    return QGMAP_OK;
}

// ---- one-call solver == gqmap_gpu_mixture(options,I1,I2) -----------------------------------------------------------------
// options.dir (gqmap_gpu_mixture.m:62): colour-coded MAP flow of the monitored iterations as <dir>/<it>.png
static thread_local std::string g_dump_dir;
extern "C" int qgmap_solve_set_dump_dir(const char *dir)
{
    return qg_guard([&]() -> int {
    g_dump_dir = dir ? dir : "";
    return QGMAP_OK;
    });
}
static int dump_map_png(qgmap_handle *h, int it)
{
    const int M = h->M, N = h->N;
    std::vector<double> map((size_t)M * N * 2);
    QG_CUDA(h, cudaMemcpyAsync(map.data(), h->d_map, map.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QG_CUDA(h, cudaStreamSynchronize(h->stream));
    int Mi = M, Ni = N;
    std::vector<double> flow;
    if (h->cfg.variant == QGMAP_VARIANT_SUPER) {                 // flow = repelem(map,4,4); flc = flowToColor_mex(flow(5:end-4,5:end-4,:)) (S:58-59)
        Mi = 4 * M - 8; Ni = 4 * N - 8;
        if (Mi < 1 || Ni < 1) return QGMAP_OK;
        flow.resize((size_t)Mi * Ni * 2);
        for (int c = 0; c < 2; ++c)
            for (int n = 0; n < Ni; ++n)
                for (int m = 0; m < Mi; ++m)
                    flow[(size_t)m + (size_t)Mi * n + (size_t)Mi * Ni * c] = map[(size_t)((m + 4) / 4) + (size_t)M * ((n + 4) / 4) + (size_t)M * N * c];
    } else flow.swap(map);
    std::vector<uint8_t> img((size_t)Mi * Ni * 3);
    int rc = qgmap_flow_to_color(flow.data(), Mi, Ni, -1.0, img.data(), nullptr, nullptr, nullptr);
    if (rc == QGMAP_OK) rc = qgmap_write_png((g_dump_dir + "/" + std::to_string(it) + ".png").c_str(), img.data(), Mi, Ni);
    if (rc != QGMAP_OK) QG_FAIL(h, rc, "could not write %s/%d.png (options.dir must exist)", g_dump_dir.c_str(), it);
    return QGMAP_OK;
}

extern "C" int qgmap_solve(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int its,
                           const double *const *init, uint64_t seed, const double *tflow, const uint8_t *unknown,
                           double *mu, double *sigma, double *alpha, double *AEPE, double *Energy, double *logP,
                           int *its_done)
{
    return qg_guard([&]() -> int {
    qgmap_handle *nh = nullptr;
    if (its < 1) QG_FAIL(nh, QGMAP_ERR_ARG, "qgmap_solve: its=%d", its);
    qgmap_handle *h = nullptr;
    int rc = qgmap_create(cfg, I1, I2, Mo, No, &h);
    if (rc) return rc;
    auto bail = [&](int code) { g_last_error = h->err; qgmap_destroy(h); return code; };
    if (init) rc = qgmap_set_state(h, init[0], init[1], init[2], init[3], init[4], init[5], init[6], nullptr, cfg->temperature, 1);
    else rc = qgmap_init_state(h, seed);
    if (rc) return bail(rc);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int i = 0; i < its; ++i) {                                                  // :16
        if (AEPE) AEPE[i] = nan;
        if (Energy) Energy[i] = 0.0;
        if (logP) logP[i] = nan;
    }
    if (tflow && (rc = upload_truth(h, tflow, unknown)) != QGMAP_OK) return bail(rc);
    const int every = cfg->log_every > 0 ? cfg->log_every : 300;
    int it = 1, stopped = 0;
    long long launches = 0;
    float ms = 0.f;
    while (!stopped && it <= its) {
        // run up to and including the next monitored iteration (:52: mod(it,300)==0 || it==1)
        const int next_mon = (it == 1) ? 1 : ((it + every - 1) / every) * every;
        const int n = std::min(next_mon, its) - it + 1;
        int done = 0;
        rc = qgmap_step(h, n, its, Energy ? Energy + (it - 1) : nullptr, nullptr, nullptr, &done, &stopped);
        if (rc) return bail(rc);
        launches += h->last_launches; ms += h->last_ms;
        it += done;
        const int last = it - 1;                                                      // last executed iteration
        if (done > 0 && (last == 1 || last % every == 0)) {                           // :52-68
            if ((rc = map_device(h)) != QGMAP_OK) return bail(rc);
            if (!g_dump_dir.empty() && (rc = dump_map_png(h, last)) != QGMAP_OK) return bail(rc);    // :59-62
            if (tflow && AEPE && (rc = aepe_device(h, h->d_map, &AEPE[last - 1])) != QGMAP_OK) return bail(rc);
            if (logP && (rc = logp_device(h, h->d_map, &logP[last - 1])) != QGMAP_OK) return bail(rc);
            launches += 1 + (tflow && AEPE ? 1 : 0) + (logP ? 1 : 0);
        }
        if (done < n) break;
    }
    if (its_done) *its_done = it - 1;
    const size_t n3 = (size_t)h->M * h->N * h->L;
    std::vector<double> a(h->L);
    rc = qgmap_get_state(h, mu, mu ? mu + n3 : nullptr, sigma, sigma ? sigma + n3 : nullptr, nullptr, nullptr, nullptr,
                         a.data(), nullptr, nullptr);                                 // :183-185
    if (rc) return bail(rc);
    if (alpha) std::copy(a.begin(), a.end(), alpha);
    g_solve_launches = launches; g_solve_ms = ms;
    qgmap_destroy(h);
    return QGMAP_OK;
    });
}

// gqmap_gpu_mixture(options,I1,I2) with options.devices: the loop of qgmap_solve over a band group
extern "C" int qgmap_group_solve(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int its,
                                 int nbands, const int *devices,
                                 const double *const *init, uint64_t seed, const double *tflow, const uint8_t *unknown,
                                 double *mu, double *sigma, double *alpha, double *AEPE, double *Energy, double *logP,
                                 int *its_done)
{
    return qg_guard([&]() -> int {
    qgmap_handle *nh = nullptr;
    if (its < 1 || nbands < 1) QG_FAIL(nh, QGMAP_ERR_ARG, "qgmap_group_solve: its=%d nbands=%d", its, nbands);
    qgmap_group *g = nullptr;
    int rc = qgmap_group_create(cfg, I1, I2, Mo, No, nbands, devices, &g);
    if (rc) return rc;
    auto bail = [&](int code) { qgmap_group_destroy(g); return code; };
    if (init) rc = qgmap_group_set_state(g, init[0], init[1], init[2], init[3], init[4], init[5], init[6], nullptr, cfg->temperature, 1);
    else rc = qgmap_group_init_state(g, seed);
    if (rc) return bail(rc);
    int M = 0, N = 0, L = 0;
    qgmap_group_dims(g, &M, &N, &L, nullptr);
    const size_t n3 = (size_t)M * N * L;
    std::vector<double> mu_tmp, sg_tmp, al(L), map((size_t)M * N * 2);
    if (!mu) { mu_tmp.resize(2 * n3); mu = mu_tmp.data(); }
    if (!sigma) { sg_tmp.resize(2 * n3); sigma = sg_tmp.data(); }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int i = 0; i < its; ++i) {
        if (AEPE) AEPE[i] = nan;
        if (Energy) Energy[i] = 0.0;
        if (logP) logP[i] = nan;
    }
    qgmap_handle *h0 = qgmap_group_band(g, 0);                                       // the PNG dump is rendered from band 0's device
    if (tflow && AEPE)
        for (int b = 0; b < nbands; ++b)
            if ((rc = qgmap_set_truth(qgmap_group_band(g, b), tflow, unknown)) != QGMAP_OK) { g_last_error = qgmap_group_band(g, b)->err; return bail(rc); }
    const int every = cfg->log_every > 0 ? cfg->log_every : 300;
    int it = 1, stopped = 0;
    float ms = 0.f;
    while (!stopped && it <= its) {
        const int next_mon = (it == 1) ? 1 : ((it + every - 1) / every) * every;
        const int n = std::min(next_mon, its) - it + 1;
        int done = 0;
        rc = qgmap_group_step(g, n, its, Energy ? Energy + (it - 1) : nullptr, nullptr, nullptr, &done, &stopped);
        if (rc) return bail(rc);
        float m1 = 0.f;
        qgmap_group_last_step_ms(g, &m1);
        ms += m1;
        it += done;
        const int last = it - 1;
        if (done > 0 && (last == 1 || last % every == 0) && (AEPE || logP || !g_dump_dir.empty())) {     // :52-68
            // every band extracts the MAP of the rows it stores and sums its own share of logP / AEPE on its own GPU (:52-67);
            // only the scalars travel.  The colour-coded PNG (:59-62) needs the whole map: gathered only when options.dir is set.
            double lp = 0.0, ae = 0.0;
            const bool want_ae = tflow && AEPE;
            for (int b = 0; b < nbands; ++b) {
                qgmap_handle *hb = qgmap_group_band(g, b);
                double lpb = 0.0, aeb = 0.0;
                if ((rc = qgmap_monitor_partial(hb, logP ? &lpb : nullptr, want_ae ? &aeb : nullptr)) != QGMAP_OK) { g_last_error = hb->err; return bail(rc); }
                lp += lpb; ae += aeb;                                                                    // fixed band order
            }
            if (logP) logP[last - 1] = lp;
            if (want_ae) {
                const int bw = cfg->variant == QGMAP_VARIANT_SUPER ? 4 : 1;
                AEPE[last - 1] = ae / ((double)(Mo - 2 * bw) * (double)(No - 2 * bw));
            }
            if (!g_dump_dir.empty()) {
                for (int b = 0; b < nbands; ++b) {
                    qgmap_handle *hb = qgmap_group_band(g, b);
                    if ((rc = qgmap_get_map(hb, map.data())) != QGMAP_OK) { g_last_error = hb->err; return bail(rc); }
                }
                QG_CUDA(h0, cudaSetDevice(h0->device));
                if ((rc = ensure_map(h0)) != QGMAP_OK) return bail(rc);
                QG_CUDA(h0, cudaMemcpyAsync(h0->d_map, map.data(), map.size() * sizeof(double), cudaMemcpyHostToDevice, h0->stream));
                if ((rc = dump_map_png(h0, last)) != QGMAP_OK) { g_last_error = h0->err; return bail(rc); }
            }
        }
        if (done < n) break;
    }
    if (its_done) *its_done = it - 1;
    rc = qgmap_group_get_state(g, mu, mu + n3, sigma, sigma + n3, nullptr, nullptr, nullptr, al.data(), nullptr, nullptr);   // :183-185
    if (rc) return bail(rc);
    if (alpha) std::copy(al.begin(), al.end(), alpha);
    g_solve_launches = 0; g_solve_ms = ms;
    qgmap_group_destroy(g);
    return QGMAP_OK;
    });
}

void qgmap_launch_iteration(const qgmap_handle *h) { launch_iter(h, false); }
void qgmap_launch_advance(const qgmap_handle *h) { qgmap_advance_kernel<<<1, 32, 0, h->stream>>>(h->params); }

extern "C" int qgmap_last_solve_stats(long long *launches, float *kernel_ms)
{
    if (launches) *launches = g_solve_launches;
    if (kernel_ms) *kernel_ms = g_solve_ms;
    return QGMAP_OK;
}
