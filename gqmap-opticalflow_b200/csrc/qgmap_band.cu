// qgmap_band.cu -- row-band decomposition of ONE frame pair over several GPUs (BASELINE configs[3]).
//
// A band handle owns rows [row_begin,row_end) of the belief grid and stores one halo row on each side.  The 4-neighbour
// coupling reaches exactly one row, so per iteration each band needs
//   * from the band ABOVE: its last row's (mu_u,mu_v,sig_u,sig_v) and down-edge correlations -- the halo warp recomputes that
//     row's down edge, whose endpoint-2 gradient belongs to our first row (gqmap_gpu_mixture.m:37-40);
//   * from the band BELOW: its first row's (mu_u,mu_v,sig_u,sig_v) -- the neighbour values of our last row's down edges (:32,:34);
//   * the GLOBAL sums of this iteration (Energy, d(alpha) per component, sum|G_mu_u|, sum|G_sig_u|: 4L doubles), because
//     alpha(t+1) and the stop test depend on them (:36,:48,:50,:69-75).
// Whole rows of all 9 fields are exchanged (9*L*N floats per boundary and direction; 4K, L=3: 415 KB) -- bandwidth is
// irrelevant at NVLink speed, latency is what counts.
//
// Two transports share the iteration structure:
//   (1) qgmap_group_*: ONE process drives all bands (the natural model for a MATLAB host): streams + events, halo rows
//       copied directly between the bands' state buffers (cudaMemcpy2DAsync, peer-to-peer over NVLink when the bands sit on
//       different GPUs), sums read through peer pointers in fixed band order (bit-deterministic).  Also runs with all bands
//       on one GPU, which is how the decomposition is tested on a 1-GPU box.
//   (2) qgmap_band_connect: one process per GPU (torchrun): NCCL send/recv of packed boundary rows + a 4L-double all-reduce.
#include "qgmap_internal.h"
#include "qgmap_advance.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define QGB_FAIL(h, code, ...)                                        \
    do {                                                              \
        char _b[512];                                                 \
        snprintf(_b, sizeof _b, __VA_ARGS__);                         \
        if (h) (h)->err = _b;                                         \
        qgmap_set_last_error(_b);                                     \
        return (code);                                                \
    } while (0)
#define QGB_CUDA(h, expr)                                             \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) QGB_FAIL(h, QGMAP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL, resolved lazily so that a process which already loaded an NCCL (torch) keeps using that one -------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load(std::string *why)
{
    if (g_nccl.lib) return true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { *why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define QGB_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { *why = std::string("missing symbol ") + name; return false; }
    QGB_SYM(GetUniqueId, "ncclGetUniqueId") QGB_SYM(CommInitRank, "ncclCommInitRank") QGB_SYM(CommDestroy, "ncclCommDestroy")
    QGB_SYM(AllReduce, "ncclAllReduce") QGB_SYM(Send, "ncclSend") QGB_SYM(Recv, "ncclRecv") QGB_SYM(GroupStart, "ncclGroupStart")
    QGB_SYM(GroupEnd, "ncclGroupEnd") QGB_SYM(GetErrorString, "ncclGetErrorString")
#undef QGB_SYM
    g_nccl.lib = lib;
    return true;
}
#define QGB_NCCL(h, expr)                                             \
    do {                                                              \
        ncclResult_t _r = (expr);                                     \
        if (_r != ncclSuccess) QGB_FAIL(h, QGMAP_ERR_COMM, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

struct QgBand {
    ncclComm_t comm = nullptr;
    float *send_up = nullptr, *send_dn = nullptr, *recv_up = nullptr, *recv_dn = nullptr;   // [9L][N] each
};

void qgmap_band_release(qgmap_handle *h)
{
    if (!h->band) return;
    if (h->band->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->band->comm);
    for (float *p : {h->band->send_up, h->band->send_dn, h->band->recv_up, h->band->recv_dn}) if (p) cudaFree(p);
    delete h->band;
    h->band = nullptr;
}

int qgmap_band_refresh(qgmap_handle *h) { (void)h; return QGMAP_OK; }   // set_state imports the halo rows with the band

extern "C" int qgmap_band_unique_id(void *id128)
{
    return qg_guard([&]() -> int {
    qgmap_handle *nh = nullptr;
    if (!id128) return QGMAP_ERR_ARG;
    std::string why;
    if (!nccl_load(&why)) QGB_FAIL(nh, QGMAP_ERR_COMM, "NCCL unavailable: %s", why.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    QGB_NCCL(nh, g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return QGMAP_OK;
    });
}

extern "C" int qgmap_band_connect(qgmap_handle *h, int rank, int nranks, const void *id128)
{
    return qg_guard([&]() -> int {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return QGMAP_ERR_ARG;
    if (h->in_group) QGB_FAIL(h, QGMAP_ERR_STATE, "handle belongs to a qgmap_group");
    std::string why;
    if (!nccl_load(&why)) QGB_FAIL(h, QGMAP_ERR_COMM, "NCCL unavailable: %s", why.c_str());
    QGB_CUDA(h, cudaSetDevice(h->device));
    qgmap_band_release(h);
    h->band = new QgBand();
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    QGB_NCCL(h, g_nccl.CommInitRank(&h->band->comm, nranks, id, rank));
    const size_t n = (size_t)F_COUNT * h->L * h->N * sizeof(float);
    QGB_CUDA(h, cudaMalloc(&h->band->send_up, n)); QGB_CUDA(h, cudaMalloc(&h->band->send_dn, n));
    QGB_CUDA(h, cudaMalloc(&h->band->recv_up, n)); QGB_CUDA(h, cudaMalloc(&h->band->recv_dn, n));
    h->rank = rank; h->nranks = nranks;
    h->params.band = nranks > 1 ? 1 : 0;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return QGMAP_OK;
    });
}

// row `grow` (global index) of all 9L planes of state buffer b  <->  contiguous [9L][N]
static cudaError_t pack_row(qgmap_handle *h, int b, int grow, float *dst)
{
    return cudaMemcpy2DAsync(dst, (size_t)h->N * sizeof(float), h->buf[b] + (size_t)(grow - h->g0) * h->P, (size_t)h->plane * sizeof(float),
                             (size_t)h->N * sizeof(float), (size_t)F_COUNT * h->L, cudaMemcpyDeviceToDevice, h->stream);
}
static cudaError_t unpack_row(qgmap_handle *h, int b, int grow, const float *src)
{
    return cudaMemcpy2DAsync(h->buf[b] + (size_t)(grow - h->g0) * h->P, (size_t)h->plane * sizeof(float), src, (size_t)h->N * sizeof(float),
                             (size_t)h->N * sizeof(float), (size_t)F_COUNT * h->L, cudaMemcpyDeviceToDevice, h->stream);
}

// One iteration in multi-process band mode (called by qgmap_step_begin when nranks > 1).
int qgmap_band_iteration(qgmap_handle *h, long long *launches)
{
    QgBand *bd = h->band;
    if (!bd || !bd->comm) QGB_FAIL(h, QGMAP_ERR_COMM, "band handle not connected (qgmap_band_connect)");
    // host-side iteration count: all ranks stop together (they see the same all-reduced sums), so parity stays consistent
    const int it = h->ctrl_host->it + h->band_steps_enqueued;
    const int nb = it & 1;                                       // buffer this iteration writes
    qgmap_launch_iteration(h);
    QGB_NCCL(h, g_nccl.AllReduce(h->ctrl->sums, h->ctrl->sums, (size_t)QG_NRED * h->L, ncclDouble, ncclSum, bd->comm, h->stream));
    qgmap_launch_advance(h);
    const size_t cnt = (size_t)F_COUNT * h->L * h->N;
    const bool up = h->rank > 0, dn = h->rank < h->nranks - 1;
    if (up) QGB_CUDA(h, pack_row(h, nb, h->row_begin, bd->send_up));
    if (dn) QGB_CUDA(h, pack_row(h, nb, h->row_end - 1, bd->send_dn));
    QGB_NCCL(h, g_nccl.GroupStart());
    if (up) { QGB_NCCL(h, g_nccl.Send(bd->send_up, cnt, ncclFloat, h->rank - 1, bd->comm, h->stream));
              QGB_NCCL(h, g_nccl.Recv(bd->recv_up, cnt, ncclFloat, h->rank - 1, bd->comm, h->stream)); }
    if (dn) { QGB_NCCL(h, g_nccl.Send(bd->send_dn, cnt, ncclFloat, h->rank + 1, bd->comm, h->stream));
              QGB_NCCL(h, g_nccl.Recv(bd->recv_dn, cnt, ncclFloat, h->rank + 1, bd->comm, h->stream)); }
    QGB_NCCL(h, g_nccl.GroupEnd());
    if (up) QGB_CUDA(h, unpack_row(h, nb, h->row_begin - 1, bd->recv_up));
    if (dn) QGB_CUDA(h, unpack_row(h, nb, h->row_end, bd->recv_dn));
    h->band_steps_enqueued++;
    *launches += 2;
    return QGMAP_OK;
}

// ================================== single-process group of bands ==========================================================
struct qgmap_group {
    std::vector<qgmap_handle *> bands;
    std::vector<cudaEvent_t> ev_iter, ev_adv;     // per band: iteration kernel done / control block advanced + halos written
    std::vector<const double **> d_sumptrs;       // per band (device): pointers to every band's ctrl->sums
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int M = 0, N = 0, L = 0;
    float last_ms = 0.f;
    bool p2p = false;                             // bands on distinct GPUs: publish-kernel exchange (qgmap_p2p.cu)
    std::string err;
};

__global__ void qgmap_group_advance_kernel(const __grid_constant__ QgIterParams p, const double *const *sums, int nb)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || p.ctrl->stop) return;
    double tot[QG_LMAX * QG_NRED];
    for (int k = 0; k < p.L * QG_NRED; ++k) {
        double s = 0.0;
        for (int b = 0; b < nb; ++b) s += __ldcg(sums[b] + k);        // fixed band order: every band computes the same bits
        tot[k] = s;
    }
    qg_advance(p, p.ctrl, tot);
}

qgmap_handle *qgmap_group_band(qgmap_group *g, int b) { return g->bands[b]; }

extern "C" int qgmap_group_destroy(qgmap_group *g)
{
    if (!g) return QGMAP_ERR_ARG;
    for (size_t b = 0; b < g->bands.size(); ++b) {
        if (g->bands[b]) cudaSetDevice(g->bands[b]->device);
        if (b < g->ev_iter.size() && g->ev_iter[b]) cudaEventDestroy(g->ev_iter[b]);
        if (b < g->ev_adv.size() && g->ev_adv[b]) cudaEventDestroy(g->ev_adv[b]);
        if (b < g->d_sumptrs.size() && g->d_sumptrs[b]) cudaFree((void *)g->d_sumptrs[b]);
        if (g->bands[b]) qgmap_destroy(g->bands[b]);
    }
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    delete g;
    return QGMAP_OK;
}

extern "C" int qgmap_group_create(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int nbands,
                                  const int *devices, qgmap_group **out)
{
    return qg_guard([&]() -> int {
    qgmap_handle *nh = nullptr;
    if (!cfg || !I1 || !I2 || !out || nbands < 1) QGB_FAIL(nh, QGMAP_ERR_ARG, "qgmap_group_create: bad argument");
    *out = nullptr;
    const int M = cfg->variant == QGMAP_VARIANT_SUPER ? Mo / 4 : Mo;
    if (M / nbands < 2) QGB_FAIL(nh, QGMAP_ERR_ARG, "qgmap_group_create: %d bands over %d rows (need >= 2 rows per band)", nbands, M);
    qgmap_group *g = new qgmap_group();
    g->bands.assign(nbands, nullptr);
    g->ev_iter.assign(nbands, nullptr); g->ev_adv.assign(nbands, nullptr); g->d_sumptrs.assign(nbands, nullptr);
    auto bail = [&](int rc) { qgmap_group_destroy(g); return rc; };
    for (int b = 0; b < nbands; ++b) {
        qgmap_config c = *cfg;
        c.row_begin = (int)((long long)M * b / nbands);
        c.row_end = (int)((long long)M * (b + 1) / nbands);
        if (devices) c.device = devices[b];
        int rc = qgmap_create(&c, I1, I2, Mo, No, &g->bands[b]);
        if (rc) return bail(rc);
        qgmap_handle *h = g->bands[b];
        h->in_group = true; h->rank = b; h->nranks = nbands; h->params.band = nbands > 1 ? 1 : 0;
        if (cudaEventCreateWithFlags(&g->ev_iter[b], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->ev_adv[b], cudaEventDisableTiming) != cudaSuccess) return bail(QGMAP_ERR_CUDA);
    }
    g->M = g->bands[0]->M; g->N = g->bands[0]->N; g->L = g->bands[0]->L;
    // peer access between the bands' devices (NVLink P2P); bands may also share a device
    for (int a = 0; a < nbands; ++a)
        for (int b = 0; b < nbands; ++b) {
            const int da = g->bands[a]->device, db = g->bands[b]->device;
            if (da == db) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, da, db);
            if (!can) { qgmap_set_last_error("devices of a band group cannot access each other (no P2P)"); return bail(QGMAP_ERR_COMM); }
            cudaSetDevice(da);
            cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { qgmap_set_last_error(cudaGetErrorString(e)); return bail(QGMAP_ERR_CUDA); }
            cudaGetLastError();
        }
    std::vector<const double *> ptrs(nbands);
    for (int b = 0; b < nbands; ++b) ptrs[b] = g->bands[b]->ctrl->sums;
    for (int b = 0; b < nbands; ++b) {
        cudaSetDevice(g->bands[b]->device);
        if (cudaMalloc((void **)&g->d_sumptrs[b], nbands * sizeof(double *)) != cudaSuccess ||
            cudaMemcpy((void *)g->d_sumptrs[b], ptrs.data(), nbands * sizeof(double *), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(QGMAP_ERR_CUDA);
    }
    // every band on its own GPU: exchange through the publish kernel over peer memory (qgmap_p2p.cu) -- two launches per band
    // and iteration, no events.  Bands sharing a GPU keep the event-ordered copies below: a kernel that waits for another
    // kernel's flag must not rely on the two being co-scheduled on one device.
    bool distinct = nbands > 1;
    for (int a = 0; a < nbands && distinct; ++a)
        for (int b = a + 1; b < nbands; ++b) if (g->bands[a]->device == g->bands[b]->device) distinct = false;
    // QGMAP_GROUP_TRANSPORT=events forces the copies; =p2p-shared (tests only) runs the peer-memory kernels although bands share a
    // GPU -- their flag waits then rely on the device running the bands' streams concurrently (they time out into QGMAP_ERR_COMM
    // under a serialising tool).  A plain "p2p" never overrides the shared-GPU rule.
    if (const char *env = getenv("QGMAP_GROUP_TRANSPORT")) { if (!strcmp(env, "events")) distinct = false; else if (!strcmp(env, "p2p-shared") && nbands > 1) distinct = true; }
    if (distinct && nbands <= QGMAP_P2P_RANKS_MAX) {
        std::vector<char> blobs((size_t)nbands * QGMAP_P2P_BLOB_BYTES);
        for (int b = 0; b < nbands; ++b) { int rc = qgmap_band_p2p_export(g->bands[b], blobs.data() + (size_t)b * QGMAP_P2P_BLOB_BYTES); if (rc) return bail(rc); }
        for (int b = 0; b < nbands; ++b) { int rc = qgmap_band_p2p_connect(g->bands[b], b, nbands, blobs.data()); if (rc) return bail(rc); }
        g->p2p = true;
    }
    cudaSetDevice(g->bands[0]->device);
    if (cudaEventCreate(&g->ev0) != cudaSuccess || cudaEventCreate(&g->ev1) != cudaSuccess) return bail(QGMAP_ERR_CUDA);
    *out = g;
    return QGMAP_OK;
    });
}

extern "C" int qgmap_group_dims(const qgmap_group *g, int *M, int *N, int *L, int *nbands)
{
    if (!g) return QGMAP_ERR_ARG;
    if (M) *M = g->M;
    if (N) *N = g->N;
    if (L) *L = g->L;
    if (nbands) *nbands = (int)g->bands.size();
    return QGMAP_OK;
}

extern "C" int qgmap_group_set_state(qgmap_group *g, const double *muu, const double *muv, const double *sigu, const double *sigv,
                                     const double *pn, const double *rou, const double *w, const double *alpha, double T, int it)
{
    if (!g) return QGMAP_ERR_ARG;
    for (qgmap_handle *h : g->bands) {
        int rc = qgmap_set_state(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it);
        if (rc) { g->err = h->err; return rc; }
    }
    return QGMAP_OK;
}

extern "C" int qgmap_group_init_state(qgmap_group *g, uint64_t seed)
{
    return qg_guard([&]() -> int {
    if (!g) return QGMAP_ERR_ARG;
    qgmap_handle *h0 = g->bands[0];              // the full-grid arrays are drawn once; every band keeps its rows
    const size_t n = (size_t)g->M * g->N * g->L;
    std::vector<double> w, muu, muv, sigu, sigv, pn(n, 0.0), rou(n * 4, 0.0);
    qgmap_random_state(h0->cfg, n, g->L, seed, w, muu, muv, sigu, sigv);
    return qgmap_group_set_state(g, muu.data(), muv.data(), sigu.data(), sigv.data(), pn.data(), rou.data(), w.data(), nullptr,
                                 h0->cfg.temperature, 1);
    });
}

extern "C" int qgmap_group_get_state(qgmap_group *g, double *muu, double *muv, double *sigu, double *sigv, double *pn, double *rou,
                                     double *w, double *alpha, double *T, int *it)
{
    if (!g) return QGMAP_ERR_ARG;
    for (qgmap_handle *h : g->bands) {           // each band writes only the rows it owns
        int rc = qgmap_get_state(h, muu, muv, sigu, sigv, pn, rou, w, alpha, T, it);
        if (rc) { g->err = h->err; return rc; }
    }
    return QGMAP_OK;
}

extern "C" int qgmap_group_step(qgmap_group *g, int n, int its, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped)
{
    if (!g || n < 0 || its < 1) return QGMAP_ERR_ARG;
    const int nb = (int)g->bands.size();
    qgmap_handle *h0 = g->bands[0];
    for (qgmap_handle *h : g->bands) {
        if (!h->has_state) QGB_FAIL(h0, QGMAP_ERR_STATE, "qgmap_group_step before set_state/init_state");
        int rc = qgmap_prepare_step(h, n, its);
        if (rc) { g->err = h->err; return rc; }
    }
    const int it_start = h0->ctrl_host->it;
    QGB_CUDA(h0, cudaSetDevice(h0->device));
    QGB_CUDA(h0, cudaEventRecord(g->ev0, h0->stream));
    for (int b = 1; b < nb; ++b) { cudaSetDevice(g->bands[b]->device); QGB_CUDA(g->bands[b], cudaStreamWaitEvent(g->bands[b]->stream, g->ev0, 0)); }
    if (g->p2p) {
        for (qgmap_handle *h : g->bands) { QGB_CUDA(h, cudaSetDevice(h->device)); qgmap_p2p_begin_step(h); }
        if (qgmap_p2p_fused(h0)) {                                          // one launch per iteration: each band's stream gets its
            for (qgmap_handle *h : g->bands) {                             // whole run (graphs); the kernels pace each other by flags
                long long nl = 0;
                QGB_CUDA(h, cudaSetDevice(h->device));
                int rc = qgmap_enqueue_iterations(h, n, &nl);
                if (rc) { g->err = h->err; return rc; }
            }
        } else {
            for (int k = 0; k < n; ++k)
                for (int b = 0; b < nb; ++b) {
                    qgmap_handle *h = g->bands[b];
                    long long nl = 0;
                    QGB_CUDA(h, cudaSetDevice(h->device));
                    int rc = qgmap_p2p_iteration(h, &nl);
                    if (rc) { g->err = h->err; return rc; }
                }
        }
    }
    for (int k = 0; k < (g->p2p ? 0 : n); ++k) {
        const int wbuf = (it_start + k) & 1;                               // buffer written by this iteration
        for (int b = 0; b < nb; ++b) {                                     // 1. iteration kernels (need neighbours' halos: ev_adv waits below)
            qgmap_handle *h = g->bands[b];
            QGB_CUDA(h, cudaSetDevice(h->device));
            qgmap_launch_iteration(h);
            QGB_CUDA(h, cudaEventRecord(g->ev_iter[b], h->stream));
        }
        for (int b = 0; b < nb; ++b) {                                     // 2. global sums + control advance + halo pull
            qgmap_handle *h = g->bands[b];
            QGB_CUDA(h, cudaSetDevice(h->device));
            for (int o = 0; o < nb; ++o) if (o != b) QGB_CUDA(h, cudaStreamWaitEvent(h->stream, g->ev_iter[o], 0));
            if (nb > 1) qgmap_group_advance_kernel<<<1, 32, 0, h->stream>>>(h->params, g->d_sumptrs[b], nb);
            const size_t wbytes = (size_t)h->N * sizeof(float), rows = (size_t)F_COUNT * h->L;
            if (b > 0) {                                                   // pull the last row of the band above into our top halo
                qgmap_handle *u = g->bands[b - 1];
                QGB_CUDA(h, cudaMemcpy2DAsync(h->buf[wbuf] + (size_t)(h->row_begin - 1 - h->g0) * h->P, (size_t)h->plane * sizeof(float),
                                              u->buf[wbuf] + (size_t)(u->row_end - 1 - u->g0) * u->P, (size_t)u->plane * sizeof(float),
                                              wbytes, rows, cudaMemcpyDefault, h->stream));
            }
            if (b < nb - 1) {                                              // pull the first row of the band below into our bottom halo
                qgmap_handle *d = g->bands[b + 1];
                QGB_CUDA(h, cudaMemcpy2DAsync(h->buf[wbuf] + (size_t)(h->row_end - h->g0) * h->P, (size_t)h->plane * sizeof(float),
                                              d->buf[wbuf] + (size_t)(d->row_begin - d->g0) * d->P, (size_t)d->plane * sizeof(float),
                                              wbytes, rows, cudaMemcpyDefault, h->stream));
            }
            QGB_CUDA(h, cudaEventRecord(g->ev_adv[b], h->stream));
        }
        for (int b = 0; b < nb; ++b) {                                     // 3. nobody overwrites sums / rows a neighbour still reads
            qgmap_handle *h = g->bands[b];
            QGB_CUDA(h, cudaSetDevice(h->device));
            for (int o = 0; o < nb; ++o) if (o != b) QGB_CUDA(h, cudaStreamWaitEvent(h->stream, g->ev_adv[o], 0));
        }
    }
    QGB_CUDA(h0, cudaSetDevice(h0->device));
    for (int b = 1; b < nb; ++b) {
        cudaSetDevice(g->bands[b]->device);
        QGB_CUDA(g->bands[b], cudaEventRecord(g->ev_adv[b], g->bands[b]->stream));
        cudaSetDevice(h0->device);
        QGB_CUDA(h0, cudaStreamWaitEvent(h0->stream, g->ev_adv[b], 0));
    }
    QGB_CUDA(h0, cudaEventRecord(g->ev1, h0->stream));
    int done = 0, stop = 0;
    for (int b = 0; b < nb; ++b) {
        int d = 0, s = 0;
        int rc = qgmap_finish_step(g->bands[b], b == 0 ? energy : nullptr, b == 0 ? ptdmu : nullptr, b == 0 ? ptdsigma : nullptr, &d, &s);
        if (rc) { g->err = g->bands[b]->err; return rc; }
        if (b == 0) { done = d; stop = s; }
        else if (d != done) QGB_FAIL(h0, QGMAP_ERR_COMM, "bands disagree on the iteration count (%d vs %d)", d, done);
    }
    cudaSetDevice(h0->device);
    QGB_CUDA(h0, cudaEventSynchronize(g->ev1));
    QGB_CUDA(h0, cudaEventElapsedTime(&g->last_ms, g->ev0, g->ev1));
    if (n_done) *n_done = done;
    if (stopped) *stopped = stop;
    return QGMAP_OK;
}

extern "C" int qgmap_group_last_step_ms(const qgmap_group *g, float *ms)
{
    if (!g || !ms) return QGMAP_ERR_ARG;
    *ms = g->last_ms;
    return QGMAP_OK;
}
