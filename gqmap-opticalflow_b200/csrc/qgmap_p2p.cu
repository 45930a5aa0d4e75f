// qgmap_p2p.cu -- row bands of ONE frame pair over several GPUs with the halo exchange and the global sums done by our own
// kernel over NVLink peer memory (BASELINE configs[3]); no NCCL and no host on the per-iteration path.
//
// Why: the NCCL transport (qgmap_band.cu) needs seven stream operations per iteration (all-reduce, advance kernel, two packs,
// grouped send/recv, two unpacks) -- ~0.13 ms per iteration at 8 GPUs against 0.58 ms of compute.  Here ONE kernel follows
// the iteration kernel:
//   1. every CTA copies a slice of the band's first / last owned row (all 9L planes of the buffer just written) straight
//      into the neighbour's halo row of the same ping-pong buffer (peer stores, float4);
//   2. the last CTA to finish stores the band's 4L partial sums into slot [rank] of EVERY rank's mailbox, fences
//      (system scope) and raises flag[rank] = (step generation, iteration) in every mailbox;
//   3. it then waits until all flags of its own mailbox carry this iteration, adds the nranks partial sums in rank order
//      (same bits on every rank, same bits as the single-domain reduction order per band) and advances the control block
//      (alpha update, anneal, stop test: gqmap_gpu_mixture.m:36,48,50,69-75).
// Hazards: iteration t reads buffer A and writes B; peers write only halo rows of B during t.  A peer can start t+1 (writing
// our halo rows of A) only after it has seen OUR flag for t, i.e. after our iteration kernel t has completely finished
// reading A.  Flags are monotone 64-bit tags, so nothing is ever reset while a peer may still write.
// A peer that never shows up (crashed rank) must not hang the GPU: the wait gives up after ~10 s, stops the run and
// sets QgCtrl::comm_error, which qgmap_step_end turns into QGMAP_ERR_COMM.  Every qgmap_step call starts with a ready handshake
// (qgmap_p2p_ready_kernel) so that set_state on one rank can never race with a faster neighbour's first boundary rows.
#include "qgmap_internal.h"
#include "qgmap_advance.cuh"
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <vector>

#define QGP_FAIL(h, code, ...)                                        \
    do {                                                              \
        char _b[512];                                                 \
        snprintf(_b, sizeof _b, __VA_ARGS__);                         \
        if (h) (h)->err = _b;                                         \
        qgmap_set_last_error(_b);                                     \
        return (code);                                                \
    } while (0)
#define QGP_CUDA(h, expr)                                             \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) QGP_FAIL(h, QGMAP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define QG_RANKS_MAX QGMAP_P2P_RANKS_MAX

// One per band handle, in that band's device memory; slot q is written by rank q only.
struct QgMailbox {
    double sums[QG_RANKS_MAX][QG_LMAX * QG_NRED];
    unsigned long long flag[QG_RANKS_MAX];   // (generation << 32 | iteration) rank q has completely published
    unsigned int ticket;                     // CTAs of the local publish kernel that have finished their copy slice
};

struct QgP2PParams {
    QgMailbox *box[QG_RANKS_MAX];            // every rank's mailbox as mapped on this device (own included)
    float *up[2], *dn[2];                    // neighbours' ping-pong state buffers (null at the image top / bottom)
    long long up_plane, dn_plane;            // floats per plane in the neighbour's buffers
    long long up_off, dn_off;                // offset of the halo row we fill inside each of its planes
    long long first_off, last_off;           // offsets of our first / last owned row inside our planes
    int rank, nranks, row4, nplanes;         // row4 = float4 per row (P/4), nplanes = 9L
    unsigned long long gen;                  // generation of this qgmap_step call, already shifted
    long long timeout_cycles;
};

struct QgP2PBlob {                           // what a rank tells the others (QGMAP_P2P_BLOB_BYTES)
    int magic, pid, device, row_begin, row_end, g0, P, L, N, M;
    long long plane;
    void *raw_buf[2], *raw_box;              // valid inside the exporting process
    cudaIpcMemHandle_t ipc_buf[2], ipc_box;  // valid in other processes on the same node
};
static_assert(sizeof(QgP2PBlob) <= QGMAP_P2P_BLOB_BYTES, "blob too large");

struct QgP2P {
    QgMailbox *box = nullptr;                // own mailbox (cudaMalloc)
    QgP2PParams prm{};
    std::vector<void *> opened;              // cudaIpcOpenMemHandle mappings to close
    unsigned int gen = 0;
    bool connected = false;
};

__global__ void __launch_bounds__(256) qgmap_p2p_publish_kernel(const __grid_constant__ QgIterParams p, const __grid_constant__ QgP2PParams q)
{
    QgCtrl *c = p.ctrl;
    if (c->stop) return;
    const int it = c->it, wb = it & 1;                                  // the iteration kernel just wrote buffer wb
    const float *src = p.buf[wb];
    float *up = q.up[wb], *dn = q.dn[wb];
    const int total = q.nplanes * q.row4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int pl = i / q.row4, c4 = i - pl * q.row4;
        if (up) reinterpret_cast<float4 *>(up + pl * q.up_plane + q.up_off)[c4] =
                    __ldcg(reinterpret_cast<const float4 *>(src + pl * p.plane + q.first_off) + c4);
        if (dn) reinterpret_cast<float4 *>(dn + pl * q.dn_plane + q.dn_off)[c4] =
                    __ldcg(reinterpret_cast<const float4 *>(src + pl * p.plane + q.last_off) + c4);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int sh_last, sh_bad;
    QgMailbox *mine = q.box[q.rank];
    if (threadIdx.x == 0) {
        sh_last = (atomicAdd(&mine->ticket, 1u) == gridDim.x - 1);
        sh_bad = 0;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence_system();
    const int ns = p.L * QG_NRED;
    for (int k = threadIdx.x; k < q.nranks * ns; k += blockDim.x) {
        const int r = k / ns, j = k - r * ns;
        q.box[r]->sums[q.rank][j] = c->sums[j];
    }
    __threadfence_system();
    __syncthreads();
    const unsigned long long tag = q.gen | (unsigned long long)(unsigned int)it;
    if ((int)threadIdx.x < q.nranks) {
        *reinterpret_cast<volatile unsigned long long *>(&q.box[threadIdx.x]->flag[q.rank]) = tag;       // publish
        volatile unsigned long long *f = &mine->flag[threadIdx.x];                                          // wait for rank threadIdx.x
        const long long t0 = clock64();
        while (*f < tag) {
            if (clock64() - t0 > q.timeout_cycles) { sh_bad = 1; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0) {
        mine->ticket = 0;
        if (sh_bad) { c->comm_error = 1; c->stop = 1; return; }
        double tot[QG_LMAX * QG_NRED];
        for (int k = 0; k < ns; ++k) {
            double s = 0.0;
            for (int r = 0; r < q.nranks; ++r) s += *reinterpret_cast<volatile double *>(&mine->sums[r][k]);   // fixed rank order
            tot[k] = s;
        }
        qg_advance(p, c, tot);
    }
}

// First kernel of every qgmap_step call: "my state buffers are ready to be written into" (stream-ordered after this rank's
// set_state / previous step), then wait until every peer says the same.  Without it a fast rank could store its first boundary
// rows into a peer whose set_state has not imported its halo rows yet.  Host-level skew between ranks is allowed to be long.
__global__ void qgmap_p2p_ready_kernel(const __grid_constant__ QgIterParams p, const __grid_constant__ QgP2PParams q)
{
    QgCtrl *c = p.ctrl;
    __shared__ int sh_bad;
    if (threadIdx.x == 0) sh_bad = 0;
    __syncthreads();
    __threadfence_system();
    if ((int)threadIdx.x < q.nranks) {
        *reinterpret_cast<volatile unsigned long long *>(&q.box[threadIdx.x]->flag[q.rank]) = q.gen;        // tag (gen, iteration 0)
        volatile unsigned long long *f = &q.box[q.rank]->flag[threadIdx.x];
        const long long t0 = clock64();
        while (*f < q.gen) {
            if (clock64() - t0 > 12 * q.timeout_cycles) { sh_bad = 1; break; }
            __nanosleep(1000);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh_bad) { c->comm_error = 1; c->stop = 1; }
}

void qgmap_p2p_release(qgmap_handle *h)
{
    if (!h->p2p) return;
    for (void *m : h->p2p->opened) cudaIpcCloseMemHandle(m);
    if (h->p2p->box) cudaFree(h->p2p->box);
    delete h->p2p;
    h->p2p = nullptr;
}

extern "C" int qgmap_band_p2p_export(qgmap_handle *h, void *blob)
{
    if (!h || !blob) return QGMAP_ERR_ARG;
    QGP_CUDA(h, cudaSetDevice(h->device));
    qgmap_p2p_release(h);
    h->p2p = new QgP2P();
    QGP_CUDA(h, cudaMalloc(&h->p2p->box, sizeof(QgMailbox)));
    QGP_CUDA(h, cudaMemset(h->p2p->box, 0, sizeof(QgMailbox)));
    QgP2PBlob b;
    std::memset(&b, 0, sizeof b);
    b.magic = 0x51503270; b.pid = (int)getpid(); b.device = h->device;
    b.row_begin = h->row_begin; b.row_end = h->row_end; b.g0 = h->g0; b.P = h->P; b.L = h->L; b.N = h->N; b.M = h->M;
    b.plane = h->plane;
    b.raw_buf[0] = h->buf[0]; b.raw_buf[1] = h->buf[1]; b.raw_box = h->p2p->box;
    // IPC handles are only needed across processes; a failure here (e.g. IPC disabled in a container) is reported at connect
    // time, and only if a peer really lives in another process
    if (cudaIpcGetMemHandle(&b.ipc_buf[0], h->buf[0]) != cudaSuccess || cudaIpcGetMemHandle(&b.ipc_buf[1], h->buf[1]) != cudaSuccess ||
        cudaIpcGetMemHandle(&b.ipc_box, h->p2p->box) != cudaSuccess) {
        cudaGetLastError();
        b.magic = 0x51503271;                                   // "no IPC handles"
    }
    std::memset(blob, 0, QGMAP_P2P_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return QGMAP_OK;
}

static int map_peer(qgmap_handle *h, const QgP2PBlob &b, bool want_bufs, void **box, void **buf0, void **buf1)
{
    if (b.pid == (int)getpid()) {                              // same process: plain peer access
        if (b.device != h->device) {
            int can = 0;
            QGP_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, b.device));
            if (!can) QGP_FAIL(h, QGMAP_ERR_COMM, "device %d cannot access device %d (no P2P)", h->device, b.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) QGP_FAIL(h, QGMAP_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        *box = b.raw_box; *buf0 = b.raw_buf[0]; *buf1 = b.raw_buf[1];
        return QGMAP_OK;
    }
    if (b.magic != 0x51503270) QGP_FAIL(h, QGMAP_ERR_COMM, "peer rank exported no CUDA IPC handles");
    QGP_CUDA(h, cudaIpcOpenMemHandle(box, b.ipc_box, cudaIpcMemLazyEnablePeerAccess));
    h->p2p->opened.push_back(*box);
    if (want_bufs) {
        QGP_CUDA(h, cudaIpcOpenMemHandle(buf0, b.ipc_buf[0], cudaIpcMemLazyEnablePeerAccess));
        h->p2p->opened.push_back(*buf0);
        QGP_CUDA(h, cudaIpcOpenMemHandle(buf1, b.ipc_buf[1], cudaIpcMemLazyEnablePeerAccess));
        h->p2p->opened.push_back(*buf1);
    }
    return QGMAP_OK;
}

extern "C" int qgmap_band_p2p_connect(qgmap_handle *h, int rank, int nranks, const void *blobs)
{
    if (!h || !blobs || nranks < 1 || nranks > QG_RANKS_MAX || rank < 0 || rank >= nranks) return QGMAP_ERR_ARG;
    if (!h->p2p || !h->p2p->box) QGP_FAIL(h, QGMAP_ERR_STATE, "qgmap_band_p2p_connect before qgmap_band_p2p_export");
    QGP_CUDA(h, cudaSetDevice(h->device));
    std::vector<QgP2PBlob> bl(nranks);
    for (int r = 0; r < nranks; ++r) {
        std::memcpy(&bl[r], (const char *)blobs + (size_t)r * QGMAP_P2P_BLOB_BYTES, sizeof(QgP2PBlob));
        if ((bl[r].magic & ~1) != 0x51503270) QGP_FAIL(h, QGMAP_ERR_ARG, "blob %d is not a qgmap_band_p2p_export blob", r);
        if (bl[r].N != h->N || bl[r].L != h->L || bl[r].P != h->P || bl[r].M != h->M)
            QGP_FAIL(h, QGMAP_ERR_ARG, "rank %d solves a different problem (grid/components differ)", r);
        if (r > 0 && bl[r].row_begin != bl[r - 1].row_end) QGP_FAIL(h, QGMAP_ERR_ARG, "bands of ranks %d and %d are not adjacent", r - 1, r);
    }
    if (bl[rank].row_begin != h->row_begin || bl[rank].row_end != h->row_end) QGP_FAIL(h, QGMAP_ERR_ARG, "blob %d is not this handle's", rank);
    if (bl[0].row_begin != 0 || bl[nranks - 1].row_end != h->M) QGP_FAIL(h, QGMAP_ERR_ARG, "the bands do not cover the grid");
    QgP2PParams &q = h->p2p->prm;
    std::memset(&q, 0, sizeof q);
    q.rank = rank; q.nranks = nranks; q.row4 = h->P / 4; q.nplanes = F_COUNT * h->L;
    q.first_off = (long long)(h->row_begin - h->g0) * h->P;
    q.last_off = (long long)(h->row_end - 1 - h->g0) * h->P;
    int clk_khz = 2000000;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, h->device);
    q.timeout_cycles = (long long)clk_khz * 1000LL * 10LL;              // ~10 s
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { q.box[r] = h->p2p->box; continue; }
        const bool nb = (r == rank - 1 || r == rank + 1);
        void *box = nullptr, *b0 = nullptr, *b1 = nullptr;
        int rc = map_peer(h, bl[r], nb, &box, &b0, &b1);
        if (rc) return rc;
        q.box[r] = (QgMailbox *)box;
        if (r == rank - 1) {                                            // our first row -> its bottom halo row (global row row_end_up)
            q.up[0] = (float *)b0; q.up[1] = (float *)b1; q.up_plane = bl[r].plane;
            q.up_off = (long long)(bl[r].row_end - bl[r].g0) * bl[r].P;
        } else if (r == rank + 1) {                                     // our last row -> its top halo row (global row row_begin_dn - 1)
            q.dn[0] = (float *)b0; q.dn[1] = (float *)b1; q.dn_plane = bl[r].plane;
            q.dn_off = (long long)(bl[r].row_begin - 1 - bl[r].g0) * bl[r].P;
        }
    }
    h->rank = rank; h->nranks = nranks;
    h->params.band = nranks > 1 ? 1 : 0;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    h->p2p->connected = true;
    return QGMAP_OK;
}

void qgmap_p2p_begin_step(qgmap_handle *h)
{
    if (!h->p2p || !h->p2p->connected) return;
    h->p2p->prm.gen = (unsigned long long)(++h->p2p->gen) << 32;
    qgmap_p2p_ready_kernel<<<1, 32, 0, h->stream>>>(h->params, h->p2p->prm);
}

int qgmap_p2p_iteration(qgmap_handle *h, long long *launches)
{
    if (!h->p2p || !h->p2p->connected) QGP_FAIL(h, QGMAP_ERR_COMM, "band handle not connected (qgmap_band_p2p_connect)");
    qgmap_launch_iteration(h);
    const int total4 = h->p2p->prm.nplanes * h->p2p->prm.row4;
    const int nblk = std::max(1, std::min(32, (total4 + 1023) / 1024));
    qgmap_p2p_publish_kernel<<<nblk, 256, 0, h->stream>>>(h->params, h->p2p->prm);
    *launches += 1;
    return QGMAP_OK;
}
