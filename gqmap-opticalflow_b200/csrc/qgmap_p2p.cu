// qgmap_p2p.cu -- row bands of ONE frame pair over several GPUs with the halo exchange and the global sums done by our own
// kernels over NVLink peer memory (BASELINE configs[3]); no NCCL and no host on the per-iteration path.
// Protocol, hazards and the device-side types: qgmap_peer.cuh.
//
// Why: the NCCL transport (qgmap_band.cu) needs seven stream operations per iteration (all-reduce, advance kernel, two packs,
// grouped send/recv, two unpacks) -- ~0.13 ms per iteration at 8 GPUs against 0.58 ms of compute.  Two forms here:
//   * one thread per belief (qgmap_iter_kernel, qgmap_walk_kernel; QgIterParams::band == 2): the exchange is INSIDE the
//     iteration kernel -- the threads that update the band's first / last row store them straight into the neighbours' halo
//     rows (those tiles are scheduled first, so the stores overlap the rest of the band's compute), and the last CTA to retire
//     posts the partial sums, raises the flags, waits for the other bands and advances the control block (qg_peer_finish).
//     ONE launch per iteration, captured in the same 25-node CUDA graph as the single-GPU loop.
//   * four lanes per belief (qgmap_iter_kernel_g4, the super-pixel default): qgmap_p2p_publish_kernel follows the iteration
//     kernel: all its CTAs copy the boundary rows, its last CTA runs qg_peer_finish -- two launches per iteration.
// Every qgmap_step call starts with a ready handshake (qgmap_p2p_ready_kernel) so that set_state on one rank can never race
// with a faster neighbour's first boundary rows.
#include "qgmap_internal.h"
#include "qgmap_peer.cuh"
#include <unistd.h>
#include <algorithm>
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

static_assert(QG_RANKS_MAX == QGMAP_P2P_RANKS_MAX, "rank limit of the header and of the mailbox differ");

struct QgP2PBlob {                           // what a rank tells the others (QGMAP_P2P_BLOB_BYTES)
    int magic, pid, device, row_begin, row_end, g0, P, L, N, M;
    long long plane;
    void *raw_buf[2], *raw_box;              // valid inside the exporting process
    cudaIpcMemHandle_t ipc_buf[2], ipc_box;  // valid in other processes on the same node
};
static_assert(sizeof(QgP2PBlob) <= QGMAP_P2P_BLOB_BYTES, "blob too large");

struct QgP2P {
    QgMailbox *box = nullptr;                // own mailbox (cudaMalloc)
    QgPeer prm{};                            // host copy
    QgPeer *d_peer = nullptr;                // device copy read by the kernels (gen is written by the ready kernel)
    std::vector<void *> opened;              // cudaIpcOpenMemHandle mappings to close
    unsigned int gen = 0;
    bool connected = false;
};

// Tiled iteration kernel only: copy the boundary rows the iteration kernel just wrote into the neighbours' halo rows; the last
// CTA to finish exchanges the sums and advances the control block.
__global__ void __launch_bounds__(256) qgmap_p2p_publish_kernel(const __grid_constant__ QgIterParams p, const QgPeer *__restrict__ q)
{
    QgCtrl *c = p.ctrl;
    if (c->stop) return;
    const int it = c->it, wb = it & 1;                                  // the iteration kernel just wrote buffer wb
    const float *src = p.buf[wb];
    float *up = q->up[wb], *dn = q->dn[wb];
    const int row4 = q->row4, total = q->nplanes * row4;
    const long long up_plane = q->up_plane, dn_plane = q->dn_plane, up_off = q->up_off, dn_off = q->dn_off;
    const long long first_off = q->first_off, last_off = q->last_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int pl = i / row4, c4 = i - pl * row4;
        if (up) reinterpret_cast<float4 *>(up + pl * up_plane + up_off)[c4] =
                    __ldcg(reinterpret_cast<const float4 *>(src + pl * p.plane + first_off) + c4);
        if (dn) reinterpret_cast<float4 *>(dn + pl * dn_plane + dn_off)[c4] =
                    __ldcg(reinterpret_cast<const float4 *>(src + pl * p.plane + last_off) + c4);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int sh_last;
    QgMailbox *mine = q->box[q->rank];
    if (threadIdx.x == 0) sh_last = (atomicAdd(&mine->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!sh_last || threadIdx.x >= 32) return;
    if (threadIdx.x == 0) mine->ticket = 0;
    __shared__ double sh_mine[QG_LMAX * QG_NRED];
    for (int k = threadIdx.x; k < p.L * QG_NRED; k += 32) sh_mine[k] = c->sums[k];
    __syncwarp();
    qg_peer_finish(p, c, q, sh_mine, threadIdx.x);
}

// First kernel of every qgmap_step call: "my state buffers are ready to be written into" (stream-ordered after this rank's
// set_state / previous step), then wait until every peer says the same.  Without it a fast rank could store its first boundary
// rows into a peer whose set_state has not imported its halo rows yet.  Host-level skew between ranks is allowed to be long.
__global__ void qgmap_p2p_ready_kernel(const __grid_constant__ QgIterParams p, QgPeer *q, unsigned long long gen)
{
    QgCtrl *c = p.ctrl;
    __shared__ int sh_bad;
    if (threadIdx.x == 0) { sh_bad = 0; q->gen = gen; }
    __syncthreads();
    __threadfence_system();
    if ((int)threadIdx.x < q->nranks) {
        *reinterpret_cast<volatile unsigned long long *>(&q->box[threadIdx.x]->flag[q->rank]) = gen;          // tag (gen, iteration 0)
        volatile unsigned long long *f = &q->box[q->rank]->flag[threadIdx.x];
        const long long t0 = clock64();
        while (*f < gen) {
            if (clock64() - t0 > 12 * q->timeout_cycles) { sh_bad = 1; break; }
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
    if (h->p2p->d_peer) cudaFree(h->p2p->d_peer);
    delete h->p2p;
    h->p2p = nullptr;
}

extern "C" int qgmap_band_p2p_export(qgmap_handle *h, void *blob)
{
    return qg_guard([&]() -> int {
    if (!h || !blob) return QGMAP_ERR_ARG;
    QGP_CUDA(h, cudaSetDevice(h->device));
    qgmap_p2p_release(h);
    h->p2p = new QgP2P();
    QGP_CUDA(h, cudaMalloc(&h->p2p->box, sizeof(QgMailbox)));
    QGP_CUDA(h, cudaMemset(h->p2p->box, 0, sizeof(QgMailbox)));
    QGP_CUDA(h, cudaMalloc(&h->p2p->d_peer, sizeof(QgPeer)));
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
    });
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
    return qg_guard([&]() -> int {
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
    QgPeer &q = h->p2p->prm;
    std::memset(&q, 0, sizeof q);
    q.rank = rank; q.nranks = nranks; q.row4 = h->P / 4; q.nplanes = F_COUNT * h->L;
    q.first_off = (long long)(h->row_begin - h->g0) * h->P;
    q.last_off = (long long)(h->row_end - 1 - h->g0) * h->P;
    q.first_row = h->row_begin; q.last_row = h->row_end - 1;
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
    QGP_CUDA(h, cudaMemcpy(h->p2p->d_peer, &q, sizeof q, cudaMemcpyHostToDevice));
    h->rank = rank; h->nranks = nranks;
    // one thread per belief (tiled or row-walking kernel): the iteration kernel carries the exchange itself (band 2); the four-lane
    // kernel of the super-pixel variant leaves its sums for the publish kernel (band 1)
    h->params.band = nranks > 1 ? (h->lanes_per_belief == 1 ? 2 : 1) : 0;
    h->params.peer = h->p2p->d_peer;
    h->params.pub_row[0] = q.up[0] ? h->row_begin : -1;
    h->params.pub_row[1] = q.dn[0] ? h->row_end - 1 : -1;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    h->p2p->connected = true;
    return QGMAP_OK;
    });
}

void qgmap_p2p_begin_step(qgmap_handle *h)
{
    if (!h->p2p || !h->p2p->connected) return;
    const unsigned long long gen = (unsigned long long)(++h->p2p->gen) << 32;
    qgmap_p2p_ready_kernel<<<1, 32, 0, h->stream>>>(h->params, h->p2p->d_peer, gen);
}

bool qgmap_p2p_fused(const qgmap_handle *h) { return h->p2p && h->p2p->connected && h->params.band == 2; }

int qgmap_p2p_iteration(qgmap_handle *h, long long *launches)
{
    if (!h->p2p || !h->p2p->connected) QGP_FAIL(h, QGMAP_ERR_COMM, "band handle not connected (qgmap_band_p2p_connect)");
    qgmap_launch_iteration(h);
    if (h->params.band == 2) return QGMAP_OK;                  // row-walking kernel: the exchange is inside the iteration kernel
    const int total4 = h->p2p->prm.nplanes * h->p2p->prm.row4;
    const int nblk = std::max(1, std::min(32, (total4 + 1023) / 1024));
    qgmap_p2p_publish_kernel<<<nblk, 256, 0, h->stream>>>(h->params, h->p2p->d_peer);
    *launches += 1;
    return QGMAP_OK;
}
