// qgmap_peer.cuh -- device-side types of the peer-memory row-band exchange (host side: qgmap_p2p.cu).
//
// Every band (rank) owns one QgMailbox in its device memory and maps every other rank's mailbox and its two neighbours' state
// buffers (CUDA IPC across processes, plain peer access inside one process).  Per iteration a rank
//   1. stores its first / last owned row (9L planes of the buffer just written) into the neighbours' halo rows,
//   2. posts its 4L partial sums into slot [parity][rank] of EVERY mailbox, fences at system scope and raises
//      flag[rank] = (step generation << 32 | iteration) in every mailbox,
//   3. waits until all flags of its own mailbox carry this iteration, adds the partial sums in rank order (same bits on every
//      rank) and advances its control block (alpha update, anneal, stop test: gqmap_gpu_mixture.m:36,48,50,69-75).
// Hazards.  Iteration t reads buffer A and writes B; peers write only halo rows of B during t.  A peer can start t+1 (writing
// our halo rows of A) only after it has seen OUR flag for t, which is raised when every CTA of our iteration t has retired.
// The sums slots are double-buffered by iteration parity: a peer can overwrite slot [t&1] only in iteration t+2, whose flag wait
// needs our flag t+1, which we raise after we have read the sums of t.  Flags are monotone 64-bit tags: nothing is ever reset
// while a peer may still write.  A peer that never shows up must not hang the GPU: waits give up after ~10 s, stop the run and
// set QgCtrl::comm_error (-> QGMAP_ERR_COMM).
#pragma once
#include "qgmap_device.cuh"
#include "qgmap_advance.cuh"

#define QG_RANKS_MAX 16      // == QGMAP_P2P_RANKS_MAX (include/qgmap.h)

struct QgMailbox {
    double sums[2][QG_RANKS_MAX][QG_LMAX * QG_NRED];   // [iteration parity][writer rank]
    unsigned long long flag[QG_RANKS_MAX];             // (generation << 32 | iteration) rank q has completely published
    unsigned int ticket;                               // tiled kernel: CTAs of the publish kernel that finished their copy slice
};

struct QgPeer {
    QgMailbox *box[QG_RANKS_MAX];            // every rank's mailbox as mapped on this device (own included)
    float *up[2], *dn[2];                    // neighbours' ping-pong state buffers (null at the image top / bottom)
    long long up_plane, dn_plane;            // floats per plane in the neighbour's buffers
    long long up_off, dn_off;                // offset of the halo row we fill inside each of its planes
    long long first_off, last_off;           // offsets of our first / last owned row inside our planes
    int rank, nranks, row4, nplanes;         // row4 = float4 per row (P/4), nplanes = 9L
    int first_row, last_row;                 // global indices of those rows
    unsigned long long gen;                  // generation of the current qgmap_step call, already shifted (ready kernel)
    long long timeout_cycles;
};

// Steps 2 and 3 above, executed by ONE warp once this rank's partial sums `mine` (L x QG_NRED, shared memory) are complete and
// all boundary-row stores of the iteration have been ordered before this call (ticket chain + fences).
__device__ __forceinline__ void qg_peer_finish(const QgIterParams &p, QgCtrl *c, const QgPeer *q, const double *mine, int lane)
{
    const int it = c->it, par = it & 1, ns = p.L * QG_NRED, nr = q->nranks, me = q->rank;
    __threadfence_system();
    for (int k = lane; k < nr * ns; k += 32) {
        const int r = k / ns, j = k - r * ns;
        q->box[r]->sums[par][me][j] = mine[j];
    }
    __threadfence_system();
    __syncwarp();
    const unsigned long long tag = q->gen | (unsigned long long)(unsigned int)it;
    int bad = 0;
    QgMailbox *box = q->box[me];
    for (int r = lane; r < nr; r += 32) {
        *reinterpret_cast<volatile unsigned long long *>(&q->box[r]->flag[me]) = tag;                      // publish
        volatile unsigned long long *f = &box->flag[r];                                                     // wait for rank r
        const long long t0 = clock64();
        while (*f < tag) {
            if (clock64() - t0 > q->timeout_cycles) { bad = 1; break; }
            __nanosleep(100);
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    __threadfence_system();
    if (lane == 0) {
        if (bad) { c->comm_error = 1; c->stop = 1; return; }
        double tot[QG_LMAX * QG_NRED];
        for (int k = 0; k < ns; ++k) {
            double s = 0.0;
            for (int r = 0; r < nr; ++r) s += *reinterpret_cast<volatile double *>(&box->sums[par][r][k]);   // fixed rank order
            tot[k] = s;
        }
        qg_advance(p, c, tot);
    }
}
