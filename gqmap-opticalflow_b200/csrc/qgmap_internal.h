// qgmap_internal.h -- private handle layout shared by the translation units of libqgmap.so.
#pragma once
#include "../../include/qgmap.h"
#include "qgmap_device.cuh"
#include <cuda_runtime.h>
#include <string>

struct QgMonParams;
struct QgBand;
struct QgP2P;

struct qgmap_handle {
    qgmap_config cfg{};
    int Mo = 0, No = 0, M = 0, N = 0, L = 0, K = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evb0 = nullptr, evb1 = nullptr;
    bool pending = false;      // qgmap_step_begin without qgmap_step_end
    int it0 = 0;               // iteration counter when the pending step began
    // band geometry: rows [row_begin,row_end) owned; [g0,g1) stored; [out_r0,out_r1) updated
    int row_begin = 0, row_end = 0, g0 = 0, g1 = 0, rows_local = 0, out_r0 = 0, out_r1 = 0;
    int P = 0;
    long long plane = 0;
    float *I1f = nullptr;
    QgTap8 *VVf = nullptr;     // packed padded second frame (see QgIterParams::VV8)
    QgTap16h *VVh = nullptr;   // fp16 4x4-block layout of the same frame; null unless the frame is fp16-exact (QgIterParams::VVh)
    int pitch4 = 0;
    double *I1d = nullptr, *VVd = nullptr;
    int pitchI = 0, pitchV = 0;
    float *buf[2] = {nullptr, nullptr};
    float *dbg = nullptr;
    QgCtrl *ctrl = nullptr, *ctrl_host = nullptr;
    double *partials = nullptr, *gpartials = nullptr;
    unsigned int *tickets = nullptr;
    bool pdl = false;             // iteration launches carry the programmatic-serialization attribute (QGMAP_PDL; qg_pdl_enter)
    bool walk = false;            // full-resolution variant: row-walking kernel (qgmap_walk.cuh) instead of the tiled one
    double *hist[3] = {nullptr, nullptr, nullptr};
    int hist_cap = 0;
    double *stage = nullptr;
    size_t stage_cap = 0;
    double *mon_partials = nullptr, *d_map = nullptr, *d_tflow = nullptr;
    unsigned char *d_unknown = nullptr;
    bool has_unknown = false, has_state = false;
    QgIterParams params{};
    dim3 grid{};
    int lanes_per_belief = 1;     // 1: qgmap_iter_kernel, 4: qgmap_iter_kernel_g4
    cudaGraphExec_t graph = nullptr;
    float last_ms = 0.f;
    long long last_launches = 0;
    int rank = 0, nranks = 1;
    bool in_group = false;         // member of a single-process qgmap_group (stepped by the group only)
    int band_steps_enqueued = 0;   // NCCL band mode: iterations enqueued since the last control-block read-back
    QgBand *band = nullptr;
    QgP2P *p2p = nullptr;          // band mode over peer memory (qgmap_band_p2p_connect / multi-device qgmap_group)
    std::string err;
};

// qgmap_map.cu (compiled with -fmad=false: fp64 monitoring arithmetic must not be contracted)
void qgmap_launch_find_map_f32(const double *alpha, const float *mu_u, const float *sig_u, const float *mu_v,
                               const float *sig_v, long long comp_stride, int M, int N, int L, int pitch, int row_off,
                               int r0, int r1, double *map, long long total, cudaStream_t s);
void qgmap_launch_find_map_f64(const double *alpha, const double *mu_u, const double *sig_u, const double *mu_v,
                               const double *sig_v, long long comp_stride, int M, int N, int L, double *map,
                               long long total, cudaStream_t s);
struct QgMonArgs {
    const double *I1; int pitchI; const double *VV; int pitchV;
    int Mo, No, M, N, super; double lambdad, lambdas, epsn;
    int r0, r1;                    // belief rows [r0,r1) summed (a band's share)
};
void qgmap_launch_logp(const QgMonArgs &q, const double *uv, double *partials, int nblk, cudaStream_t s);
void qgmap_launch_aepe(const QgMonArgs &q, const double *map, const double *tflow, const unsigned char *unknown,
                       double *partials, int nblk, cudaStream_t s);

// qgmap_band.cu: row-band decomposition over NCCL
void qgmap_band_release(qgmap_handle *h);
int qgmap_band_refresh(qgmap_handle *h);                       // exchange halo rows of the current state
int qgmap_band_iteration(qgmap_handle *h, long long *launches); // one iteration incl. all-reduce + halo exchange
void qgmap_launch_iteration(const qgmap_handle *h);             // plain iteration kernel launch on h->stream
void qgmap_launch_advance(const qgmap_handle *h);
void qgmap_p2p_release(qgmap_handle *h);
void qgmap_p2p_begin_step(qgmap_handle *h);                     // new generation tag for the flags of this qgmap_step call
int qgmap_p2p_iteration(qgmap_handle *h, long long *launches);  // iteration kernel (+ publish/advance kernel for the tiled form)
bool qgmap_p2p_fused(const qgmap_handle *h);                    // exchange inside the iteration kernel (row-walking form)
int qgmap_enqueue_iterations(qgmap_handle *h, int n, long long *launches);   // n iterations of a single domain or a p2p band
int qgmap_prepare_step(qgmap_handle *h, int n, int its);
int qgmap_finish_step(qgmap_handle *h, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped);
void qgmap_set_last_error(const char *msg);
#include <vector>
void qgmap_random_state(const qgmap_config &c, size_t n, int L, uint64_t seed, std::vector<double> &w, std::vector<double> &muu,
                        std::vector<double> &muv, std::vector<double> &sigu, std::vector<double> &sigv);
struct qgmap_group;
qgmap_handle *qgmap_group_band(qgmap_group *g, int b);           // band b's handle (monitoring kernels of qgmap_group_solve)

#include "qgmap_guard.h"

extern thread_local long long g_solve_launches;
extern thread_local float g_solve_ms;
