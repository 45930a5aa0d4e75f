/*
 * qgmap.h -- C ABI of libqgmap.so: the B200-native (sm_100a CUDA) QGMAP optical-flow inference path.
 *
 * This is the drop-in boundary for ONE hot path of motionlife/gqmap-opticalflow: the per-pixel
 * Gauss-Hermite-quadrature MAP gradient-ascent loop.  The reference has no source-level plug-in API;
 * its de-facto operator boundary is (i) the MATLAB solver signature and (ii) two MEX call signatures.
 * Each entry point below names the reference interface (file:line, relative to the reference root)
 * it replaces.  The MEX gateways (gqmap-opticalflow_b200/mex/) and the ctypes host (host.py) both bind
 * exactly these symbols; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every array crossing the boundary is a HOST buffer in MATLAB layout:
 *    column-major IEEE fp64 (uint8 where stated); element (m,n,l) of an M x N x L array (0-based) is at
 *    m + M*n + M*N*l.  rou is M x N x L x 2 x 2 (edge e: 0=down,1=right; layer c: 0=u,1=v), as
 *    gqmap_gpu_mixture.m:24 allocates it.  Device layout is private to the library.
 *  - Every function returns 0 (QGMAP_OK) or a negative qgmap_status; no exceptions cross the boundary;
 *    no global state except per-handle.  One host thread per handle at a time.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with QGMAP_ERR_CUDA.
 *  - "belief grid" = M x N = image size Mo x No (variant FULL) or Mo/4 x No/4 (variant SUPER).
 */
#ifndef QGMAP_H
#define QGMAP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGMAP_VERSION 100
#define QGMAP_LMAX 10      /* get_map_mex.mexw64 uses 80-byte (10 double) stack buffers per pixel        */
#define QGMAP_KMAX 32      /* quadrature order per axis                                                   */

typedef enum {
    QGMAP_OK = 0,
    QGMAP_ERR_ARG = -1,       /* bad argument (NULL, size, range)                                          */
    QGMAP_ERR_CUDA = -2,      /* CUDA runtime error or no device; qgmap_last_error() has the text          */
    QGMAP_ERR_STATE = -3,     /* call order (e.g. step before set_state/init_state)                        */
    QGMAP_ERR_NOMEM = -4,
    QGMAP_ERR_COMM = -5       /* NCCL / band-exchange error                                                */
} qgmap_status;

enum { QGMAP_VARIANT_FULL = 0,      /* gqmap_gpu_mixture.m          */
       QGMAP_VARIANT_SUPER = 1 };   /* gqmap_gpuSuper_mix_entropy.m */
enum { QGMAP_ALPHA_SOFTMAX = 0,     /* updateAlpha, gqmap_gpu_mixture.m:78-86 (live code)                 */
       QGMAP_ALPHA_PROJSPLX = 1 };  /* projsplx.m:15-32 via the commented line gqmap_gpu_mixture.m:49     */

/* The `options` struct of the reference (gqmap_gpu_mixture.m:3-6) plus the constants the reference
 * hard-codes, exposed as fields.  qgmap_config_defaults() fills the reference's values. */
typedef struct {
    int32_t struct_size;            /* = sizeof(qgmap_config), ABI check                                   */
    int32_t variant;                /* QGMAP_VARIANT_*                                                     */
    int32_t L;                      /* options.L  mixture components                                       */
    int32_t K;                      /* options.K  Gauss-Hermite order per axis (K*K points)                */
    double lambdad, lambdas, epsn;  /* options.lambdad / lambdas / epsn                                    */
    double temperature, drate;      /* options.temperature / drate                                         */
    double minu, maxu, minv, maxv;  /* options.minu..maxv  (clamp range of the means)                      */
    double sigma_min, sigma_max;    /* 0.01 ; 23 (gqmap_gpu_mixture.m:43) / 25 (gqmap_gpuSuper..m:42)       */
    double corr_tor;                /* 1-1e-5   (:7)                                                       */
    double step0, step_tau;         /* step = step0/(1+it/step_tau): 0.1,8000 (:27) / 0.001,4000 (S:26)    */
    double alpha_scale;             /* 1e-7 (:83)                                                          */
    double T_floor;                 /* 0.001 (S:72)                                                        */
    double tor;                     /* 1e-4 (:25) stop when mean|dmu_u| < tor                              */
    double sigma_step_scale;        /* 1 (:43-44); legacy/gqmap_ctf.m:34-35 steps sigma with step*0.3.  NOTE: run with
                                       gqmap_ctf's constants the library still samples the second frame with the exact
                                       bicubic of :156-179, NOT gqmap_ctf's nearest lookup into a 64x upsampled frame
                                       (legacy/gqmap_ctf.m:10,96) -- a deliberate difference, measured in
                                       tests/test_refsrc_ctf.py                                            */
    int32_t alpha_start;            /* alpha updates when it > 500 (:50)                                   */
    int32_t alpha_mode;             /* QGMAP_ALPHA_*                                                       */
    int32_t anneal_every;           /* 0 = never (full-res, :73 commented) ; 500 (S:72)                    */
    int32_t device;                 /* CUDA device ordinal, -1 = current device                            */
    int32_t row_begin, row_end;     /* rows [row_begin,row_end) of the belief grid owned by this handle
                                       (row-band decomposition); 0,0 = whole grid                          */
    int32_t log_every;              /* monitoring cadence of qgmap_solve: 300 (:52)                        */
    int32_t strip_rows;             /* full-resolution kernel: rows one warp walks; 0 = chosen from the grid size */
} qgmap_config;

typedef struct qgmap_handle qgmap_handle;

/* Fill *cfg with the reference's hard-coded constants for the given variant (L=1,K=3 placeholders). */
int qgmap_config_defaults(qgmap_config *cfg, int variant);

/* Setup, gqmap_gpu_mixture.m:3-14 / gqmap_gpuSuper_mix_entropy.m:3-14: upload I1, build VV=getVV(I2)
 * (:191-208), Gauss-Hermite tables (GaussHermite_2.m:21-32), allocate state.  I1,I2: Mo x No. */
int qgmap_create(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, qgmap_handle **out);
int qgmap_destroy(qgmap_handle *h);

/* Belief-grid dimensions of a handle; size of the frames it was created with. */
int qgmap_dims(const qgmap_handle *h, int *M, int *N, int *L);
int qgmap_image_dims(const qgmap_handle *h, int *Mo, int *No);

/* State in/out (the arrays of gqmap_gpu_mixture.m:18-24).  The reference hard-wires a random init and
 * cannot take one; these calls are what make runs reproducible and double as checkpoint/resume.
 * muu,muv,sigu,sigv,pn: M x N x L; rou: M x N x L x 2 x 2; w: L.  alpha = softmax(w) unless alpha!=NULL.
 * set_state resets the iteration counter to `it` (1 = fresh run) and the temperature to T. */
int qgmap_set_state(qgmap_handle *h, const double *muu, const double *muv, const double *sigu, const double *sigv,
                    const double *pn, const double *rou, const double *w, const double *alpha, double T, int it);
int qgmap_get_state(qgmap_handle *h, double *muu, double *muv, double *sigu, double *sigv,
                    double *pn, double *rou, double *w, double *alpha, double *T, int *it);
/* The same with single-precision belief arrays (MATLAB `single`; w, alpha, T stay double).  The device keeps the beliefs in fp32, so
 * this boundary format is lossless -- get_state_f32 -> set_state_f32 restores the state bit for bit -- at half the host<->device
 * bytes (2160 x 3840, L=3: 0.9 GB instead of 1.8 GB per direction), which is what a short call on a large frame pays for. */
int qgmap_set_state_f32(qgmap_handle *h, const float *muu, const float *muv, const float *sigu, const float *sigv,
                        const float *pn, const float *rou, const double *w, const double *alpha, double T, int it);
int qgmap_get_state_f32(qgmap_handle *h, float *muu, float *muv, float *sigu, float *sigv,
                        float *pn, float *rou, double *w, double *alpha, double *T, int *it);
/* Random init exactly as :18-24 (uniform draws; generator = splitmix64-seeded xoshiro256**, not MATLAB's). */
int qgmap_init_state(qgmap_handle *h, uint64_t seed);

/* The loop body gqmap_gpu_mixture.m:26-50,69-75 (without the monitoring block :52-68), run for at most
 * n iterations or until the reference would `break` (:75: it>its or mean|dmu_u|<tor).
 * energy/ptdmu/ptdsigma: n doubles each (may be NULL): Energy(it) (:48), mean|dmuu| and mean|dsigmau| (:69-70)
 * of each executed iteration.  *n_done iterations executed, *stopped = 1 if the break condition fired. */
int qgmap_step(qgmap_handle *h, int n, int its, double *energy, double *ptdmu, double *ptdsigma,
               int *n_done, int *stopped);
/* The same split in two so that independent handles overlap: begin enqueues (no wait), end waits and fetches. */
int qgmap_step_begin(qgmap_handle *h, int n, int its);
int qgmap_step_end(qgmap_handle *h, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped);
/* Independent frame pairs (BASELINE config "batch of frame pairs"): n iterations on each of nh handles of ONE device,
 * concurrently on their own streams; *device_ms = device time of the whole batch (CUDA events), *launches = kernels. */
int qgmap_batch_step(qgmap_handle **hs, int nh, int n, int its, float *device_ms, long long *launches);
/* Measured FP32 FMA throughput (TFLOP/s) of `device`: the roofline denominator for the FP32-bound iteration kernel. */
int qgmap_fp32_peak(int device, double *tflops);

/* Device time (CUDA events on the handle's stream) of the kernels launched by the last qgmap_step, in ms. */
int qgmap_last_step_ms(const qgmap_handle *h, float *ms);
/* Number of kernels the last qgmap_step / qgmap_solve launched (for the bench's gpu_launches claim). */
int qgmap_last_launches(const qgmap_handle *h, long long *launches);

/* MAP flow of the current beliefs: gqmap_gpu_mixture.m:53-58 (L==1: cat(3,mu_u,mu_v); else get_map_mex).
 * map: M x N x 2. */
int qgmap_get_map(qgmap_handle *h, double *map);   /* a row-band handle fills only the rows it owns */
/* profile_logP(map), gqmap_gpu_mixture.m:148-154 / gqmap_gpuSuper_mix_entropy.m:152-169.  map: M x N x 2. */
int qgmap_logp(qgmap_handle *h, const double *map, double *lp);
/* AEPE of a map against ground truth, gqmap_gpu_mixture.m:63-64 / gqmap_gpuSuper_mix_entropy.m:58-63.
 * map M x N x 2 (belief grid), tflow Mo x No x 2, unknown Mo x No uint8 (may be NULL). */
int qgmap_aepe(qgmap_handle *h, const double *map, const double *tflow, const uint8_t *unknown, double *aepe);
/* The same monitoring block for the rows a handle OWNS (row bands; gqmap_gpu_mixture.m:52-67), computed where the rows live:
 * qgmap_set_truth uploads options.trueFlow / options.unknownIdx once; qgmap_monitor_partial extracts the MAP of the handle's
 * stored rows on its device and returns its share of profile_logP (:148-154) and of the AEPE numerator (:63-64; divide the sum of
 * all shares by (Mo-2b)(No-2b), b = 1, or 4 for the super-pixel variant).  Either output may be NULL. */
int qgmap_set_truth(qgmap_handle *h, const double *tflow, const uint8_t *unknown);
int qgmap_monitor_partial(qgmap_handle *h, double *logp_share, double *aepe_sum_share);

/* One-call solver == [mu,sigma,alpha,AEPE,Energy,logP] = gqmap_gpu_mixture(options,I1,I2)
 * (gqmap_gpu_mixture.m:1) or gqmap_gpuSuper_mix_entropy (same signature), all host buffers.
 * its = options.its.  init (optional, may be NULL -> qgmap_init_state(seed)): the 7 state arrays in the order
 * muu,muv,sigu,sigv,pn,rou,w.  tflow Mo x No x 2 and unknown Mo x No (options.trueFlow/unknownIdx) may be NULL
 * (then AEPE stays NaN).  Outputs: mu,sigma M x N x L x 2 (:183-184); alpha L; AEPE,Energy,logP: its doubles,
 * prefilled NaN / 0 / NaN as :16.  *its_done = iterations executed. */
int qgmap_solve(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int its,
                const double *const *init, uint64_t seed, const double *tflow, const uint8_t *unknown,
                double *mu, double *sigma, double *alpha, double *AEPE, double *Energy, double *logP,
                int *its_done);

/* The same one-call solver with the frame pair split into `nbands` row bands, band b on devices[b] (options.devices of the
 * MATLAB-level call; NULL: all bands on the current device): qgmap_group_* underneath -- peer-memory exchange kernel when
 * every band has its own GPU -- and the reference's monitoring cadence (MAP of the gathered beliefs via qgmap_find_map,
 * AEPE / logP on band 0's device, PNG dumps).  Beliefs are bit-identical to qgmap_solve's. */
int qgmap_group_solve(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int its,
                      int nbands, const int *devices,
                      const double *const *init, uint64_t seed, const double *tflow, const uint8_t *unknown,
                      double *mu, double *sigma, double *alpha, double *AEPE, double *Energy, double *logP,
                      int *its_done);

/* Kernels launched / device ms (CUDA events) of the last qgmap_solve on this thread. */
int qgmap_last_solve_stats(long long *launches, float *kernel_ms);

/* Stateless get_map_mex replacement: map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v)
 * (call sites gqmap_gpu_mixture.m:57, gqmap_gpuSuper_mix_entropy.m:56; algorithm legacy/findMixMax.m:1-38 +
 * MATLAB R2018b fminbnd).  alpha: L; mu_*,sig_*: M x N x L; map: M x N x 2.  Runs on `device` (-1 current). */
int qgmap_find_map(const double *alpha, const double *mu_u, const double *sig_u, const double *mu_v,
                   const double *sig_v, int M, int N, int L, double *map, int device);

/* flowToColor_mex replacement (host C++): [img,flo,minu,maxu,minv,maxv,idxUnknown] = flowToColor(flow[,maxFlow])
 * (legacy/flowToColor.m:1,37-87 + legacy/computeColor.m:33-115; call sites optical_flow.m:12-13).
 * flow M x N x 2; img M x N x 3 uint8; flo M x N x 2; stats[4]=minu,maxu,minv,maxv; unknown M x N uint8.
 * max_flow <= 0: normalise by the largest flow present. */
int qgmap_flow_to_color(const double *flow, int M, int N, double max_flow,
                        uint8_t *img, double *flo, double *stats, uint8_t *unknown);

/* imwrite(flc, [options.dir '/' num2str(it) '.png']) replacement (gqmap_gpu_mixture.m:62, gqmap_gpuSuper_mix_entropy.m:61):
 * 8-bit RGB PNG of a column-major M x N x 3 uint8 image (what qgmap_flow_to_color returns).  Host C++, no zlib needed
 * (stored deflate blocks). */
int qgmap_write_png(const char *path, const uint8_t *rgb, int M, int N);
/* options.dir of the one-call solver (gqmap_gpu_mixture.m:62): while set (per host thread; NULL or "" clears it), qgmap_solve
 * writes <dir>/<it>.png -- the colour-coded MAP flow, for the super-pixel variant repelem(map,4,4) cropped 5:end-4 (S:58-61) --
 * at it == 1 and every log_every iterations.  The directory must exist (the reference's drivers mkdir it, optical_flow.m:25). */
int qgmap_solve_set_dump_dir(const char *dir);
/* readFlowFile.m:33-81 / legacy/writeFlowFile.m: Middlebury .flo <-> column-major H x W x 2 doubles.
 * qgmap_read_flo: call with flow == NULL to get the size, then with a buffer of H*W*2 doubles. */
int qgmap_read_flo(const char *path, int *H, int *W, double *flow);
int qgmap_write_flo(const char *path, const double *flow, int H, int W);

/* Gauss-Hermite nodes/weights, GaussHermite_2.m:21-32 ([x,w] = GaussHermite_2(n)), n <= QGMAP_KMAX. */
int qgmap_gauss_hermite(int n, double *x, double *w);
/* projsplx.m:15-32: Euclidean projection of y (length m) onto the probability simplex. */
int qgmap_projsplx(const double *y, int m, double *x);

/* Debug/parity: run ONE gradient pass on the current state WITHOUT updating it and return the assembled
 * per-pixel quantities of gqmap_gpu_mixture.m:36-40 (fp32 results widened to fp64):
 * G_muu,G_muv,G_sigu,G_sigv,dpn: M x N x L; drou: M x N x L x 2 x 2; e_px = nEnergy+sum eEnergy and
 * da_px = dan+sum dae: M x N x L.  Border entries are 0.  Any pointer may be NULL. */
int qgmap_debug_gradients(qgmap_handle *h, double *G_muu, double *G_muv, double *G_sigu, double *G_sigv,
                          double *dpn, double *drou, double *e_px, double *da_px);

/* Row-band decomposition of ONE frame pair over several GPUs (one handle per GPU/process).  The handle must
 * have been created with row_begin/row_end.  `nccl_unique_id` is the 128-byte ncclUniqueId produced by
 * qgmap_band_unique_id() on rank 0 and distributed by the host (torch.distributed / MATLAB parpool / files). */
int qgmap_band_unique_id(void *id128);
int qgmap_band_connect(qgmap_handle *h, int rank, int nranks, const void *nccl_unique_id);
/* The same decomposition WITHOUT NCCL on the data path: every rank maps its two neighbours' state buffers and every rank's
 * small mailbox (CUDA IPC across processes, plain peer access inside one process).  Per iteration ONE extra kernel stores the
 * band's boundary rows straight into the neighbours' halo rows over NVLink, posts the band's 4L partial sums into every
 * rank's mailbox, raises a flag, waits for the other ranks' flags, sums in fixed rank order (bit-identical on all ranks) and
 * advances the control block.  No host involvement, two launches per iteration.  Protocol: every rank calls
 * qgmap_band_p2p_export (fills QGMAP_P2P_BLOB_BYTES), the host all-gathers the blobs (torch.distributed / files), every rank
 * calls qgmap_band_p2p_connect with the nranks blobs in rank order, then a host barrier before the first step.  All ranks must
 * issue the same qgmap_step calls.  A rank that does not hear from a peer within ~10 s (~2 min for the handshake that opens every qgmap_step call) stops with
 * QGMAP_ERR_COMM; after a re-export every rank must reconnect. */
#define QGMAP_P2P_BLOB_BYTES 512
#define QGMAP_P2P_RANKS_MAX 16
int qgmap_band_p2p_export(qgmap_handle *h, void *blob);
int qgmap_band_p2p_connect(qgmap_handle *h, int rank, int nranks, const void *blobs);

/* The same decomposition driven by ONE process (the natural model for a MATLAB host; also how it is tested on one GPU):
 * nbands handles, band b on devices[b] (NULL: all on the current device), halo rows copied peer-to-peer between the bands'
 * state buffers, global sums combined in fixed band order.  State calls take / return the FULL-grid arrays of
 * qgmap_set_state / qgmap_get_state; qgmap_group_step has the semantics of qgmap_step. */
typedef struct qgmap_group qgmap_group;
int qgmap_group_create(const qgmap_config *cfg, const double *I1, const double *I2, int Mo, int No, int nbands,
                       const int *devices, qgmap_group **out);
int qgmap_group_destroy(qgmap_group *g);
int qgmap_group_dims(const qgmap_group *g, int *M, int *N, int *L, int *nbands);
int qgmap_group_set_state(qgmap_group *g, const double *muu, const double *muv, const double *sigu, const double *sigv,
                          const double *pn, const double *rou, const double *w, const double *alpha, double T, int it);
int qgmap_group_init_state(qgmap_group *g, uint64_t seed);
int qgmap_group_get_state(qgmap_group *g, double *muu, double *muv, double *sigu, double *sigv, double *pn, double *rou,
                          double *w, double *alpha, double *T, int *it);
int qgmap_group_step(qgmap_group *g, int n, int its, double *energy, double *ptdmu, double *ptdsigma, int *n_done, int *stopped);
int qgmap_group_last_step_ms(const qgmap_group *g, float *ms);

const char *qgmap_last_error(const qgmap_handle *h);   /* h may be NULL: last error of a handle-less call */
const char *qgmap_status_string(int status);
int qgmap_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QGMAP_H */
