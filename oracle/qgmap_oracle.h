/*
 * qgmap_oracle.h -- CPU fp64 restatement of the QGMAP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libqgmap.so, the CUDA path) never links or calls it.
 *
 * PINNED against the reference, executed:
 *   - qo_find_map (get_map_mex) and qo_flow_to_color (flowToColor_mex) against the shipped .mexw64 binaries, run natively by
 *     oracle/refbin/ -- bit-identical (tests/test_refbin_get_map.py, tests/test_refbin_flow_to_color.py);
 *   - the iteration loop (qo_run, qo_gradients, qo_get_vv, qo_gauss_hermite, qo_projsplx, qo_profile_logp, qo_aepe) against the
 *     reference's own .m files, run unmodified by the mini-MATLAB interpreter oracle/mlab/minimat.py -- fp64 rounding
 *     (tests/test_refsrc_parity.py, vectors tests/golden/refsrc_*.npz).  The interpreter is ours, MATLAB itself is not
 *     available: MATLAB's libm, summation order and LAPACK stay unpinned (all at rounding level).
 * It is additionally held in place by (i) closed-form known answers derived from the maths (tests/test_oracle_*.py), and
 * (ii) an independently written NumPy twin (oracle/numpy_twin.py) that must agree to fp64 rounding.
 *
 * All arrays are MATLAB column-major fp64: element (m,n,l) of an M x N x L array (1-based) lives at
 * (m-1) + M*(n-1) + M*N*(l-1).  rou is M x N x L x 2 x 2 (edge e: 1=down,2=right; layer c: 1=u,2=v).
 */
#ifndef QGMAP_ORACLE_H
#define QGMAP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int Mo, No;          /* image rows, cols                                                     */
    int M, N;            /* belief grid: Mo x No (full-res) or Mo/4 x No/4 (super-pixel)          */
    int L, K;            /* mixture components, quadrature order per axis (reference names)       */
    int super;           /* 0: gqmap_gpu_mixture.m   1: gqmap_gpuSuper_mix_entropy.m              */
    double lambdad, lambdas, epsn;
    double minu, maxu, minv, maxv;
    double sigma_min, sigma_max;   /* 0.01, 23 (full :43-44) / 25 (super :42-43)                  */
    double corr_tor;               /* 1-1e-5  (:7)                                                */
    double step0, step_tau;        /* 0.1, 8000 (full :27) / 0.001, 4000 (super :26)              */
    int    alpha_start;            /* 500  (:50)                                                  */
    double alpha_scale;            /* 1e-7 (:83)                                                  */
    int    alpha_mode;             /* 0 softmax (live code), 1 projsplx (commented :49)           */
    int    anneal_every;           /* 0 = off (full :73 is commented), 500 (super :72)            */
    double drate, T_floor;         /* T = max(T*drate, T_floor); floor 0.001 (super :72)          */
    double tor;                    /* 1e-4 stop tolerance (:25)                                   */
    double sigma_step_scale;       /* 1 (:43-44); legacy/gqmap_ctf.m:48-49 uses step*0.3          */
    int    guard_a0;               /* full-res node/edge loops skip accumulators when a==0 (:98)  */
    int    nthreads;               /* OpenMP threads (0 = default)                                */
} qo_config;

typedef struct {
    double *muu, *muv, *sigu, *sigv, *pn;   /* M*N*L each        */
    double *rou;                            /* M*N*L*2*2         */
    double *w, *alpha;                      /* L each            */
    double  T;                              /* current temperature */
} qo_state;

/* the 14 gradient arrays of one iteration, exactly as the reference materialises them (:29,:31) */
typedef struct {
    double *dan, *dmuu, *dmuv, *dsigmau, *dsigmav, *dpn, *nEnergy;        /* M*N*L      */
    double *dae, *dmu1, *dmu2, *dsigma1, *dsigma2, *drou, *eEnergy;       /* M*N*L*2*2  */
} qo_grads;

/* GaussHermite_2.m:21-32 */
int  qo_gauss_hermite(int n, double *x, double *w);
/* gqmap_gpu_mixture.m:191-208 : VV is (M+2)x(N+2) column-major */
void qo_get_vv(const double *V, int M, int N, double *VV);
/* gqmap_gpu_mixture.m:156-179 ; i,j 1-based image row/col */
double qo_node_pot(const qo_config *c, const double *I1, const double *VV, double x1, double x2, int i, int j);
/* legacy/gqmap_ctf.m:10,96: optional nearest lookup into I2_cont = interp2(I2,rfc,'cubic') as the data term (NULL switches it off);
 * interp2(V,k,'cubic') restated through getVV + the weights of node_pot (what the reference hand-copies from interp2.m) */
void qo_set_nearest_lookup(const double *I2_cont, int MM, int NN, int rfc2);
void qo_interp2_cubic_refine(const double *V, int M, int N, int k, double *out);
/* gqmap_gpu_mixture.m:180-182 */
double qo_edge_pot(const qo_config *c, double x1, double x2);

/* :29-34  raw node + edge gradient arrays (before assembly) */
void qo_gradients(const qo_config *c, const double *I1, const double *VV, const qo_state *s, qo_grads *g);
/* :36-40  assembled gradient: overwrites g->dmuu,dmuv,dsigmau,dsigmav in place, fills dalpha[L] */
void qo_assemble(const qo_config *c, qo_grads *g, double *dalpha);

/* :26-76 main loop body run for at most nsteps iterations starting at iteration counter *it
 * (1-based, as the reference).  Per executed iteration i (0-based within this call) writes
 * Energy[i], ptdmu[i], ptdsigma[i].  Monitoring (:52-68) is NOT run here; the host does it.
 * Returns number of iterations executed; *stopped = 1 when the reference would `break`
 * (:75) given `its`. */
int  qo_run(const qo_config *c, const double *I1, const double *VV, qo_state *s,
            int *it, int its, int nsteps, double *Energy, double *ptdmu, double *ptdsigma, int *stopped);

/* updateAlpha :78-86 / projsplx.m:15-32 */
void qo_update_alpha(const qo_config *c, qo_state *s, const double *dalpha, double step);
void qo_projsplx(const double *y, int m, double *x);

/* get_map_mex == legacy/findMixMax.m:1-38 + MATLAB R2018b fminbnd */
void qo_find_map(const double *alpha, const double *mu_u, const double *sig_u,
                 const double *mu_v, const double *sig_v, int M, int N, int L, double *map, int nthreads);
double qo_fminbnd_mixture(const double *a, const double *u, const double *o, int L,
                          double ax, double bx, double *fval, int *funccount);

/* profile_logP :148-154 (full) / super :152-169 ; uv is M x N x 2 */
double qo_profile_logp(const qo_config *c, const double *I1, const double *VV, const double *uv);
/* AEPE :63-64 (full, interior of Mo x No) / super :58-63 (repelem + crop 5:end-4).
 * map is M x N x 2 (belief grid); tflow Mo x No x 2; unknown Mo x No (uint8 0/1). */
double qo_aepe(const qo_config *c, const double *map, const double *tflow, const unsigned char *unknown);

/* legacy/flowToColor.m:37-87 + legacy/computeColor.m:33-115.  flow M x N x 2 in, img M x N x 3 uint8,
 * flo M x N x 2 (unknown zeroed), stats[4] = minu,maxu,minv,maxv, unknown M x N.  maxFlow<=0: auto. */
void qo_flow_to_color(const double *flow, int M, int N, double maxFlow,
                      unsigned char *img, double *flo, double *stats, unsigned char *unknown);

#ifdef __cplusplus
}
#endif
#endif
