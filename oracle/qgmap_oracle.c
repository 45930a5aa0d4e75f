/*
 * qgmap_oracle.c -- CPU fp64 restatement of the QGMAP hot path.  TEST INFRASTRUCTURE ONLY
 * (see qgmap_oracle.h: who may load it, which functions are pinned against the reference binaries and which are not).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * Compile with -ffp-contract=off so no FMA contraction changes the fp64 arithmetic MATLAB performs.
 */
#include "qgmap_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QO_PI 3.14159265358979323846

/* ------------------------------------------------------------------------------------------------
 * GaussHermite_2.m:21-32.  Golub-Welsch: eigen-decomposition of the symmetric Jacobi matrix with zero
 * diagonal and off-diagonal sqrt(i/2); nodes = eigenvalues ascending; w = sqrt(pi) * (first component
 * of each normalised eigenvector)^2.  MATLAB's `eig` is replaced by an implicit-shift QL sweep that
 * carries only the first row of the eigenvector matrix.
 * ---------------------------------------------------------------------------------------------- */
int qo_gauss_hermite(int n, double *x, double *w)
{
    if (n < 1 || n > 256) return -1;
    double *d = (double *)calloc((size_t)n, sizeof(double));
    double *e = (double *)calloc((size_t)n + 1, sizeof(double));
    double *z = (double *)calloc((size_t)n, sizeof(double));
    for (int i = 1; i < n; ++i) e[i - 1] = sqrt((double)i / 2.0);   /* :21-23 */
    z[0] = 1.0;
    int rc = 0;
    for (int l = 0; l < n && rc == 0; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= DBL_EPSILON * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) { rc = -2; break; }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, cc = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = cc * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r; cc = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * cc * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = cc * r - b;
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + cc * f;
                    z[i] = cc * z[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    if (rc == 0) {
        /* :30-32 sort ascending, weights from first eigenvector components */
        for (int i = 0; i < n; ++i) {           /* selection sort, n is tiny */
            int k = i;
            for (int j = i + 1; j < n; ++j) if (d[j] < d[k]) k = j;
            double t = d[i]; d[i] = d[k]; d[k] = t;
            t = z[i]; z[i] = z[k]; z[k] = t;
        }
        for (int i = 0; i < n; ++i) { x[i] = d[i]; w[i] = sqrt(QO_PI) * z[i] * z[i]; }
    }
    free(d); free(e); free(z);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * getVV  gqmap_gpu_mixture.m:191-208 (== gqmap_gpuSuper_mix_entropy.m:206-223).
 * (M+2)x(N+2) copy of V with a one-pixel border extrapolated as 3a-3b+c: first the top/bottom rows of
 * every column (corner inputs are still 0 there), then the left/right columns of every row.
 * ---------------------------------------------------------------------------------------------- */
void qo_get_vv(const double *V, int M, int N, double *VV)
{
    const long M2 = M + 2, N2 = N + 2, M2N2 = M2 * N2;
    memset(VV, 0, sizeof(double) * (size_t)M2N2);
    for (long n = 0; n < N; ++n)
        for (long m = 0; m < M; ++m) VV[(m + 1) + M2 * (n + 1)] = V[m + (long)M * n];     /* :197 */
    for (long i = 1; i <= N2; ++i) {                                                        /* :198-203 */
        long ix = M2 * (i - 1) + 1, iy = ix + M2 - 1;     /* 1-based linear indices */
        VV[ix - 1] = (3.0 * VV[ix] - 3.0 * VV[ix + 1]) + VV[ix + 2];
        VV[iy - 1] = (3.0 * VV[iy - 2] - 3.0 * VV[iy - 3]) + VV[iy - 4];
    }
    for (long i = 1; i <= M2; ++i) {                                                        /* :204-207 */
        VV[i - 1] = (3.0 * VV[M2 + i - 1] - 3.0 * VV[M2 * 2 + i - 1]) + VV[M2 * 3 + i - 1];
        VV[M2N2 - M2 + i - 1] = (3.0 * VV[M2N2 - M2 * 2 + i - 1] - 3.0 * VV[M2N2 - M2 * 3 + i - 1])
                                + VV[M2N2 - M2 * 4 + i - 1];
    }
}

/* ------------------------------------------------------------------------------------------------
 * node_pot  gqmap_gpu_mixture.m:156-179 (super :171-194 uses Mo,No -- here c->Mo,c->No always are the
 * image dims).  x1 = horizontal (column) displacement, x2 = vertical (row) displacement.
 * ---------------------------------------------------------------------------------------------- */
/* legacy/gqmap_ctf.m:10,66,96 -- the coarse-to-fine solver's data term: a NEAREST lookup into the second frame upsampled 2^rfc
 * times per axis by interp2(I2,rfc,'cubic') ("I2_cont"), instead of an exact bicubic sample.  Optional (off unless set): */
static const double *g_cont = 0;
static int g_cont_MM = 0, g_cont_NN = 0, g_cont_rfc2 = 0;
void qo_set_nearest_lookup(const double *I2_cont, int MM, int NN, int rfc2)
{
    g_cont = I2_cont; g_cont_MM = MM; g_cont_NN = NN; g_cont_rfc2 = rfc2;
}
static double matlab_round(double x) { return x < 0.0 ? -floor(-x + 0.5) : floor(x + 0.5); }

/* The bicubic sample of :156-176 at (row Yq, column Xq), 1-based, already clamped to the image. */
static double bicubic_vv(const double *VV, int M, int N, double Xq, double Yq)
{
    const long M2 = M + 2;
    double ix, iy;
    if (Xq <= 1.0) ix = 1.0; else if (Xq <= (double)(N - 1)) ix = floor(Xq); else ix = (double)(N - 1);  /* :160 */
    if (Yq <= 1.0) iy = 1.0; else if (Yq <= (double)(M - 1)) iy = floor(Yq); else iy = (double)(M - 1);  /* :161 */
    double so = Xq - ix, to = Yq - iy;
    double t0 = ((2.0 - to) * to - 1.0) * to;
    double t1 = (3.0 * to - 5.0) * to * to + 2.0;
    double t2 = ((4.0 - 3.0 * to) * to + 1.0) * to;
    double t3 = (to - 1.0) * to * to;
    long iy1 = (long)iy + M2 * ((long)ix - 1);          /* 1-based linear index into VV  :165 */
    long iy2 = iy1 + M2, iy3 = iy2 + M2, iy4 = iy3 + M2;
    const double *P = VV - 1;                           /* 1-based view */
    double ss = ((2.0 - so) * so - 1.0) * so;                                                    /* :164 */
    double Vq = ((P[iy1] * ss * t0 + P[iy1 + 1] * ss * t1) + P[iy1 + 2] * ss * t2) + P[iy1 + 3] * ss * t3;  /* :169 */
    ss = (3.0 * so - 5.0) * so * so + 2.0;                                                       /* :170 */
    Vq = Vq + P[iy2] * ss * t0 + P[iy2 + 1] * ss * t1 + P[iy2 + 2] * ss * t2 + P[iy2 + 3] * ss * t3;
    ss = ((4.0 - 3.0 * so) * so + 1.0) * so;                                                     /* :172 */
    Vq = Vq + P[iy3] * ss * t0 + P[iy3 + 1] * ss * t1 + P[iy3 + 2] * ss * t2 + P[iy3 + 3] * ss * t3;
    ss = (so - 1.0) * so * so;                                                                   /* :174 */
    Vq = Vq + P[iy4] * ss * t0 + P[iy4 + 1] * ss * t1 + P[iy4 + 2] * ss * t2 + P[iy4 + 3] * ss * t3;
    return Vq / 4.0;                                                                             /* :176 */
}

/* interp2(V,k,'cubic') restated: MATLAB's classic cubic-convolution interp2 (toolbox/matlab/polyfun/interp2.m, subfunction `cubic`;
 * third party, not shipped with the reference) pads V by one quadratically extrapolated ring (3a-3b+c) and applies the Keys
 * a=-0.5 weights -- which is exactly what the reference's own getVV (:191-208) and node_pot (:160-176) hand-copy.  So the refined
 * grid is evaluated with those two restated functions; out is ((M-1)*2^k+1) x ((N-1)*2^k+1), column-major. */
void qo_interp2_cubic_refine(const double *V, int M, int N, int k, double *out)
{
    const int r = 1 << k, MM = (M - 1) * r + 1, NN = (N - 1) * r + 1;
    double *VV = (double *)malloc(sizeof(double) * (size_t)(M + 2) * (N + 2));
    qo_get_vv(V, M, N, VV);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < NN; ++b)
        for (int a = 0; a < MM; ++a)
            out[a + (size_t)MM * b] = bicubic_vv(VV, M, N, 1.0 + (double)b / r, 1.0 + (double)a / r);
    free(VV);
}

double qo_node_pot(const qo_config *c, const double *I1, const double *VV, double x1, double x2, int i, int j)
{
    const int M = c->Mo, N = c->No;
    double Vq;
    if (g_cont) {                                                                                /* legacy/gqmap_ctf.m:96 */
        long a = (long)fmin(fmax(matlab_round(((double)i + x2 - 1.0) * g_cont_rfc2 + 1.0), 1.0), (double)g_cont_MM);
        long b = (long)fmin(fmax(matlab_round(((double)j + x1 - 1.0) * g_cont_rfc2 + 1.0), 1.0), (double)g_cont_NN);
        Vq = g_cont[(a - 1) + (size_t)g_cont_MM * (b - 1)];
    } else {
        double Xq = fmin(fmax((double)j + x1, 1.0), (double)N);                 /* :157 */
        double Yq = fmin(fmax((double)i + x2, 1.0), (double)M);                 /* :158 */
        Vq = bicubic_vv(VV, M, N, Xq, Yq);
    }
    double d = I1[(i - 1) + (long)M * (j - 1)] - Vq;
    return -c->lambdad * sqrt(c->epsn + d * d);                                                  /* :178 */
}

/* edge_pot  gqmap_gpu_mixture.m:180-182 */
double qo_edge_pot(const qo_config *c, double x1, double x2)
{
    double d = x1 - x2;
    return -c->lambdas * sqrt(c->epsn + d * d);
}

/* quadrature tables  gqmap_gpu_mixture.m:8-10.  [XI,XJ]=meshgrid(X): XI(r,c)=X(c), XJ(r,c)=X(r);
 * linear index k = r + K*c (column-major), so XI varies slowest. */
typedef struct { int K2; double *XI, *XJ, *WIWJ, *XIXJ, *XI2aXJ2, *XI2mXJ2; } qo_tables;

static int tables_make(int K, qo_tables *t)
{
    double *X = (double *)malloc(sizeof(double) * (size_t)K), *W = (double *)malloc(sizeof(double) * (size_t)K);
    if (qo_gauss_hermite(K, X, W) != 0) { free(X); free(W); return -1; }
    t->K2 = K * K;
    double *buf = (double *)malloc(sizeof(double) * 6 * (size_t)t->K2);
    t->XI = buf; t->XJ = buf + t->K2; t->WIWJ = buf + 2 * t->K2; t->XIXJ = buf + 3 * t->K2;
    t->XI2aXJ2 = buf + 4 * t->K2; t->XI2mXJ2 = buf + 5 * t->K2;
    for (int cidx = 0; cidx < K; ++cidx)
        for (int r = 0; r < K; ++r) {
            int k = r + K * cidx;
            double xi = X[cidx], xj = X[r];
            t->XI[k] = xi; t->XJ[k] = xj;
            t->WIWJ[k] = W[cidx] * W[r];
            t->XIXJ[k] = xi * xj;
            t->XI2aXJ2[k] = xi * xi + xj * xj;
            t->XI2mXJ2[k] = xi * xi - xj * xj;
        }
    free(X); free(W);
    return 0;
}
static void tables_free(qo_tables *t) { free(t->XI); }

/* node_grad_spectral  gqmap_gpu_mixture.m:87-116 ; super :87-122 (16-pixel sum :94-105, no a~=0 guard) */
static void node_grad(const qo_config *c, const qo_tables *t, const double *I1, const double *VV, double T,
                      double a, double u1, double u2, double o1, double o2, double p, int m, int n, double out[7])
{
    const double sqrt2 = sqrt(2.0), const1 = 1.0 + log(2.0 * QO_PI);
    double du1 = 0, du2 = 0, do1 = 0, do2 = 0, dp = 0, Ei = 0;
    double s = (sqrt(1.0 + p) + sqrt(1.0 - p)) / 2.0;
    double tt = (sqrt(1.0 + p) - sqrt(1.0 - p)) / 2.0;
    double pr = 1.0 - p * p, sqrtpr = sqrt(pr);
    double o1pr = sqrt2 / (o1 * pr), o2pr = sqrt2 / (o2 * pr);
    int bottom = 4 * m, top = bottom - 3, right = 4 * n, left = right - 3;
    int acc = (!c->guard_a0) || (a != 0.0);
    for (int k = 0; k < t->K2; ++k) {
        double zi = s * t->XI[k] + tt * t->XJ[k], zj = tt * t->XI[k] + s * t->XJ[k];
        double x1 = sqrt2 * o1 * zi + u1, x2 = sqrt2 * o2 * zj + u2;
        double fval;
        if (c->super) {
            double sup = 0.0;
            for (int i = top; i <= bottom; ++i)
                for (int j = left; j <= right; ++j) sup = sup + qo_node_pot(c, I1, VV, x1, x2, i, j);
            fval = t->WIWJ[k] * sup;
        } else {
            fval = t->WIWJ[k] * qo_node_pot(c, I1, VV, x1, x2, m, n);
        }
        if (acc) {
            dp  = dp + fval * (p - p * t->XI2aXJ2[k] + 2.0 * t->XIXJ[k]);
            du1 = du1 + fval * (zi - p * zj);
            du2 = du2 + fval * (zj - p * zi);
            do1 = do1 + fval * (t->XI2aXJ2[k] - 1.0 + t->XI2mXJ2[k] / sqrtpr);
            do2 = do2 + fval * (t->XI2aXJ2[k] - 1.0 - t->XI2mXJ2[k] / sqrtpr);
        }
        Ei = Ei + fval;
    }
    du1 = a * du1 * o1pr / QO_PI;
    du2 = a * du2 * o2pr / QO_PI;
    double da = Ei / QO_PI - 3.0 * T * (const1 + log(sqrtpr * o1 * o2));
    do1 = a * (do1 / QO_PI - 3.0 * T) / o1;
    do2 = a * (do2 / QO_PI - 3.0 * T) / o2;
    dp = a * (dp / QO_PI + 3.0 * T * p) / pr;
    Ei = a * da;
    out[0] = da; out[1] = du1; out[2] = du2; out[3] = do1; out[4] = do2; out[5] = dp; out[6] = Ei;
}

/* edge_grad_spectral  gqmap_gpu_mixture.m:118-146 ; super :124-150 */
static void edge_grad(const qo_config *c, const qo_tables *t, double T,
                      double a, double u1, double u2, double o1, double o2, double p, double out[7])
{
    const double sqrt2 = sqrt(2.0), const1 = 1.0 + log(2.0 * QO_PI);
    double du1 = 0, du2 = 0, do1 = 0, do2 = 0, dp = 0, Ei = 0;
    double s = (sqrt(1.0 + p) + sqrt(1.0 - p)) / 2.0;
    double tt = (sqrt(1.0 + p) - sqrt(1.0 - p)) / 2.0;
    double pr = 1.0 - p * p, sqrtpr = sqrt(pr);
    double o1pr = sqrt2 / (o1 * pr), o2pr = sqrt2 / (o2 * pr);
    int acc = (!c->guard_a0) || (a != 0.0);
    for (int k = 0; k < t->K2; ++k) {
        double zi = s * t->XI[k] + tt * t->XJ[k], zj = tt * t->XI[k] + s * t->XJ[k];
        double x1 = sqrt2 * o1 * zi + u1, x2 = sqrt2 * o2 * zj + u2;
        double fval = t->WIWJ[k] * qo_edge_pot(c, x1, x2);
        if (acc) {
            dp  = dp + fval * (p - p * t->XI2aXJ2[k] + 2.0 * t->XIXJ[k]);
            du1 = du1 + fval * (zi - p * zj);
            du2 = du2 + fval * (zj - p * zi);
            do1 = do1 + fval * (t->XI2aXJ2[k] - 1.0 + t->XI2mXJ2[k] / sqrtpr);
            do2 = do2 + fval * (t->XI2aXJ2[k] - 1.0 - t->XI2mXJ2[k] / sqrtpr);
        }
        Ei = Ei + fval;
    }
    du1 = a * du1 * o1pr / QO_PI;
    du2 = a * du2 * o2pr / QO_PI;
    double da = Ei / QO_PI + T * (const1 + log(sqrtpr * o1 * o2));
    do1 = a * (do1 / QO_PI + T) / o1;
    do2 = a * (do2 / QO_PI + T) / o2;
    dp = a * (dp / QO_PI - T * p) / pr;
    Ei = a * da;
    out[0] = da; out[1] = du1; out[2] = du2; out[3] = do1; out[4] = do2; out[5] = dp; out[6] = Ei;
}

static void gradients_t(const qo_config *c, const qo_tables *t, const double *I1, const double *VV,
                        const qo_state *s, qo_grads *g)
{
    const long M = c->M, N = c->N, L = c->L, MN = M * N, MNL = MN * L;
#ifdef _OPENMP
    int nth = c->nthreads > 0 ? c->nthreads : omp_get_max_threads();
#pragma omp parallel for collapse(2) schedule(static) num_threads(nth)
#endif
    for (long l = 0; l < L; ++l)
        for (long n = 0; n < N; ++n)
            for (long m = 0; m < M; ++m) {
                long idx = m + M * n + MN * l;
                double a = s->alpha[l], out[7];
                /* :29 node kernel over M x N x L */
                node_grad(c, t, I1, VV, s->T, a, s->muu[idx], s->muv[idx], s->sigu[idx], s->sigv[idx], s->pn[idx],
                          (int)m + 1, (int)n + 1, out);
                g->dan[idx] = out[0]; g->dmuu[idx] = out[1]; g->dmuv[idx] = out[2];
                g->dsigmau[idx] = out[3]; g->dsigmav[idx] = out[4]; g->dpn[idx] = out[5]; g->nEnergy[idx] = out[6];
                /* :31-34 edge kernel over M x N x L x 2(e) x 2(c); neighbour via circshift(.,-1[,2]) (wraps) */
                for (int e = 0; e < 2; ++e) {
                    long mn = (e == 0) ? ((m + 1) % M) : m;
                    long nn = (e == 0) ? n : ((n + 1) % N);
                    long nidx = mn + M * nn + MN * l;
                    for (int cc = 0; cc < 2; ++cc) {
                        const double *mu = cc == 0 ? s->muu : s->muv;
                        const double *sg = cc == 0 ? s->sigu : s->sigv;
                        long eidx = idx + MNL * e + 2 * MNL * cc;
                        edge_grad(c, t, s->T, a, mu[idx], mu[nidx], sg[idx], sg[nidx], s->rou[eidx], out);
                        g->dae[eidx] = out[0]; g->dmu1[eidx] = out[1]; g->dmu2[eidx] = out[2];
                        g->dsigma1[eidx] = out[3]; g->dsigma2[eidx] = out[4]; g->drou[eidx] = out[5];
                        g->eEnergy[eidx] = out[6];
                    }
                }
            }
}

void qo_gradients(const qo_config *c, const double *I1, const double *VV, const qo_state *s, qo_grads *g)
{
    qo_tables t;
    if (tables_make(c->K, &t) != 0) return;
    gradients_t(c, &t, I1, VV, s, g);
    tables_free(&t);
}

/* gqmap_gpu_mixture.m:36-40 */
void qo_assemble(const qo_config *c, qo_grads *g, double *dalpha)
{
    const long M = c->M, N = c->N, L = c->L, MN = M * N, MNL = MN * L;
    for (long l = 0; l < L; ++l) {                                           /* :36 */
        double sn = 0.0, se = 0.0;
        for (long n = 1; n < N - 1; ++n)
            for (long m = 1; m < M - 1; ++m) {
                long idx = m + M * n + MN * l;
                sn += g->dan[idx];
                for (int e = 0; e < 2; ++e)
                    for (int cc = 0; cc < 2; ++cc) se += g->dae[idx + MNL * e + 2 * MNL * cc];
            }
        dalpha[l] = sn + se;
    }
    double *outs[4] = { g->dmuu, g->dmuv, g->dsigmau, g->dsigmav };
    for (int q = 0; q < 4; ++q) {                                            /* :37-40 */
        int cc = q & 1;                               /* dmuu,dsigmau -> layer u ; dmuv,dsigmav -> layer v */
        const double *d1 = (q < 2) ? g->dmu1 : g->dsigma1;
        const double *d2 = (q < 2) ? g->dmu2 : g->dsigma2;
        double *o = outs[q];
        for (long l = 0; l < L; ++l)
            for (long n = 0; n < N; ++n)
                for (long m = 0; m < M; ++m) {
                    long idx = m + M * n + MN * l;
                    long up = ((m + M - 1) % M) + M * n + MN * l;            /* circshift(.,1)   */
                    long lf = m + M * ((n + N - 1) % N) + MN * l;            /* circshift(.,1,2) */
                    double v = o[idx];
                    v = v + (d1[idx + 2 * MNL * cc] + d1[idx + MNL + 2 * MNL * cc]);   /* sum(dmu1(:,:,:,:,c),4) */
                    v = v + d2[up + 2 * MNL * cc];                                     /* down edge of (m-1,n)   */
                    v = v + d2[lf + MNL + 2 * MNL * cc];                               /* right edge of (m,n-1)  */
                    o[idx] = v;
                }
    }
}

/* projsplx.m:15-32 */
void qo_projsplx(const double *y, int m, double *x)
{
    double *s = (double *)malloc(sizeof(double) * (size_t)m);
    memcpy(s, y, sizeof(double) * (size_t)m);
    for (int i = 0; i < m; ++i)                 /* sort descending */
        for (int j = i + 1; j < m; ++j) if (s[j] > s[i]) { double t = s[i]; s[i] = s[j]; s[j] = t; }
    int bget = 0; double tmpsum = 0.0, tmax = 0.0;
    for (int ii = 1; ii <= m - 1; ++ii) {
        tmpsum = tmpsum + s[ii - 1];
        tmax = (tmpsum - 1.0) / (double)ii;
        if (tmax >= s[ii]) { bget = 1; break; }
    }
    if (!bget) tmax = (tmpsum + s[m - 1] - 1.0) / (double)m;
    for (int i = 0; i < m; ++i) x[i] = fmax(y[i] - tmax, 0.0);
    free(s);
}

/* updateAlpha  gqmap_gpu_mixture.m:78-86 ; projsplx alternative :49 */
void qo_update_alpha(const qo_config *c, qo_state *s, const double *dalpha, double step)
{
    const int L = c->L;
    if (c->alpha_mode == 1) {
        double *y = (double *)malloc(sizeof(double) * (size_t)L);
        for (int l = 0; l < L; ++l) y[l] = s->alpha[l] + dalpha[l] * step * c->alpha_scale;
        qo_projsplx(y, L, s->alpha);
        free(y);
        return;
    }
    double dot = 0.0;
    for (int l = 0; l < L; ++l) dot += dalpha[l] * s->alpha[l];
    double se = 0.0;
    for (int l = 0; l < L; ++l) {
        double dw = s->alpha[l] * (dalpha[l] - dot);
        s->w[l] = fmin(fmax(s->w[l] + dw * step * c->alpha_scale, -300.0), 300.0);
    }
    for (int l = 0; l < L; ++l) se += exp(s->w[l]);
    for (int l = 0; l < L; ++l) s->alpha[l] = exp(s->w[l]) / se;
}

static inline double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

/* main loop  gqmap_gpu_mixture.m:26-76 (super :25-75) without the monitoring block :52-68 */
int qo_run(const qo_config *c, const double *I1, const double *VV, qo_state *s,
           int *it_io, int its, int nsteps, double *Energy, double *ptdmu_o, double *ptdsigma_o, int *stopped)
{
    const long M = c->M, N = c->N, L = c->L, MN = M * N, MNL = MN * L;
    qo_tables t;
    if (tables_make(c->K, &t) != 0) return -1;
    qo_grads g;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(7 * MNL + 7 * 4 * MNL));
    g.dan = buf; g.dmuu = buf + MNL; g.dmuv = buf + 2 * MNL; g.dsigmau = buf + 3 * MNL; g.dsigmav = buf + 4 * MNL;
    g.dpn = buf + 5 * MNL; g.nEnergy = buf + 6 * MNL;
    double *eb = buf + 7 * MNL;
    g.dae = eb; g.dmu1 = eb + 4 * MNL; g.dmu2 = eb + 8 * MNL; g.dsigma1 = eb + 12 * MNL; g.dsigma2 = eb + 16 * MNL;
    g.drou = eb + 20 * MNL; g.eEnergy = eb + 24 * MNL;
    double *dalpha = (double *)malloc(sizeof(double) * (size_t)L);
    int it = *it_io, done = 0;
    *stopped = 0;
    while (done < nsteps) {
        double step = c->step0 / (1.0 + (double)it / c->step_tau);                         /* :27 */
        gradients_t(c, &t, I1, VV, s, &g);                                                 /* :29-34 */
        qo_assemble(c, &g, dalpha);                                                        /* :36-40 */
        double E = 0.0, sdm = 0.0, sds = 0.0;
        for (long l = 0; l < L; ++l)
            for (long n = 1; n < N - 1; ++n)
                for (long m = 1; m < M - 1; ++m) {
                    long idx = m + M * n + MN * l;
                    s->muu[idx]  = clampd(s->muu[idx]  + g.dmuu[idx] * step, c->minu, c->maxu);          /* :41 */
                    s->muv[idx]  = clampd(s->muv[idx]  + g.dmuv[idx] * step, c->minv, c->maxv);          /* :42 */
                    s->sigu[idx] = clampd(s->sigu[idx] + g.dsigmau[idx] * step * c->sigma_step_scale, c->sigma_min, c->sigma_max);   /* :43 */
                    s->sigv[idx] = clampd(s->sigv[idx] + g.dsigmav[idx] * step * c->sigma_step_scale, c->sigma_min, c->sigma_max);   /* :44 */
                    for (int q = 0; q < 4; ++q) {                                                         /* :45 */
                        long eidx = idx + MNL * q;
                        s->rou[eidx] = clampd(s->rou[eidx] + g.drou[eidx] * step, -c->corr_tor, c->corr_tor);
                        E += g.eEnergy[eidx];
                    }
                    s->pn[idx] = clampd(s->pn[idx] + g.dpn[idx] * step, -c->corr_tor, c->corr_tor);      /* :46 */
                    E += g.nEnergy[idx];                                                                  /* :48 */
                    sdm += fabs(g.dmuu[idx]); sds += fabs(g.dsigmau[idx]);                                /* :69 */
                }
        Energy[done] = E;
        if (it > c->alpha_start && L != 1) qo_update_alpha(c, s, dalpha, step);            /* :50 */
        double cnt = (double)((M - 2) * (N - 2) * L);
        double ptdmu = sdm / cnt, ptdsigma = sds / cnt;                                    /* :70 */
        ptdmu_o[done] = ptdmu; ptdsigma_o[done] = ptdsigma;
        if (c->anneal_every > 0 && it % c->anneal_every == 0)                              /* super :72 */
            s->T = fmax(s->T * c->drate, c->T_floor);
        it = it + 1; ++done;                                                               /* :74 */
        if (it > its || ptdmu < c->tor) { *stopped = 1; break; }                           /* :75 */
    }
    *it_io = it;
    free(dalpha); free(buf); tables_free(&t);
    return done;
}

/* ------------------------------------------------------------------------------------------------
 * get_map_mex: legacy/findMixMax.m:1-38 (findmin / neg_mixture) + MATLAB R2018b fminbnd with default
 * options (TolX 1e-4, MaxFunEvals = MaxIter = 500), control flow as in SURVEY.md Appendix A.
 * fminbnd is third-party (MathWorks toolbox/matlab/optimfun/fminbnd.m, R2018b 9.5.0.944444): its
 * published Brent golden-section/parabolic algorithm is restated here.
 * ---------------------------------------------------------------------------------------------- */
static double neg_mixture(double x, const double *a, const double *u, const double *o, int L)
{
    const double sqrt2pi = 2.5066282746310002;
    double v = 0.0;
    for (int l = 0; l < L; ++l) {                              /* findMixMax.m:33-37 */
        double d = x - u[l];
        v = v + a[l] * exp(-(d * d) / (2.0 * (o[l] * o[l]))) / (sqrt2pi * o[l]);
    }
    return -v;
}

static inline double sgn(double v) { return (v > 0.0) - (v < 0.0); }

double qo_fminbnd_mixture(const double *al, const double *u, const double *o, int L,
                          double ax, double bx, double *fval, int *funccount_o)
{
    const double tol = 1e-4;
    double a = ax, b = bx;
    const double seps = sqrt(DBL_EPSILON), c = 0.5 * (3.0 - sqrt(5.0));
    double v = a + c * (b - a), w = v, xf = v, d = 0.0, e = 0.0, x = xf;
    double fx = neg_mixture(x, al, u, o, L);
    int funccount = 1, iter = 0;
    double fv = fx, fw = fx, xm = 0.5 * (a + b), tol1 = seps * fabs(xf) + tol / 3.0, tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        int gs = 1;
        if (fabs(e) > tol1) {
            gs = 0;
            double r = (xf - w) * (fx - fv);
            double q = (xf - v) * (fx - fw);
            double p = (xf - v) * q - (xf - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e; e = d;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                d = p / q; x = xf + d;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    double si = sgn(xm - xf) + ((xm - xf) == 0.0);
                    d = tol1 * si;
                }
            } else gs = 1;
        }
        if (gs) {
            if (xf >= xm) e = a - xf; else e = b - xf;
            d = c * e;
        }
        double si = sgn(d) + (d == 0.0);
        x = xf + si * fmax(fabs(d), tol1);
        double fu = neg_mixture(x, al, u, o, L);
        funccount++; iter++;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            v = w; fv = fw; w = xf; fw = fx; xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fw) || (w == xf)) { v = w; fv = fw; w = x; fw = fu; }
            else if ((fu <= fv) || (v == xf) || (v == w)) { v = x; fv = fu; }
        }
        xm = 0.5 * (a + b); tol1 = seps * fabs(xf) + tol / 3.0; tol2 = 2.0 * tol1;
        if (funccount >= 500 || iter >= 500) break;
    }
    *fval = fx;
    if (funccount_o) *funccount_o = funccount;
    return xf;
}

static double findmin(const double *a, const double *u, const double *o, int L)   /* findMixMax.m:15-30 */
{
    const double sqrt2pi = 2.5066282746310002;
    double spk = INFINITY; int sid = 0;
    double umin = u[0], umax = u[0];
    for (int l1 = 0; l1 < L; ++l1) {
        double vl = 0.0;
        for (int l2 = 0; l2 < L; ++l2) {
            double d = u[l1] - u[l2];
            vl = vl - a[l2] * exp(-(d * d) / (2.0 * (o[l2] * o[l2]))) / (sqrt2pi * o[l2]);
        }
        if (vl < spk) { spk = vl; sid = l1; }
        umin = fmin(umin, u[l1]); umax = fmax(umax, u[l1]);
    }
    double fval;
    double x = qo_fminbnd_mixture(a, u, o, L, umin, umax, &fval, NULL);
    return (fval < spk) ? x : u[sid];
}

void qo_find_map(const double *alpha, const double *mu_u, const double *sig_u,
                 const double *mu_v, const double *sig_v, int M, int N, int L, double *map, int nthreads)
{
    const long MN = (long)M * N;
#ifdef _OPENMP
    int nth = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nth)
#endif
    for (long i = 0; i < MN; ++i) {
        double u[64], o[64];
        for (int l = 0; l < L; ++l) { u[l] = mu_u[i + MN * l]; o[l] = sig_u[i + MN * l]; }
        map[i] = findmin(alpha, u, o, L);
        for (int l = 0; l < L; ++l) { u[l] = mu_v[i + MN * l]; o[l] = sig_v[i + MN * l]; }
        map[i + MN] = findmin(alpha, u, o, L);
    }
    (void)nthreads;
}

/* profile_logP  gqmap_gpu_mixture.m:148-154 ; super :152-169 (node_lp = 16-pixel sum) */
double qo_profile_logp(const qo_config *c, const double *I1, const double *VV, const double *uv)
{
    const long M = c->M, N = c->N, MN = M * N;
    double lp = 0.0;
    for (long n = 1; n < N - 1; ++n)
        for (long m = 1; m < M - 1; ++m) {
            double us = uv[m + M * n], vs = uv[m + M * n + MN];
            if (c->super) {
                for (int i = 4 * (int)(m + 1) - 3; i <= 4 * (int)(m + 1); ++i)
                    for (int j = 4 * (int)(n + 1) - 3; j <= 4 * (int)(n + 1); ++j) lp += qo_node_pot(c, I1, VV, us, vs, i, j);
            } else lp += qo_node_pot(c, I1, VV, us, vs, (int)m + 1, (int)n + 1);
            for (int cc = 0; cc < 2; ++cc) {
                double x = uv[m + M * n + MN * cc];
                lp += qo_edge_pot(c, x, uv[((m + 1) % M) + M * n + MN * cc]);          /* circshift(uv,-1)   */
                lp += qo_edge_pot(c, x, uv[m + M * ((n + 1) % N) + MN * cc]);          /* circshift(uv,-1,2) */
            }
        }
    return lp;
}

/* AEPE  gqmap_gpu_mixture.m:59-64 ; super gqmap_gpuSuper_mix_entropy.m:58-63 */
double qo_aepe(const qo_config *c, const double *map, const double *tflow, const unsigned char *unknown)
{
    const long Mo = c->Mo, No = c->No, M = c->M, N = c->N;
    const long r0 = c->super ? 4 : 1, r1 = c->super ? Mo - 4 : Mo - 1;       /* 0-based half-open row range */
    const long c0 = c->super ? 4 : 1, c1 = c->super ? No - 4 : No - 1;
    const int sc = c->super ? 4 : 1;
    double sum = 0.0;
    for (long j = c0; j < c1; ++j)
        for (long i = r0; i < r1; ++i) {
            long mi = i / sc, mj = j / sc;                 /* repelem(map,4,4) */
            double fu = map[mi + M * mj], fv = map[mi + M * mj + M * N];
            if (unknown && unknown[i + Mo * j]) { fu = 0.0; fv = 0.0; }       /* flow(unidx)=0 */
            double du = tflow[i + Mo * j] - fu, dv = tflow[i + Mo * j + Mo * No] - fv;
            sum += sqrt(du * du + dv * dv);
        }
    return sum / (double)((r1 - r0) * (c1 - c0));
}

/* legacy/computeColor.m:67-115 makeColorwheel */
static void make_colorwheel(double cw[55][3])
{
    const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
    memset(cw, 0, sizeof(double) * 55 * 3);
    int col = 0;
    for (int i = 0; i < RY; ++i) { cw[i][0] = 255; cw[i][1] = floor(255.0 * i / RY); }
    col += RY;
    for (int i = 0; i < YG; ++i) { cw[col + i][0] = 255 - floor(255.0 * i / YG); cw[col + i][1] = 255; }
    col += YG;
    for (int i = 0; i < GC; ++i) { cw[col + i][1] = 255; cw[col + i][2] = floor(255.0 * i / GC); }
    col += GC;
    for (int i = 0; i < CB; ++i) { cw[col + i][1] = 255 - floor(255.0 * i / CB); cw[col + i][2] = 255; }
    col += CB;
    for (int i = 0; i < BM; ++i) { cw[col + i][2] = 255; cw[col + i][0] = floor(255.0 * i / BM); }
    col += BM;
    for (int i = 0; i < MR; ++i) { cw[col + i][2] = 255 - floor(255.0 * i / MR); cw[col + i][0] = 255; }
}

/* legacy/flowToColor.m:37-87 + legacy/computeColor.m:33-65 */
void qo_flow_to_color(const double *flow, int M, int N, double maxFlow,
                      unsigned char *img, double *flo, double *stats, unsigned char *unknown)
{
    const long MN = (long)M * N;
    const double TH = 1e9;
    double maxu = -999, maxv = -999, minu = 999, minv = 999, maxrad = -1;
    for (long i = 0; i < MN; ++i) {
        double u = flow[i], v = flow[i + MN];
        unsigned char unk = (fabs(u) > TH) || (fabs(v) > TH);
        if (unk) { u = 0; v = 0; }
        unknown[i] = unk; flo[i] = u; flo[i + MN] = v;
        maxu = fmax(maxu, u); minu = fmin(minu, u); maxv = fmax(maxv, v); minv = fmin(minv, v);
        maxrad = fmax(maxrad, sqrt(u * u + v * v));
    }
    if (maxFlow > 0) maxrad = maxFlow;
    stats[0] = minu; stats[1] = maxu; stats[2] = minv; stats[3] = maxv;
    double cw[55][3];
    make_colorwheel(cw);
    const int ncols = 55;
    for (long i = 0; i < MN; ++i) {
        double u = flo[i] / (maxrad + DBL_EPSILON), v = flo[i + MN] / (maxrad + DBL_EPSILON);
        int nan = isnan(u) || isnan(v);
        if (nan) { u = 0; v = 0; }
        double rad = sqrt(u * u + v * v);
        double a = atan2(-v, -u) / QO_PI;
        double fk = (a + 1.0) / 2.0 * (ncols - 1) + 1.0;
        int k0 = (int)floor(fk), k1 = k0 + 1;
        if (k1 == ncols + 1) k1 = 1;
        double f = fk - k0;
        for (int ch = 0; ch < 3; ++ch) {
            double col0 = cw[k0 - 1][ch] / 255.0, col1 = cw[k1 - 1][ch] / 255.0;
            double col = (1.0 - f) * col0 + f * col1;
            if (rad <= 1.0) col = 1.0 - rad * (1.0 - col); else col = col * 0.75;
            double val = floor(255.0 * col * (1.0 - nan));
            if (val < 0) val = 0;
            if (val > 255) val = 255;
            img[i + MN * ch] = unknown[i] ? 0 : (unsigned char)val;
        }
    }
}
