// qgmap_advance.cuh -- the scalar part of one iteration: control-block advance (alpha update, anneal, stop test).
#pragma once
#include "qgmap_device.cuh"

// projsplx.m:15-32 on <= QG_LMAX doubles
__device__ inline void qg_projsplx(const double *y, int m, double *x) {
    double s[QG_LMAX];
    for (int i = 0; i < m; ++i) s[i] = y[i];
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j)
            if (s[j] > s[i]) { double t = s[i]; s[i] = s[j]; s[j] = t; }
    bool bget = false;
    double tmpsum = 0.0, tmax = 0.0;
    for (int ii = 1; ii <= m - 1; ++ii) {
        tmpsum += s[ii - 1];
        tmax = (tmpsum - 1.0) / (double)ii;
        if (tmax >= s[ii]) { bget = true; break; }
    }
    if (!bget) tmax = (tmpsum + s[m - 1] - 1.0) / (double)m;
    for (int i = 0; i < m; ++i) x[i] = fmax(y[i] - tmax, 0.0);
}

// Control-block advance, executed by ONE thread once the global sums of this iteration are known.
// sums[l*QG_NRED + {0:energy,1:dalpha,2:sum|G_muu|,3:sum|G_sigu|}].
__device__ inline void qg_advance(const QgIterParams &p, QgCtrl *c, const double *sums)
{
    const int L = p.L, it = c->it;
    double E = 0.0, sdm = 0.0, sds = 0.0, dalpha[QG_LMAX];
    for (int l = 0; l < L; ++l) {
        E += sums[l * QG_NRED + 0];
        dalpha[l] = sums[l * QG_NRED + 1];
        sdm += sums[l * QG_NRED + 2];
        sds += sums[l * QG_NRED + 3];
        c->dalpha[l] = dalpha[l];
    }
    const double cnt = (double)(p.M - 2) * (double)(p.N - 2) * (double)L;
    const double ptdmu = sdm / cnt, ptdsig = sds / cnt;
    p.hist_energy[it - 1] = E;                                                   // :48
    p.hist_dmu[it - 1] = ptdmu;                                                  // :69-70
    p.hist_dsig[it - 1] = ptdsig;
    const double step = p.step0 / (1.0 + (double)it / p.step_tau);               // :27
    if (it > p.alpha_start && L != 1) {                                          // :50
        if (p.alpha_mode == 1) {                                                 // :49 (commented alternative)
            double y[QG_LMAX];
            for (int l = 0; l < L; ++l) y[l] = c->alpha[l] + dalpha[l] * step * p.alpha_scale;
            qg_projsplx(y, L, c->alpha);
        } else {                                                                 // updateAlpha :78-86
            double dot = 0.0, se = 0.0;
            for (int l = 0; l < L; ++l) dot += dalpha[l] * c->alpha[l];
            for (int l = 0; l < L; ++l) {
                double dw = c->alpha[l] * (dalpha[l] - dot);
                c->w[l] = fmin(fmax(c->w[l] + dw * step * p.alpha_scale, -300.0), 300.0);
            }
            for (int l = 0; l < L; ++l) se += exp(c->w[l]);
            for (int l = 0; l < L; ++l) c->alpha[l] = exp(c->w[l]) / se;
        }
    }
    if (p.anneal_every > 0 && it % p.anneal_every == 0) c->T = fmax(c->T * p.drate, p.T_floor);   // S:72
    c->it = it + 1;                                                              // :74
    if (it + 1 > c->its || ptdmu < p.tor) c->stop = 1;                           // :75
}

