// gqmap_mex.cpp -- stateful MEX gateway over libqgmap.so; the loop body of the rewritten gqmap_gpu_mixture.m /
// gqmap_gpuSuper_mix_entropy.m (matlab/) is one call into here: no gpuArray, no arrayfun.
//   [mu,sigma,alpha,AEPE,Energy,logP] = gqmap_mex('solve', variant, options, I1, I2)   one-call solver (qgmap_solve)
//   h = gqmap_mex('create', variant, options, I1, I2)          variant: 0 full-res, 1 super-pixel
//   gqmap_mex('set_state', h, S)   S: struct muu,muv,sigmau,sigmav,pn,rou,w [,T,it]     | gqmap_mex('init_state', h, seed)
//   S = gqmap_mex('get_state', h)
//   [E,dmu,dsig,nit,stopped] = gqmap_mex('step', h, n, its)
//   map = gqmap_mex('map', h) ; lp = gqmap_mex('logp', h, map) ; a = gqmap_mex('aepe', h, map, tflow, unknownIdx)
//   gqmap_mex('destroy', h)
// Handles are uint64 scalars; live handles are destroyed by the mexAtExit hook.
#include "mex_util.h"
#include <math.h>
#include <stdlib.h>

#define QG_MAXH 64
static qgmap_handle *g_handles[QG_MAXH];
static int g_registered = 0;

static void at_exit(void) { for (int i = 0; i < QG_MAXH; ++i) if (g_handles[i]) { qgmap_destroy(g_handles[i]); g_handles[i] = NULL; } }

static qgmap_handle *get_handle(const mxArray *a) {
    if (!a || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("qgmap:arg", "handle must be a scalar.");
    int slot = (int)mxGetScalar(a) - 1;
    if (slot < 0 || slot >= QG_MAXH || !g_handles[slot]) mexErrMsgIdAndTxt("qgmap:state", "invalid or destroyed handle.");
    return g_handles[slot];
}

static void images(const mxArray *a, const mxArray *b, const double **I1, const double **I2, int *Mo, int *No) {
    *I1 = qg_real_double(a, "I1"); *I2 = qg_real_double(b, "I2");
    if (mxGetM(a) != mxGetM(b) || mxGetN(a) != mxGetN(b)) mexErrMsgIdAndTxt("qgmap:arg", "I1 and I2 must have equal size.");
    *Mo = (int)mxGetM(a); *No = (int)mxGetN(a);
}

static const char *kState[7] = {"muu", "muv", "sigmau", "sigmav", "pn", "rou", "w"};

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("EMLRT:runTime:WrongNumberOfInputs", "gqmap_mex(command, ...)");
    if (!g_registered) { mexAtExit(at_exit); g_registered = 1; }
    char *cmd = mxArrayToString(prhs[0]);
    char c[32];
    strncpy(c, cmd ? cmd : "", sizeof c - 1); c[sizeof c - 1] = 0;
    mxFree(cmd);

    if (!strcmp(c, "solve")) {
        qg_nargchk(nrhs, 5, 5, nlhs, 6);
        const int variant = (int)mxGetScalar(prhs[1]);
        qgmap_config cfg;
        qg_config_from_options(prhs[2], variant, &cfg);
        const double *I1, *I2; int Mo, No;
        images(prhs[3], prhs[4], &I1, &I2, &Mo, &No);
        const int its = (int)qg_field(prhs[2], "its", 1, 0);
        if (its < 1) mexErrMsgIdAndTxt("qgmap:arg", "options.its must be >= 1.");
        if (cfg.L < 1 || cfg.L > QGMAP_LMAX) mexErrMsgIdAndTxt("qgmap:arg", "options.L must be 1..%d.", QGMAP_LMAX);
        const int M = variant ? Mo / 4 : Mo, N = variant ? No / 4 : No, L = cfg.L;
        const size_t n3 = (size_t)M * N * L;
        const mxArray *tf = mxGetField(prhs[2], 0, "trueFlow"), *uk = mxGetField(prhs[2], 0, "unknownIdx");
        const mxArray *ini = mxGetField(prhs[2], 0, "init");
        const double *init[7]; int have_init = 0;
        if (ini && mxIsStruct(ini)) {
            have_init = 1;
            for (int k = 0; k < 7; ++k)         /* muu,muv,sigmau,sigmav,pn: M x N x L; rou: M x N x L x 2 x 2; w: L */
                init[k] = qg_real_double_n(mxGetField(ini, 0, kState[k]), k < 5 ? n3 : (k == 5 ? 4 * n3 : (size_t)L), kState[k]);
        }
        const mwSize d4[4] = {(mwSize)M, (mwSize)N, (mwSize)L, 2}, d3[3] = {1, 1, (mwSize)L};
        mxArray *mu = mxCreateNumericArray(4, d4, mxDOUBLE_CLASS, mxREAL), *sg = mxCreateNumericArray(4, d4, mxDOUBLE_CLASS, mxREAL);
        mxArray *al = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
        mxArray *ae = mxCreateDoubleMatrix(its, 1, mxREAL), *en = mxCreateDoubleMatrix(its, 1, mxREAL), *lp = mxCreateDoubleMatrix(its, 1, mxREAL);
        int done = 0;
        /* options.dir (gqmap_gpu_mixture.m:62): <dir>/<it>.png at every monitored iteration */
        const mxArray *dir = mxGetField(prhs[2], 0, "dir");
        char *dirs = (dir && mxIsChar(dir) && !mxIsEmpty(dir)) ? mxArrayToString(dir) : NULL;
        qgmap_solve_set_dump_dir(dirs);
        if (dirs) mxFree(dirs);
        /* options.devices (new, optional): one row band of the frame pair per listed GPU */
        const mxArray *dv = mxGetField(prhs[2], 0, "devices");
        int devs[QGMAP_P2P_RANKS_MAX], ndev = 0;
        if (dv && mxIsDouble(dv)) {
            ndev = (int)mxGetNumberOfElements(dv);
            if (ndev > QGMAP_P2P_RANKS_MAX) mexErrMsgIdAndTxt("qgmap:arg", "options.devices lists more than %d devices.", QGMAP_P2P_RANKS_MAX);
            for (int k = 0; k < ndev; ++k) devs[k] = (int)mxGetPr(dv)[k];
        }
        const double *tfp = tf && !mxIsEmpty(tf) ? qg_real_double_n(tf, (size_t)Mo * No * 2, "options.trueFlow") : NULL;
        const uint8_t *ukp = qg_logical_n(uk, (size_t)Mo * No, "options.unknownIdx");
        const uint64_t seed = (uint64_t)qg_field(prhs[2], "seed", 0, 0);
        int rc;
        if (ndev > 1)
            rc = qgmap_group_solve(&cfg, I1, I2, Mo, No, its, ndev, devs, have_init ? init : NULL, seed, tfp, ukp,
                                   mxGetPr(mu), mxGetPr(sg), mxGetPr(al), mxGetPr(ae), mxGetPr(en), mxGetPr(lp), &done);
        else
            rc = qgmap_solve(&cfg, I1, I2, Mo, No, its, have_init ? init : NULL, seed, tfp, ukp,
                             mxGetPr(mu), mxGetPr(sg), mxGetPr(al), mxGetPr(ae), mxGetPr(en), mxGetPr(lp), &done);
        qgmap_solve_set_dump_dir(NULL);
        qg_check(rc, NULL);
        mxArray *outs[6] = {mu, sg, al, ae, en, lp};
        for (int k = 0; k < 6 && (k == 0 || k < nlhs); ++k) plhs[k] = outs[k];
    } else if (!strcmp(c, "create")) {
        qg_nargchk(nrhs, 5, 5, nlhs, 1);
        const int variant = (int)mxGetScalar(prhs[1]);
        qgmap_config cfg;
        qg_config_from_options(prhs[2], variant, &cfg);
        const double *I1, *I2; int Mo, No;
        images(prhs[3], prhs[4], &I1, &I2, &Mo, &No);
        int slot = 0;
        while (slot < QG_MAXH && g_handles[slot]) ++slot;
        if (slot == QG_MAXH) mexErrMsgIdAndTxt("qgmap:state", "too many live handles (%d).", QG_MAXH);
        qg_check(qgmap_create(&cfg, I1, I2, Mo, No, &g_handles[slot]), NULL);
        plhs[0] = mxCreateDoubleScalar((double)(slot + 1));
    } else if (!strcmp(c, "destroy")) {
        qg_nargchk(nrhs, 2, 2, nlhs, 0);
        int slot = (int)mxGetScalar(prhs[1]) - 1;
        if (slot >= 0 && slot < QG_MAXH && g_handles[slot]) { qgmap_destroy(g_handles[slot]); g_handles[slot] = NULL; }
    } else if (!strcmp(c, "set_state")) {
        qg_nargchk(nrhs, 3, 3, nlhs, 0);
        qgmap_handle *h = get_handle(prhs[1]);
        if (!mxIsStruct(prhs[2])) mexErrMsgIdAndTxt("qgmap:arg", "S must be a struct.");
        const double *f[7];
        int M, N, L;
        qgmap_dims(h, &M, &N, &L);
        const size_t n3 = (size_t)M * N * L;
        for (int k = 0; k < 7; ++k)
            f[k] = qg_real_double_n(mxGetField(prhs[2], 0, kState[k]), k < 5 ? n3 : (k == 5 ? 4 * n3 : (size_t)L), kState[k]);
        const mxArray *al = mxGetField(prhs[2], 0, "alpha");
        qg_check(qgmap_set_state(h, f[0], f[1], f[2], f[3], f[4], f[5], f[6], al && !mxIsEmpty(al) ? qg_real_double_n(al, (size_t)L, "alpha") : NULL,
                                 qg_field(prhs[2], "T", 0, 0.0), (int)qg_field(prhs[2], "it", 0, 1)), h);
    } else if (!strcmp(c, "init_state")) {
        qg_nargchk(nrhs, 2, 3, nlhs, 0);
        qgmap_handle *h = get_handle(prhs[1]);
        qg_check(qgmap_init_state(h, nrhs > 2 ? (uint64_t)mxGetScalar(prhs[2]) : 0), h);
    } else if (!strcmp(c, "get_state")) {
        qg_nargchk(nrhs, 2, 2, nlhs, 1);
        qgmap_handle *h = get_handle(prhs[1]);
        int M, N, L;
        qgmap_dims(h, &M, &N, &L);
        const char *names[10] = {"muu", "muv", "sigmau", "sigmav", "pn", "rou", "w", "alpha", "T", "it"};
        mxArray *S = mxCreateStructMatrix(1, 1, 10, names);
        const mwSize d3[3] = {(mwSize)M, (mwSize)N, (mwSize)L}, d5[5] = {(mwSize)M, (mwSize)N, (mwSize)L, 2, 2}, dl[3] = {1, 1, (mwSize)L};
        mxArray *a[8];
        for (int k = 0; k < 5; ++k) a[k] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
        a[5] = mxCreateNumericArray(5, d5, mxDOUBLE_CLASS, mxREAL);
        a[6] = mxCreateNumericArray(3, dl, mxDOUBLE_CLASS, mxREAL);
        a[7] = mxCreateNumericArray(3, dl, mxDOUBLE_CLASS, mxREAL);
        double T; int it;
        qg_check(qgmap_get_state(h, mxGetPr(a[0]), mxGetPr(a[1]), mxGetPr(a[2]), mxGetPr(a[3]), mxGetPr(a[4]), mxGetPr(a[5]),
                                 mxGetPr(a[6]), mxGetPr(a[7]), &T, &it), h);
        for (int k = 0; k < 8; ++k) mxSetField(S, 0, names[k], a[k]);
        mxSetField(S, 0, "T", mxCreateDoubleScalar(T));
        mxSetField(S, 0, "it", mxCreateDoubleScalar((double)it));
        plhs[0] = S;
    } else if (!strcmp(c, "step")) {
        qg_nargchk(nrhs, 3, 4, nlhs, 5);
        qgmap_handle *h = get_handle(prhs[1]);
        const int n = (int)mxGetScalar(prhs[2]);
        const int its = nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 2147483647;
        if (n < 0 || its < 1) mexErrMsgIdAndTxt("qgmap:arg", "step: n must be >= 0 and its >= 1.");
        mxArray *E = mxCreateDoubleMatrix(n > 0 ? n : 1, 1, mxREAL), *dm = mxCreateDoubleMatrix(n > 0 ? n : 1, 1, mxREAL),
                *ds = mxCreateDoubleMatrix(n > 0 ? n : 1, 1, mxREAL);
        int done = 0, stopped = 0;
        qg_check(qgmap_step(h, n, its, mxGetPr(E), mxGetPr(dm), mxGetPr(ds), &done, &stopped), h);
        plhs[0] = E;
        if (nlhs > 1) plhs[1] = dm;
        if (nlhs > 2) plhs[2] = ds;
        if (nlhs > 3) plhs[3] = mxCreateDoubleScalar((double)done);
        if (nlhs > 4) plhs[4] = mxCreateDoubleScalar((double)stopped);
    } else if (!strcmp(c, "map")) {
        qg_nargchk(nrhs, 2, 2, nlhs, 1);
        qgmap_handle *h = get_handle(prhs[1]);
        int M, N, L;
        qgmap_dims(h, &M, &N, &L);
        const mwSize d[3] = {(mwSize)M, (mwSize)N, 2};
        plhs[0] = mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
        qg_check(qgmap_get_map(h, mxGetPr(plhs[0])), h);
    } else if (!strcmp(c, "logp")) {
        qg_nargchk(nrhs, 3, 3, nlhs, 1);
        qgmap_handle *h = get_handle(prhs[1]);
        double lp = 0;
        int M, N, L;
        qgmap_dims(h, &M, &N, &L);
        qg_check(qgmap_logp(h, qg_real_double_n(prhs[2], (size_t)M * N * 2, "map"), &lp), h);
        plhs[0] = mxCreateDoubleScalar(lp);
    } else if (!strcmp(c, "aepe")) {
        qg_nargchk(nrhs, 4, 5, nlhs, 1);
        qgmap_handle *h = get_handle(prhs[1]);
        double v = 0;
        int M, N, L, Mo, No;
        qgmap_dims(h, &M, &N, &L);
        qgmap_image_dims(h, &Mo, &No);
        qg_check(qgmap_aepe(h, qg_real_double_n(prhs[2], (size_t)M * N * 2, "map"), qg_real_double_n(prhs[3], (size_t)Mo * No * 2, "trueFlow"),
                            nrhs > 4 ? qg_logical_n(prhs[4], (size_t)Mo * No, "unknownIdx") : NULL, &v), h);
        plhs[0] = mxCreateDoubleScalar(v);
    } else {
        mexErrMsgIdAndTxt("qgmap:arg", "unknown command '%s'.", c);
    }
}
