// mex_util.h -- argument marshalling shared by the MEX gateways (thin layer over include/qgmap.h).
// Error behaviour mirrors the reference binaries: wrong argument counts raise the EMLRT ids the MATLAB Coder
// gateways used; library failures are reported as qgmap:<kind> after all resources are released.  No C++ exception
// crosses mexFunction.
#pragma once
#include "mex.h"
#include "../../include/qgmap.h"
#include <string.h>

static inline void qg_nargchk(int nrhs, int lo, int hi, int nlhs, int maxlhs) {
    if (nrhs < lo || nrhs > hi) mexErrMsgIdAndTxt("EMLRT:runTime:WrongNumberOfInputs", "%d inputs required, %d given.", lo, nrhs);
    if (nlhs > maxlhs) mexErrMsgIdAndTxt("EMLRT:runTime:TooManyOutputArguments", "Too many output arguments.");
}
static inline const double *qg_real_double(const mxArray *a, const char *what) {
    if (!a || !mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("qgmap:arg", "%s must be a real double array.", what);
    return mxGetPr(a);
}
// a real double array with exactly `n` elements (the C ABI takes plain pointers: sizes must be checked here)
static inline const double *qg_real_double_n(const mxArray *a, size_t n, const char *what) {
    const double *p = qg_real_double(a, what);
    if (mxGetNumberOfElements(a) != n)
        mexErrMsgIdAndTxt("qgmap:arg", "%s has %lu elements, %lu expected.", what, (unsigned long)mxGetNumberOfElements(a), (unsigned long)n);
    return p;
}
// options.unknownIdx / the mask of 'aepe': a LOGICAL array with `n` elements, or empty / absent (NULL)
static inline const uint8_t *qg_logical_n(const mxArray *a, size_t n, const char *what) {
    if (!a || mxIsEmpty(a)) return NULL;
    if (!mxIsLogical(a)) mexErrMsgIdAndTxt("qgmap:arg", "%s must be a logical array.", what);
    if (mxGetNumberOfElements(a) != n)
        mexErrMsgIdAndTxt("qgmap:arg", "%s has %lu elements, %lu expected.", what, (unsigned long)mxGetNumberOfElements(a), (unsigned long)n);
    return (const uint8_t *)mxGetLogicals(a);
}
static inline void qg_dims3(const mxArray *a, size_t d[3]) {
    mwSize nd = mxGetNumberOfDimensions(a);
    const mwSize *dd = mxGetDimensions(a);
    d[0] = dd[0]; d[1] = nd > 1 ? dd[1] : 1; d[2] = nd > 2 ? dd[2] : 1;
}
static inline void qg_check(int rc, qgmap_handle *h) {
    if (rc != QGMAP_OK) {
        const char *id = rc == QGMAP_ERR_CUDA ? "qgmap:cuda" : rc == QGMAP_ERR_ARG ? "qgmap:arg" : rc == QGMAP_ERR_STATE ? "qgmap:state"
                       : rc == QGMAP_ERR_COMM ? "qgmap:comm" : "qgmap:error";
        mexErrMsgIdAndTxt(id, "%s: %s", qgmap_status_string(rc), qgmap_last_error(h));
    }
}
static inline double qg_field(const mxArray *s, const char *name, int required, double dflt) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || mxIsEmpty(f)) {
        if (required) mexErrMsgIdAndTxt("qgmap:arg", "options.%s is required (gqmap_gpu_mixture.m:3-6).", name);
        return dflt;
    }
    return mxGetScalar(f);
}
// options struct (gqmap_gpu_mixture.m:3-6) -> qgmap_config
static inline void qg_config_from_options(const mxArray *opt, int variant, qgmap_config *cfg) {
    if (!mxIsStruct(opt)) mexErrMsgIdAndTxt("qgmap:arg", "options must be a struct.");
    qgmap_config_defaults(cfg, variant);
    cfg->K = (int)qg_field(opt, "K", 1, 0);
    cfg->L = (int)qg_field(opt, "L", 1, 0);
    cfg->temperature = qg_field(opt, "temperature", 1, 0);
    cfg->drate = qg_field(opt, "drate", 1, 0);
    cfg->epsn = qg_field(opt, "epsn", 1, 0);
    cfg->lambdad = qg_field(opt, "lambdad", 1, 0);
    cfg->lambdas = qg_field(opt, "lambdas", 1, 0);
    cfg->minu = qg_field(opt, "minu", 1, 0); cfg->maxu = qg_field(opt, "maxu", 1, 0);
    cfg->minv = qg_field(opt, "minv", 1, 0); cfg->maxv = qg_field(opt, "maxv", 1, 0);
    cfg->device = (int)qg_field(opt, "device", 0, -1);
    cfg->log_every = (int)qg_field(opt, "log_every", 0, 300);
    /* the constants the reference hard-codes, overridable through optional fields of the same names as qgmap_config */
    cfg->sigma_min = qg_field(opt, "sigma_min", 0, cfg->sigma_min); cfg->sigma_max = qg_field(opt, "sigma_max", 0, cfg->sigma_max);
    cfg->corr_tor = qg_field(opt, "corr_tor", 0, cfg->corr_tor);
    cfg->step0 = qg_field(opt, "step0", 0, cfg->step0); cfg->step_tau = qg_field(opt, "step_tau", 0, cfg->step_tau);
    cfg->alpha_scale = qg_field(opt, "alpha_scale", 0, cfg->alpha_scale); cfg->T_floor = qg_field(opt, "T_floor", 0, cfg->T_floor);
    cfg->tor = qg_field(opt, "tor", 0, cfg->tor); cfg->sigma_step_scale = qg_field(opt, "sigma_step_scale", 0, cfg->sigma_step_scale);
    cfg->alpha_start = (int)qg_field(opt, "alpha_start", 0, cfg->alpha_start);
    cfg->anneal_every = (int)qg_field(opt, "anneal_every", 0, cfg->anneal_every);
    cfg->strip_rows = (int)qg_field(opt, "strip_rows", 0, cfg->strip_rows);
    const mxArray *am = mxGetField(opt, 0, "alpha_mode");
    if (am && mxIsChar(am)) {
        char *sname = mxArrayToString(am);
        cfg->alpha_mode = (sname && strcmp(sname, "projsplx") == 0) ? QGMAP_ALPHA_PROJSPLX : QGMAP_ALPHA_SOFTMAX;
        mxFree(sname);
    }
}
