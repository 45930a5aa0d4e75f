// get_map_mex.cpp -- Linux MEX gateway replacing the Windows-only get_map_mex.mexw64:
//   map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v)      (call sites gqmap_gpu_mixture.m:57, gqmap_gpuSuper_mix_entropy.m:56)
// 5 real double inputs (1x1xL, MxNxL x4), 1 output MxNx2 double.  Pure marshalling over qgmap_find_map (CUDA, fp64).
#include "mex_util.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    qg_nargchk(nrhs, 5, 5, nlhs, 1);
    const double *alf = qg_real_double(prhs[0], "alf");
    const double *in[4];
    size_t d[3], e[3];
    qg_dims3(prhs[1], d);
    static const char *names[4] = {"mu_u", "sig_u", "mu_v", "sig_v"};
    for (int i = 0; i < 4; ++i) {
        in[i] = qg_real_double(prhs[1 + i], names[i]);
        qg_dims3(prhs[1 + i], e);
        if (e[0] != d[0] || e[1] != d[1] || e[2] != d[2])
            mexErrMsgIdAndTxt("Coder:MATLAB:catenate_dimensionMismatch", "Dimensions of %s do not match mu_u.", names[i]);
    }
    if (mxGetNumberOfElements(prhs[0]) != d[2]) mexErrMsgIdAndTxt("qgmap:arg", "numel(alf) must equal size(mu_u,3).");
    const mwSize od[3] = {d[0], d[1], 2};
    plhs[0] = mxCreateNumericArray(3, od, mxDOUBLE_CLASS, mxREAL);
    qg_check(qgmap_find_map(alf, in[0], in[1], in[2], in[3], (int)d[0], (int)d[1], (int)d[2], mxGetPr(plhs[0]), -1), NULL);
}
