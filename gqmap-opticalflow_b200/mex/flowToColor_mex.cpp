// flowToColor_mex.cpp -- Linux MEX gateway replacing flowToColor_mex.mexw64:
//   [img,flo,minu,maxu,minv,maxv,idxUnknown] = flowToColor_mex(flow[,maxFlow])   (optical_flow.m:12-13, gqmap_gpu_mixture.m:60)
// flow MxNx2 double -> img MxNx3 uint8, flo MxNx2 double, 4 scalars, idxUnknown MxN logical (legacy/flowToColor.m:1,37-87).
#include "mex_util.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    qg_nargchk(nrhs, 1, 2, nlhs, 7);
    const double *flow = qg_real_double(prhs[0], "flow");
    size_t d[3];
    qg_dims3(prhs[0], d);
    if (d[2] != 2) mexErrMsgIdAndTxt("flowToColor:bands", "flowToColor: image must have two bands");   // legacy/flowToColor.m:41-43
    const double maxFlow = nrhs > 1 ? mxGetScalar(prhs[1]) : -1.0;
    const mwSize d3[3] = {d[0], d[1], 3}, d2[3] = {d[0], d[1], 2}, d1[2] = {d[0], d[1]};
    mxArray *img = mxCreateNumericArray(3, d3, mxUINT8_CLASS, mxREAL);
    mxArray *flo = mxCreateNumericArray(3, d2, mxDOUBLE_CLASS, mxREAL);
    mxArray *unk = mxCreateLogicalArray(2, d1);
    double stats[4];
    qg_check(qgmap_flow_to_color(flow, (int)d[0], (int)d[1], maxFlow, (uint8_t *)mxGetData(img), mxGetPr(flo), stats,
                                 (uint8_t *)mxGetLogicals(unk)), NULL);
    plhs[0] = img;
    if (nlhs > 1) plhs[1] = flo;
    for (int k = 0; k < 4; ++k) if (nlhs > 2 + k) plhs[2 + k] = mxCreateDoubleScalar(stats[k]);
    if (nlhs > 6) plhs[6] = unk;
}
