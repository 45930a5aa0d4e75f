// qgmap_map.cu -- fp64 MAP extraction + monitoring kernels; compiled with -fmad=false (no FMA contraction) so the
// arithmetic sequence matches the fp64 reference operation by operation.
#include "qgmap_map.cuh"
#include "qgmap_internal.h"

void qgmap_launch_find_map_f32(const double *alpha, const float *mu_u, const float *sig_u, const float *mu_v,
                               const float *sig_v, long long comp_stride, int M, int N, int L, int pitch, int row_off,
                               int r0, int r1, double *map, long long total, cudaStream_t s)
{
    const int tb = 128;
    qgmap_find_map_kernel<float><<<(unsigned)((total + tb - 1) / tb), tb, 0, s>>>(alpha, mu_u, sig_u, mu_v, sig_v, comp_stride,
                                                                                   M, N, L, 1, pitch, row_off, r0, r1, map);
}
void qgmap_launch_find_map_f64(const double *alpha, const double *mu_u, const double *sig_u, const double *mu_v,
                               const double *sig_v, long long comp_stride, int M, int N, int L, double *map,
                               long long total, cudaStream_t s)
{
    const int tb = 128;
    qgmap_find_map_kernel<double><<<(unsigned)((total + tb - 1) / tb), tb, 0, s>>>(alpha, mu_u, sig_u, mu_v, sig_v, comp_stride,
                                                                                    M, N, L, 0, 0, 0, 0, M, map);
}
static QgMonParams conv(const QgMonArgs &a)
{
    QgMonParams q;
    q.I1 = a.I1; q.pitchI = a.pitchI; q.VV = a.VV; q.pitchV = a.pitchV; q.Mo = a.Mo; q.No = a.No; q.M = a.M; q.N = a.N;
    q.super = a.super; q.lambdad = a.lambdad; q.lambdas = a.lambdas; q.epsn = a.epsn; q.r0 = a.r0; q.r1 = a.r1;
    return q;
}
void qgmap_launch_logp(const QgMonArgs &q, const double *uv, double *partials, int nblk, cudaStream_t s)
{
    qgmap_logp_kernel<<<nblk, 256, 0, s>>>(conv(q), uv, partials);
}
void qgmap_launch_aepe(const QgMonArgs &q, const double *map, const double *tflow, const unsigned char *unknown,
                       double *partials, int nblk, cudaStream_t s)
{
    qgmap_aepe_kernel<<<nblk, 256, 0, s>>>(conv(q), map, tflow, unknown, partials);
}
