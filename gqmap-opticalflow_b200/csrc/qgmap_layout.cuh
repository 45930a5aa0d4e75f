// qgmap_layout.cuh -- layout conversion at the C-ABI boundary (MATLAB column-major fp64 <-> private row-major planes)
// and the device-side getVV.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

// ---- layout conversion kernels (MATLAB column-major fp64 at the boundary <-> private row-major planes) ------------
// src: column-major fp64 [rows x cols x planes]; dst: row-major T planes with pitch; copies rows [r0,r1) to local rows.
template <typename T, typename S = double>
__global__ void qgmap_import_kernel(const S *__restrict__ src, int rows, int cols, int planes, T *__restrict__ dst,
                                    int pitch, long long plane_stride, int r0, int r1, int g0)
{
    __shared__ S tile[32][33];
    const int pz = blockIdx.z;
    const int rb = r0 + blockIdx.y * 32, cb = blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {           // coalesced along rows (column-major source)
        int rr = rb + threadIdx.x, cc = cb + k;
        if (rr < r1 && cc < cols) tile[k][threadIdx.x] = src[rr + (long long)rows * cc + (long long)rows * cols * pz];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {           // coalesced along cols (row-major destination)
        int rr = rb + k, cc = cb + threadIdx.x;
        if (rr < r1 && cc < cols) dst[(long long)pz * plane_stride + (long long)(rr - g0) * pitch + cc] = (T)tile[threadIdx.x][k];
    }
}

template <typename T, typename D = double>
__global__ void qgmap_export_kernel(const T *__restrict__ src, int pitch, long long plane_stride, int g0, int r0, int r1,
                                    D *__restrict__ dst, int rows, int cols)
{
    __shared__ D tile[32][33];
    const int pz = blockIdx.z;
    const int rb = r0 + blockIdx.y * 32, cb = blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int rr = rb + k, cc = cb + threadIdx.x;
        if (rr < r1 && cc < cols) tile[k][threadIdx.x] = (D)src[(long long)pz * plane_stride + (long long)(rr - g0) * pitch + cc];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int rr = rb + threadIdx.x, cc = cb + k;
        if (rr < r1 && cc < cols) dst[rr + (long long)rows * cc + (long long)rows * cols * pz] = tile[threadIdx.x][k];
    }
}

// getVV (gqmap_gpu_mixture.m:191-208) on the device, fp64 row-major.  The interior is the frame itself (qgmap_import_kernel into
// VV + pitchV + 1, a coalesced transpose); this pass extrapolates the top/bottom rows of every padded column (the corner inputs are
// still zero, as in the reference: VV is zero-filled before), the next one the left/right columns of every row.
static __global__ void qgmap_vv_rows_kernel(int Mo, int No, double *__restrict__ VV, int pitchV)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;            // padded column 0..No+1
    if (c > No + 1) return;
    auto in = [&](int r) -> double { return VV[(long long)r * pitchV + c]; };             // padded row r in 1..Mo
    VV[c] = (3.0 * in(1) - 3.0 * in(2)) + in(3);                                          // :201
    VV[(long long)(Mo + 1) * pitchV + c] = (3.0 * in(Mo) - 3.0 * in(Mo - 1)) + in(Mo - 2); // :202
}
static __global__ void qgmap_vv_cols_kernel(int Mo, int No, double *__restrict__ VV, int pitchV)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;            // padded row 0..Mo+1
    if (r > Mo + 1) return;
    double *row = VV + (long long)r * pitchV;
    row[0] = (3.0 * row[1] - 3.0 * row[2]) + row[3];                                      // :205
    row[No + 1] = (3.0 * row[No] - 3.0 * row[No - 1]) + row[No - 2];                      // :206
}
template <typename T>
__global__ void qgmap_cast_kernel(const double *__restrict__ src, T *__restrict__ dst, long long n)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        dst[t] = (T)src[t];
}

// Packed second frame for the iteration kernel (QgTap8): entry (y,x) = rows y, y+1 x columns x..x+3, interleaved by row inside
// each column: out[8*(y*pitch8+x) + 2c + {0,1}] = VV(y + {0,1}, x + c)  (zero beyond the padded image), y = 0..rows_out-1.
static __global__ void qgmap_pack8_kernel(const double *__restrict__ VV, int pitchV, int rows_src, int width, float4 *__restrict__ out,
                                          int pitch8, int rows_out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= pitch8 || y >= rows_out) return;
    auto g = [&](int r, int c) -> float { return (r < rows_src && c < width) ? (float)VV[(long long)r * pitchV + c] : 0.0f; };
    float4 *o = out + 2 * ((long long)y * pitch8 + x);
    o[0] = make_float4(g(y, x), g(y + 1, x), g(y, x + 1), g(y + 1, x + 1));
    o[1] = make_float4(g(y, x + 2), g(y + 1, x + 2), g(y, x + 3), g(y + 1, x + 3));
}

// fp16 4 x 4 block layout (QgTap16h): out[(y*pitch8+x)] word c = (VV(y,x+c), VV(y+1,x+c)), word 4+c = (VV(y+2,x+c), VV(y+3,x+c)), zero beyond
// the padded image.  *inexact is raised if any value of the padded frame does not survive fp64 -> fp16 -> fp64 (then the layout is
// not used: the kernels only take it when it is bit-exact).
static __global__ void qgmap_pack16h_kernel(const double *__restrict__ VV, int pitchV, int rows_src, int width, uint4 *__restrict__ out,
                                            int pitch8, int rows_out, int *__restrict__ inexact)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= pitch8 || y >= rows_out) return;
    bool bad = false;
    auto g = [&](int r, int c) -> unsigned int {
        if (r >= rows_src || c >= width) return 0u;
        const double v = VV[(long long)r * pitchV + c];
        const __half hv = __double2half(v);
        if ((double)__half2float(hv) != v) bad = true;
        return (unsigned int)__half_as_ushort(hv);
    };
    unsigned int w[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        w[c] = g(y, x + c) | (g(y + 1, x + c) << 16);
        w[4 + c] = g(y + 2, x + c) | (g(y + 3, x + c) << 16);
    }
    uint4 *o = out + 2 * ((long long)y * pitch8 + x);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    if (bad) *inexact = 1;
}
