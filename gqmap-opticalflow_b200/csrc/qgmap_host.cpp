// qgmap_host.cpp -- host-side pieces of libqgmap.so that need no GPU: Gauss-Hermite tables, projsplx,
// flowToColor_mex replacement, status strings.
#include "../../include/qgmap.h"
#include "qgmap_guard.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

// GaussHermite_2.m:21-32.  Golub-Welsch on the symmetric Jacobi matrix (zero diagonal, off-diagonal sqrt(i/2)),
// eigen-decomposed with cyclic Jacobi rotations (n <= 32, so O(n^3) sweeps are free); nodes ascending,
// w = sqrt(pi) * (first eigenvector component)^2.
extern "C" int qgmap_gauss_hermite(int n, double *x, double *w)
{
    if (n < 1 || n > QGMAP_KMAX || !x || !w) return QGMAP_ERR_ARG;
    std::vector<double> A((size_t)n * n, 0.0), V((size_t)n * n, 0.0);
    for (int i = 1; i < n; ++i) A[(size_t)(i - 1) * n + i] = A[(size_t)i * n + i - 1] = std::sqrt(i / 2.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[(size_t)p * n + q];
                if (std::fabs(apq) < 1e-320) continue;
                double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {                      // A <- A J
                    double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {                      // A <- J^T A
                    double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {                      // V <- V J
                    double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> ord(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
    for (int i = 0; i < n; ++i) {
        int k = ord[i];
        x[i] = A[(size_t)k * n + k];
        double v0 = V[k];                                          // first row, column k
        w[i] = std::sqrt(3.14159265358979323846) * v0 * v0;
    }
    // the rule is symmetric: enforce x(i) = -x(n-1-i) exactly and an exact 0 for odd n
    for (int i = 0; i < n / 2; ++i) {
        double xm = 0.5 * (x[n - 1 - i] - x[i]), wm = 0.5 * (w[i] + w[n - 1 - i]);
        x[i] = -xm; x[n - 1 - i] = xm; w[i] = w[n - 1 - i] = wm;
    }
    if (n & 1) x[n / 2] = 0.0;
    return QGMAP_OK;
}

// projsplx.m:15-32
extern "C" int qgmap_projsplx(const double *y, int m, double *x)
{
    if (!y || !x || m < 1) return QGMAP_ERR_ARG;
    std::vector<double> s(y, y + m);
    std::sort(s.begin(), s.end(), [](double a, double b) { return a > b; });
    bool bget = false;
    double tmpsum = 0.0, tmax = 0.0;
    for (int ii = 1; ii <= m - 1; ++ii) {
        tmpsum += s[ii - 1];
        tmax = (tmpsum - 1.0) / ii;
        if (tmax >= s[ii]) { bget = true; break; }
    }
    if (!bget) tmax = (tmpsum + s[m - 1] - 1.0) / m;
    for (int i = 0; i < m; ++i) x[i] = std::max(y[i] - tmax, 0.0);
    return QGMAP_OK;
}

// legacy/computeColor.m:67-115 (makeColorwheel): 55 hues, RY 15, YG 6, GC 4, CB 11, BM 13, MR 6.
static void colorwheel(double cw[55][3])
{
    static const int seg[6] = {15, 6, 4, 11, 13, 6};
    // per segment: channel held at 255, channel ramping, direction of the ramp
    static const int hold[6] = {0, 1, 1, 2, 2, 0}, ramp[6] = {1, 0, 2, 1, 0, 2}, down[6] = {0, 1, 0, 1, 0, 1};
    std::memset(cw, 0, sizeof(double) * 55 * 3);
    int col = 0;
    for (int s = 0; s < 6; ++s) {
        for (int i = 0; i < seg[s]; ++i) {
            double v = std::floor(255.0 * i / seg[s]);
            cw[col + i][hold[s]] = 255.0;
            cw[col + i][ramp[s]] = down[s] ? 255.0 - v : v;
        }
        col += seg[s];
    }
}

// legacy/flowToColor.m:37-87 + legacy/computeColor.m:33-65
extern "C" int qgmap_flow_to_color(const double *flow, int M, int N, double max_flow,
                                   uint8_t *img, double *flo, double *stats, uint8_t *unknown)
{
    return qg_guard([&]() -> int {
    if (!flow || M < 1 || N < 1) return QGMAP_ERR_ARG;
    const size_t MN = (size_t)M * N;
    const double UNKNOWN_FLOW_THRESH = 1e9;
    std::vector<double> ub(MN), vb(MN);
    std::vector<uint8_t> unk(MN);
    double maxu = -999, maxv = -999, minu = 999, minv = 999, maxrad = -1;
    for (size_t i = 0; i < MN; ++i) {
        double u = flow[i], v = flow[i + MN];
        bool bad = std::fabs(u) > UNKNOWN_FLOW_THRESH || std::fabs(v) > UNKNOWN_FLOW_THRESH;
        if (bad) u = v = 0.0;
        unk[i] = bad; ub[i] = u; vb[i] = v;
        maxu = std::max(maxu, u); minu = std::min(minu, u);
        maxv = std::max(maxv, v); minv = std::min(minv, v);
        maxrad = std::max(maxrad, std::sqrt(u * u + v * v));
    }
    if (max_flow > 0) maxrad = max_flow;
    if (stats) { stats[0] = minu; stats[1] = maxu; stats[2] = minv; stats[3] = maxv; }
    if (flo) for (size_t i = 0; i < MN; ++i) { flo[i] = ub[i]; flo[i + MN] = vb[i]; }
    if (unknown) std::memcpy(unknown, unk.data(), MN);
    if (!img) return QGMAP_OK;
    double cw[55][3];
    colorwheel(cw);
    const int ncols = 55;
    const double den = maxrad + DBL_EPSILON;
    for (size_t i = 0; i < MN; ++i) {
        double u = ub[i] / den, v = vb[i] / den;
        const bool isnan_ = std::isnan(u) || std::isnan(v);
        if (isnan_) u = v = 0.0;
        const double rad = std::sqrt(u * u + v * v);
        const double a = std::atan2(-v, -u) / 3.14159265358979323846;
        const double fk = (a + 1.0) / 2.0 * (ncols - 1) + 1.0;       // 1..ncols
        const int k0 = (int)std::floor(fk);
        const int k1 = (k0 + 1 == ncols + 1) ? 1 : k0 + 1;
        const double f = fk - k0;
        for (int ch = 0; ch < 3; ++ch) {
            double col = (1.0 - f) * (cw[k0 - 1][ch] / 255.0) + f * (cw[k1 - 1][ch] / 255.0);
            col = (rad <= 1.0) ? 1.0 - rad * (1.0 - col) : col * 0.75;
            double val = std::floor(255.0 * col * (isnan_ ? 0.0 : 1.0));
            val = std::min(std::max(val, 0.0), 255.0);
            img[i + MN * ch] = unk[i] ? 0 : (uint8_t)val;
        }
    }
    return QGMAP_OK;
    });
}

// ---- PNG (imwrite replacement, gqmap_gpu_mixture.m:62): 8-bit RGB, filter 0, zlib stream of stored deflate blocks ----
static uint32_t crc32_update(uint32_t c, const uint8_t *p, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t v = i;
            for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
            table[i] = v;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c;
}
static void put_be32(std::vector<uint8_t> &v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static void png_chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data)
{
    put_be32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    put_be32(out, crc32_update(0xFFFFFFFFu, td.data(), td.size()) ^ 0xFFFFFFFFu);
}
extern "C" int qgmap_write_png(const char *path, const uint8_t *rgb, int M, int N)
{
    return qg_guard([&]() -> int {
    if (!path || !*path || !rgb || M < 1 || N < 1) return QGMAP_ERR_ARG;
    const size_t MN = (size_t)M * N, row = 1 + 3 * (size_t)N;
    std::vector<uint8_t> raw(row * M);                                  // scanlines: filter byte 0 + interleaved RGB
    for (int m = 0; m < M; ++m) {
        uint8_t *r = raw.data() + row * m;
        r[0] = 0;
        for (int n = 0; n < N; ++n)
            for (int ch = 0; ch < 3; ++ch) r[1 + 3 * n + ch] = rgb[(size_t)m + (size_t)M * n + MN * ch];
    }
    std::vector<uint8_t> z;                                             // zlib: header, stored blocks (<= 65535 B), adler32
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = off; i < off + n; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        off += n;
    }
    put_be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)N); put_be32(ihdr, (uint32_t)M);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit, RGB
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    FILE *f = std::fopen(path, "wb");
    if (!f) return QGMAP_ERR_ARG;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return (std::fclose(f) == 0 && ok) ? QGMAP_OK : QGMAP_ERR_ARG;
    });
}

// ---- Middlebury .flo (readFlowFile.m:33-81, legacy/writeFlowFile.m): tag 202021.25 ("PIEH"), int32 width, height, then
// row-major interleaved float32 (u,v) ----
extern "C" int qgmap_read_flo(const char *path, int *H, int *W, double *flow)
{
    return qg_guard([&]() -> int {
    if (!path || !H || !W) return QGMAP_ERR_ARG;
    const size_t len = std::strlen(path);
    if (len < 4 || std::strcmp(path + len - 4, ".flo") != 0) return QGMAP_ERR_ARG;         // readFlowFile.m:44-52
    FILE *f = std::fopen(path, "rb");
    if (!f) return QGMAP_ERR_ARG;
    float tag = 0.f;
    int32_t wh[2] = {0, 0};
    bool ok = std::fread(&tag, 4, 1, f) == 1 && std::fread(wh, 4, 2, f) == 2 && tag == 202021.25f &&
              wh[0] >= 1 && wh[0] <= 99999 && wh[1] >= 1 && wh[1] <= 99999;                 // readFlowFile.m:60-72
    if (ok) {
        *W = wh[0]; *H = wh[1];
        if (flow) {
            const size_t n = (size_t)wh[0] * wh[1];
            // the header is untrusted: the payload it announces must really be in the file before 8n bytes are allocated for it
            const long pos = std::ftell(f);
            long end = -1;
            if (pos >= 0 && std::fseek(f, 0, SEEK_END) == 0) { end = std::ftell(f); std::fseek(f, pos, SEEK_SET); }
            if (end < 0 || (unsigned long long)(end - pos) < 8ULL * n) { std::fclose(f); return QGMAP_ERR_ARG; }
            std::vector<float> tmp(2 * n);
            ok = std::fread(tmp.data(), 4, 2 * n, f) == 2 * n;
            if (ok)
                for (int r = 0; r < wh[1]; ++r)
                    for (int c = 0; c < wh[0]; ++c) {
                        flow[(size_t)r + (size_t)wh[1] * c] = tmp[2 * ((size_t)r * wh[0] + c)];
                        flow[(size_t)r + (size_t)wh[1] * c + n] = tmp[2 * ((size_t)r * wh[0] + c) + 1];
                    }
        }
    }
    std::fclose(f);
    return ok ? QGMAP_OK : QGMAP_ERR_ARG;
    });
}
extern "C" int qgmap_write_flo(const char *path, const double *flow, int H, int W)
{
    return qg_guard([&]() -> int {
    if (!path || !flow || H < 1 || W < 1) return QGMAP_ERR_ARG;
    const size_t len = std::strlen(path);
    if (len < 4 || std::strcmp(path + len - 4, ".flo") != 0) return QGMAP_ERR_ARG;         // writeFlowFile.m:33-41
    FILE *f = std::fopen(path, "wb");
    if (!f) return QGMAP_ERR_ARG;
    const size_t n = (size_t)H * W;
    std::vector<float> tmp(2 * n);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            tmp[2 * ((size_t)r * W + c)] = (float)flow[(size_t)r + (size_t)H * c];
            tmp[2 * ((size_t)r * W + c) + 1] = (float)flow[(size_t)r + (size_t)H * c + n];
        }
    const int32_t wh[2] = {W, H};
    const bool ok = std::fwrite("PIEH", 1, 4, f) == 4 && std::fwrite(wh, 4, 2, f) == 2 && std::fwrite(tmp.data(), 4, 2 * n, f) == 2 * n;
    return (std::fclose(f) == 0 && ok) ? QGMAP_OK : QGMAP_ERR_ARG;
    });
}

extern "C" const char *qgmap_status_string(int status)
{
    switch (status) {
        case QGMAP_OK: return "ok";
        case QGMAP_ERR_ARG: return "invalid argument";
        case QGMAP_ERR_CUDA: return "CUDA error (no device, or runtime failure)";
        case QGMAP_ERR_STATE: return "invalid call order / state not set";
        case QGMAP_ERR_NOMEM: return "out of memory";
        case QGMAP_ERR_COMM: return "NCCL / band exchange error";
        default: return "unknown status";
    }
}

extern "C" int qgmap_version(void) { return QGMAP_VERSION; }
