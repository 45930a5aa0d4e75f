/* peload.c -- TEST INFRASTRUCTURE: run the reference's own machine code for the MAP extraction.
 *
 * The reference ships get_map_mex only as a Windows x64 MEX binary (get_map_mex.mexw64, MATLAB Coder R2018b output of findMap.m).
 * Its per-pixel routine `findmax` (component-mean scoring + MATLAB's fminbnd, VA 0x180001c20) is plain x86-64 code whose only
 * outside calls are five scalar helpers of libmwmathutil (exp, abs, sign, max, isnan) and mxGetInf/mxGetNaN.  This loader maps
 * the PE image at its preferred base on Linux (no relocation needed), fills the import address table with stand-ins that follow
 * the Win64 calling convention, and calls findmax through an ms_abi function pointer -- so the oracle's restatement of
 * get_map_mex (qo_find_map) can be pinned against outputs of the reference ITSELF.  Nothing of the binary is copied: it is
 * read from /root/reference at run time, which is why the vectors it produces are committed as tests/golden/get_map_refbin.npz
 * (made by tests/golden/make_refbin_golden.py) for boxes where the reference is absent.
 *
 * findmax(sp, a_data, a_size, u_data, u_size, o_data, o_size, L) as recovered from the call site at VA 0x18000191d:
 *   rcx = emlrtStack* {site, tls, prev};  rdx = alpha data;  r8 = alpha size (int[3]);  r9 = means of this pixel (<= 10 doubles);
 *   [rsp+0x20] = size of the means (int[3] = 1,1,L);  [rsp+0x28] = sigmas;  [rsp+0x30] = size of the sigmas;  [rsp+0x38] = (double)L;
 *   returns the flow value in xmm0.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define MSABI __attribute__((ms_abi))
static int g_unexpected = 0;
static char g_last_unexpected[128];

static double MSABI st_exp(double x) { return exp(x); }
static double MSABI st_abs(double x) { return fabs(x); }
static double MSABI st_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : (x == 0 ? 0.0 : x)); }
static double MSABI st_max(double a, double b) { return isnan(a) ? b : (isnan(b) ? a : (a > b ? a : b)); }   /* muDoubleScalarMax */
static unsigned char MSABI st_isnan(double x) { return isnan(x) ? 1 : 0; }
static double MSABI st_inf(void) { return INFINITY; }
static double MSABI st_nan(void) { return NAN; }
/* any other import (emlrt error / bounds-check paths) must never run for valid inputs: count and remember it */
#define NGEN 100
static const char *g_gen_names[NGEN];
#define GEN(i) static uint64_t MSABI st_gen##i(void) { ++g_unexpected; snprintf(g_last_unexpected, sizeof g_last_unexpected, "%s", g_gen_names[i] ? g_gen_names[i] : "?"); return 0; }
#define G10(t) GEN(t##0) GEN(t##1) GEN(t##2) GEN(t##3) GEN(t##4) GEN(t##5) GEN(t##6) GEN(t##7) GEN(t##8) GEN(t##9)
GEN(0) GEN(1) GEN(2) GEN(3) GEN(4) GEN(5) GEN(6) GEN(7) GEN(8) GEN(9) G10(1) G10(2) G10(3) G10(4) G10(5) G10(6) G10(7) G10(8) G10(9)
#define P(i) (void *)st_gen##i
#define P10(t) P(t##0), P(t##1), P(t##2), P(t##3), P(t##4), P(t##5), P(t##6), P(t##7), P(t##8), P(t##9)
static void *g_gen[NGEN] = {P(0), P(1), P(2), P(3), P(4), P(5), P(6), P(7), P(8), P(9), P10(1), P10(2), P10(3), P10(4), P10(5), P10(6), P10(7), P10(8), P10(9)};

static uint8_t *g_img = NULL;
static uint64_t g_base = 0;
static char g_err[256];

const char *qref_error(void) { return g_err; }
int qref_unexpected_calls(char *name, int cap) { if (name) snprintf(name, cap, "%s", g_last_unexpected); return g_unexpected; }

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* map `path` (a PE32+ DLL) at its preferred image base and resolve its imports with the stand-ins above */
int qref_load(const char *path)
{
    if (g_img) return 0;
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return -1; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *file = malloc(n);
    if (fread(file, 1, n, f) != (size_t)n) { fclose(f); snprintf(g_err, sizeof g_err, "short read"); return -1; }
    fclose(f);
    const uint32_t pe = rd32(file + 0x3c);
    if (memcmp(file + pe, "PE\0\0", 4) || rd16(file + pe + 24) != 0x20b) { snprintf(g_err, sizeof g_err, "not a PE32+ image"); return -1; }
    const int nsec = rd16(file + pe + 6), optsz = rd16(file + pe + 20);
    const uint8_t *opt = file + pe + 24;
    g_base = rd64(opt + 24);
    const uint32_t imgsz = rd32(opt + 56), hdrsz = rd32(opt + 60);
    uint8_t *img = mmap((void *)g_base, imgsz, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
    if (img == MAP_FAILED || (uint64_t)img != g_base) { snprintf(g_err, sizeof g_err, "cannot map %u bytes at %#llx", imgsz, (unsigned long long)g_base); return -1; }
    memcpy(img, file, hdrsz);
    const uint8_t *sec = opt + optsz;
    for (int i = 0; i < nsec; ++i, sec += 40) {
        const uint32_t va = rd32(sec + 12), rsz = rd32(sec + 16), roff = rd32(sec + 20), vsz = rd32(sec + 8);
        memcpy(img + va, file + roff, rsz < vsz || vsz == 0 ? rsz : vsz);
    }
    /* imports */
    const uint32_t imp = rd32(opt + 112 + 8);
    int gen = 0;
    for (const uint8_t *d = img + imp; rd32(d) || rd32(d + 12) || rd32(d + 16); d += 20) {
        const uint32_t oft = rd32(d), ft = rd32(d + 16);
        for (int k = 0;; ++k) {
            const uint64_t ent = rd64(img + oft + 8 * k);
            if (!ent) break;
            const char *name = (ent >> 63) ? "(ordinal)" : (const char *)(img + (ent & 0x7fffffff) + 2);
            void *fn = NULL;
            if (!strcmp(name, "muDoubleScalarExp")) fn = (void *)st_exp;
            else if (!strcmp(name, "muDoubleScalarAbs")) fn = (void *)st_abs;
            else if (!strcmp(name, "muDoubleScalarSign")) fn = (void *)st_sign;
            else if (!strcmp(name, "muDoubleScalarMax")) fn = (void *)st_max;
            else if (!strcmp(name, "muDoubleScalarIsNaN")) fn = (void *)st_isnan;
            else if (!strcmp(name, "mxGetInf_800")) fn = (void *)st_inf;
            else if (!strcmp(name, "mxGetNaN_800")) fn = (void *)st_nan;
            else { if (gen >= NGEN) { snprintf(g_err, sizeof g_err, "too many imports"); return -1; } g_gen_names[gen] = name; fn = g_gen[gen++]; }
            memcpy(img + ft + 8 * k, &fn, 8);
        }
    }
    /* MATLAB Coder code polls `*emlrtBreakCheckR2012bFlagVar` (Ctrl-C) inside its loops; the pointer is a global that the MEX
     * entry point fills from emlrtGetBreakCheckFlagAddressR2012b().  We enter below the entry point, so point it at a zero byte. */
    static uint8_t never_break = 0;
    if (rd64(img + 0xc748) != 0) { snprintf(g_err, sizeof g_err, "unexpected image: break-check pointer slot is not empty"); return -1; }
    { void *pp = &never_break; memcpy(img + 0xc748, &pp, 8); }
    free(file);
    g_img = img;
    return 0;
}

typedef double MSABI (*findmax_fn)(const void *sp, const double *a, const int *asz, const double *u, const int *usz,
                                    const double *o, const int *osz, double L);

/* one pixel, one flow layer: the value get_map_mex writes (legacy/findMixMax.m:1-38 as compiled into the binary) */
double qref_findmax(const double *a, const double *u, const double *o, int L)
{
    static uint8_t tls[4096];
    const void *sp[3] = {NULL, tls, NULL};                  /* emlrtStack {site, tls, prev} */
    int asz[3] = {1, 1, L}, usz[3] = {1, 1, L}, osz[3] = {1, 1, L};
    double ab[16], ub[16], ob[16];                           /* the binary copies into 80-byte buffers: L <= 10 */
    memcpy(ab, a, L * sizeof(double)); memcpy(ub, u, L * sizeof(double)); memcpy(ob, o, L * sizeof(double));
    findmax_fn fn = (findmax_fn)(uintptr_t)(g_base + 0x1c20);
    return fn(sp, ab, asz, ub, usz, ob, osz, (double)L);
}

/* map = get_map_mex(alf, mu_u, sig_u, mu_v, sig_v): column-major M x N x L inputs, M x N x 2 output */
int qref_get_map(const double *alpha, const double *mu_u, const double *sig_u, const double *mu_v, const double *sig_v,
                 int M, int N, int L, double *map)
{
    if (!g_img || L < 1 || L > 10) return -1;
    const long MN = (long)M * N;
    double u[10], o[10];
    for (long p = 0; p < MN; ++p)
        for (int layer = 0; layer < 2; ++layer) {
            const double *mu = layer ? mu_v : mu_u, *sg = layer ? sig_v : sig_u;
            for (int l = 0; l < L; ++l) { u[l] = mu[p + MN * l]; o[l] = sg[p + MN * l]; }
            map[p + MN * layer] = qref_findmax(alpha, u, o, L);
        }
    return g_unexpected ? -2 : 0;
}
