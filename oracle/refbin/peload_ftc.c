/* peload_ftc.c -- TEST INFRASTRUCTURE: run the reference's own machine code for flowToColor_mex.
 *
 * The reference ships flowToColor_mex only as a Windows x64 MEX binary (flowToColor_mex.mexw64, MATLAB Coder R2018b output of
 * legacy/flowToColor.m + legacy/computeColor.m; it is on every driver's path, optical_flow.m:12-13, gqmap_gpu_mixture.m:60).
 * The binary exports its entry-point function
 *     void flowToColor(const emlrtStack *sp, const emxArray_real_T *flow, emxArray_uint8_T *img, emxArray_real_T *flo,
 *                      real_T *minu, real_T *maxu, real_T *minv, real_T *maxv, emxArray_boolean_T *idxUnknown)
 * (export table, RVA 0x10f0; argument order recovered from the call site in flowToColor_api at VA 0x1800069a2-0x1800069df), whose
 * outside calls on the success path are scalar helpers of libmwmathutil, memcpy/memset and the emlrt heap / size-check services.
 * This loader maps the PE image (anywhere: base relocations are applied, so it can live beside get_map_mex.mexw64, which
 * peload.c maps at the same preferred base), fills the import table with stand-ins that follow the Win64 calling convention and
 * calls flowToColor through an ms_abi pointer -- so the oracle's restatement (qo_flow_to_color) and the product's host C++
 * (qgmap_flow_to_color) can be pinned against outputs of the reference ITSELF.  Nothing of the binary is copied: it is read from
 * /root/reference at run time; the vectors it produces are committed as tests/golden/flow_to_color_refbin.npz.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define MSABI __attribute__((ms_abi))

/* MATLAB Coder's dynamic array (R2018b layout, confirmed by the accesses at [reg+0x8] / +0x10 / +0x18 in the binary) */
typedef struct {
    void *data;
    int32_t *size;
    int32_t allocatedSize;
    int32_t numDimensions;
    uint8_t canFreeData;
} emx_t;

static int g_unexpected = 0;
static char g_last_unexpected[128];
static void unexpected(const char *what) { ++g_unexpected; snprintf(g_last_unexpected, sizeof g_last_unexpected, "%s", what); }

/* ---- libmwmathutil scalars ---- */
static double MSABI st_sqrt(double x) { return sqrt(x); }
static double MSABI st_floor(double x) { return floor(x); }
static double MSABI st_atan2(double y, double x) { return atan2(y, x); }
static double MSABI st_abs(double x) { return fabs(x); }
static double MSABI st_round(double x) { return round(x); }                                   /* half away from zero, as MATLAB */
static unsigned char MSABI st_isnan(double x) { return isnan(x) ? 1 : 0; }
static double MSABI st_max(double a, double b) { return isnan(a) ? b : (isnan(b) ? a : (a > b ? a : b)); }
static double MSABI st_min(double a, double b) { return isnan(a) ? b : (isnan(b) ? a : (a < b ? a : b)); }
/* ---- C runtime ---- */
static void *MSABI st_memset(void *d, int c, size_t n) { return memset(d, c, n); }
static void *MSABI st_memcpy(void *d, const void *s, size_t n) { return memcpy(d, s, n); }
/* ---- emlrt heap ---- */
static void *MSABI st_malloc(size_t n) { return malloc(n ? n : 1); }
static void *MSABI st_calloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
static void MSABI st_free(void *p) { free(p); }
static size_t MSABI st_sizemul(size_t a, size_t b, const void *rtei, const void *sp) { (void)rtei; (void)sp; return a * b; }
/* heap reference stack: emxInit_*(..., doPush = true) registers (pointer-to-emxArray*, emxFree_*) pairs; LeaveFcn releases the
 * entries pushed since the matching EnterFcn */
typedef void MSABI (*free_fn)(void *);
#define HEAPCAP 4096
static struct { void *p; free_fn fn; } g_heap[HEAPCAP];
static int g_heap_n = 0, g_frames[256], g_frames_n = 0;
static void MSABI st_heap_enter(const void *sp) { (void)sp; if (g_frames_n < 256) g_frames[g_frames_n++] = g_heap_n; else unexpected("heap frames overflow"); }
static void MSABI st_heap_push(const void *sp, void *p, free_fn fn) { (void)sp; if (g_heap_n < HEAPCAP) { g_heap[g_heap_n].p = p; g_heap[g_heap_n++].fn = fn; } else unexpected("heap stack overflow"); }
static void MSABI st_heap_leave(const void *sp)
{
    (void)sp;
    if (!g_frames_n) { unexpected("heap leave without enter"); return; }
    const int base = g_frames[--g_frames_n];
    while (g_heap_n > base) { --g_heap_n; g_heap[g_heap_n].fn(g_heap[g_heap_n].p); }
}
/* ---- emlrt run-time checks: pass-through for valid input, recorded if they would have raised ---- */
static int MSABI st_bounds(int idx, int lo, int hi, const void *info, const void *sp)
{ (void)info; (void)sp; if (idx < lo || idx > hi) unexpected("emlrtDynamicBoundsCheckR2012b (out of range)"); return idx; }
static double MSABI st_intcheck(double d, const void *info, const void *sp)
{ (void)info; (void)sp; if (d != floor(d)) unexpected("emlrtIntegerCheckR2012b (non-integer)"); return d; }
static void MSABI st_sizeeq(const int *a, const int *b, const void *info, const void *sp)
{ (void)info; (void)sp; if (a[0] != b[0] || a[1] != b[1]) unexpected("emlrtSizeEqCheckNDR2012b (mismatch)"); }
static void MSABI st_subassign(const int *d1, int n1, const int *d2, int n2, const void *info, const void *sp)
{ (void)d1; (void)n1; (void)d2; (void)n2; (void)info; (void)sp; }
static void MSABI st_subassign1d(int a, int b, const void *info, const void *sp)
{ (void)info; (void)sp; if (a != b) unexpected("emlrtSubAssignSizeCheck1dR2017a (mismatch)"); }
static unsigned char MSABI st_breakcheck(const void *sp) { (void)sp; return 0; }

/* any other import (mx / error / MEX-gateway services) must never run when flowToColor is entered directly with valid input */
#define NGEN 100
static const char *g_gen_names[NGEN];
#define GEN(i) static uint64_t MSABI st_gen##i(void) { unexpected(g_gen_names[i] ? g_gen_names[i] : "?"); return 0; }
#define G10(t) GEN(t##0) GEN(t##1) GEN(t##2) GEN(t##3) GEN(t##4) GEN(t##5) GEN(t##6) GEN(t##7) GEN(t##8) GEN(t##9)
GEN(0) GEN(1) GEN(2) GEN(3) GEN(4) GEN(5) GEN(6) GEN(7) GEN(8) GEN(9) G10(1) G10(2) G10(3) G10(4) G10(5) G10(6) G10(7) G10(8) G10(9)
#define P(i) (void *)st_gen##i
#define P10(t) P(t##0), P(t##1), P(t##2), P(t##3), P(t##4), P(t##5), P(t##6), P(t##7), P(t##8), P(t##9)
static void *g_gen[NGEN] = {P(0), P(1), P(2), P(3), P(4), P(5), P(6), P(7), P(8), P(9), P10(1), P10(2), P10(3), P10(4), P10(5), P10(6), P10(7), P10(8), P10(9)};

static const struct { const char *name; void *fn; } g_known[] = {
    {"muDoubleScalarSqrt", (void *)st_sqrt}, {"muDoubleScalarFloor", (void *)st_floor}, {"muDoubleScalarAtan2", (void *)st_atan2},
    {"muDoubleScalarAbs", (void *)st_abs}, {"muDoubleScalarRound", (void *)st_round}, {"muDoubleScalarIsNaN", (void *)st_isnan},
    {"muDoubleScalarMax", (void *)st_max}, {"muDoubleScalarMin", (void *)st_min},
    {"memset", (void *)st_memset}, {"memcpy", (void *)st_memcpy},
    {"emlrtMallocMex", (void *)st_malloc}, {"emlrtCallocMex", (void *)st_calloc}, {"emlrtFreeMex", (void *)st_free},
    {"emlrtSizeMulR2012b", (void *)st_sizemul},
    {"emlrtHeapReferenceStackEnterFcnR2012b", (void *)st_heap_enter}, {"emlrtHeapReferenceStackLeaveFcnR2012b", (void *)st_heap_leave},
    {"emlrtPushHeapReferenceStackR2012b", (void *)st_heap_push},
    {"emlrtDynamicBoundsCheckR2012b", (void *)st_bounds}, {"emlrtIntegerCheckR2012b", (void *)st_intcheck},
    {"emlrtSizeEqCheckNDR2012b", (void *)st_sizeeq}, {"emlrtSubAssignSizeCheckR2012b", (void *)st_subassign},
    {"emlrtSubAssignSizeCheck1dR2017a", (void *)st_subassign1d}, {"emlrtBreakCheckR2012b", (void *)st_breakcheck},
};

static uint8_t *g_img = NULL;
static char g_err[256];

const char *qref_ftc_error(void) { return g_err; }
int qref_ftc_unexpected_calls(char *name, int cap) { if (name) snprintf(name, cap, "%s", g_last_unexpected); return g_unexpected; }

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

#define RVA_FLOWTOCOLOR 0x10f0          /* export "flowToColor" */
#define RVA_BREAKFLAG_PTR 0x10ba8       /* global filled from emlrtGetBreakCheckFlagAddressR2012b() by the MEX initialiser (VA 0x180001020) */

/* map `path` (a PE32+ DLL), apply its base relocations and resolve its imports with the stand-ins above */
int qref_ftc_load(const char *path)
{
    if (g_img) return 0;
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return -1; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *file = malloc(n);
    if (fread(file, 1, n, f) != (size_t)n) { fclose(f); free(file); snprintf(g_err, sizeof g_err, "short read"); return -1; }
    fclose(f);
    const uint32_t pe = rd32(file + 0x3c);
    if (memcmp(file + pe, "PE\0\0", 4) || rd16(file + pe + 24) != 0x20b) { free(file); snprintf(g_err, sizeof g_err, "not a PE32+ image"); return -1; }
    const int nsec = rd16(file + pe + 6), optsz = rd16(file + pe + 20);
    const uint8_t *opt = file + pe + 24;
    const uint64_t pref = rd64(opt + 24);
    const uint32_t imgsz = rd32(opt + 56), hdrsz = rd32(opt + 60);
    uint8_t *img = mmap(NULL, imgsz, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (img == MAP_FAILED) { free(file); snprintf(g_err, sizeof g_err, "cannot map %u bytes", imgsz); return -1; }
    memcpy(img, file, hdrsz);
    const uint8_t *sec = opt + optsz;
    for (int i = 0; i < nsec; ++i, sec += 40) {
        const uint32_t va = rd32(sec + 12), rsz = rd32(sec + 16), roff = rd32(sec + 20), vsz = rd32(sec + 8);
        memcpy(img + va, file + roff, rsz < vsz || vsz == 0 ? rsz : vsz);
    }
    /* base relocations (data directory 5): blocks of {page RVA, block size, uint16 entries type<<12 | offset}; only DIR64 (10) occurs */
    const uint64_t delta = (uint64_t)(uintptr_t)img - pref;
    const uint32_t rel = rd32(opt + 112 + 5 * 8), relsz = rd32(opt + 112 + 5 * 8 + 4);
    for (uint32_t o = 0; o + 8 <= relsz;) {
        const uint32_t page = rd32(img + rel + o), bsz = rd32(img + rel + o + 4);
        if (bsz < 8) break;
        for (uint32_t k = 8; k + 2 <= bsz; k += 2) {
            const uint16_t e = rd16(img + rel + o + k);
            const int type = e >> 12;
            if (type == 10) { uint64_t v = rd64(img + page + (e & 0xfff)) + delta; memcpy(img + page + (e & 0xfff), &v, 8); }
            else if (type != 0) { free(file); snprintf(g_err, sizeof g_err, "unsupported relocation type %d", type); return -1; }
        }
        o += bsz;
    }
    /* imports (data directory 1) */
    const uint32_t imp = rd32(opt + 112 + 8);
    int gen = 0;
    for (const uint8_t *d = img + imp; rd32(d) || rd32(d + 12) || rd32(d + 16); d += 20) {
        const uint32_t oft = rd32(d), ft = rd32(d + 16);
        for (int k = 0;; ++k) {
            const uint64_t ent = rd64(img + oft + 8 * k);
            if (!ent) break;
            const char *name = (ent >> 63) ? "(ordinal)" : (const char *)(img + (ent & 0x7fffffff) + 2);
            void *fn = NULL;
            for (size_t j = 0; j < sizeof g_known / sizeof g_known[0]; ++j)
                if (!strcmp(name, g_known[j].name)) fn = g_known[j].fn;
            if (!fn) { if (gen >= NGEN) { free(file); snprintf(g_err, sizeof g_err, "too many imports"); return -1; } g_gen_names[gen] = name; fn = g_gen[gen++]; }
            memcpy(img + ft + 8 * k, &fn, 8);
        }
    }
    /* the generated loops poll `*emlrtBreakCheckR2012bFlagVar` (Ctrl-C); we enter below the MEX initialiser, so point it at a zero byte */
    static uint8_t never_break = 0;
    if (rd64(img + RVA_BREAKFLAG_PTR) != 0) { free(file); snprintf(g_err, sizeof g_err, "unexpected image: break-check pointer slot is not empty"); return -1; }
    { void *pp = &never_break; memcpy(img + RVA_BREAKFLAG_PTR, &pp, 8); }
    free(file);
    g_img = img;
    return 0;
}

typedef void MSABI (*flowtocolor_fn)(const void *sp, const emx_t *flow, emx_t *img, emx_t *flo, double *minu, double *maxu,
                                      double *minv, double *maxv, emx_t *unknown);

static void emx_new(emx_t *e, int nd) { e->data = NULL; e->size = calloc(nd, sizeof(int32_t)); e->allocatedSize = 0; e->numDimensions = nd; e->canFreeData = 1; }
static void emx_del(emx_t *e) { if (e->canFreeData) free(e->data); free(e->size); }

/* [img, flo, minu, maxu, minv, maxv, idxUnknown] = flowToColor_mex(flow): column-major M x N x 2 doubles in; img M x N x 3 uint8,
 * flo M x N x 2, range[4] = minu,maxu,minv,maxv, unknown M x N (0/1) out.  Returns 0, -1 not loaded, -2 the binary took an
 * error path, -3 unexpected output shape. */
int qref_flow_to_color(const double *flow, int M, int N, uint8_t *img, double *flo, double *range, uint8_t *unknown)
{
    if (!g_img) return -1;
    static uint8_t tls[4096];
    const void *sp[3] = {NULL, tls, NULL};                  /* emlrtStack {site, tls, prev} */
    int32_t fsz[3] = {M, N, 2};
    emx_t in = {(void *)flow, fsz, M * N * 2, 3, 0}, eimg, eflo, eunk;
    emx_new(&eimg, 3); emx_new(&eflo, 3); emx_new(&eunk, 2);
    const int before = g_unexpected;
    flowtocolor_fn fn = (flowtocolor_fn)(uintptr_t)(g_img + RVA_FLOWTOCOLOR);
    fn(sp, &in, &eimg, &eflo, &range[0], &range[1], &range[2], &range[3], &eunk);
    int rc = g_unexpected != before ? -2 : 0;
    if (!rc && (eimg.size[0] != M || eimg.size[1] != N || eimg.size[2] != 3 || eflo.size[0] != M || eflo.size[1] != N || eflo.size[2] != 2 ||
                eunk.size[0] != M || eunk.size[1] != N)) rc = -3;
    if (!rc) {
        memcpy(img, eimg.data, (size_t)M * N * 3);
        memcpy(flo, eflo.data, (size_t)M * N * 2 * sizeof(double));
        memcpy(unknown, eunk.data, (size_t)M * N);
    }
    emx_del(&eimg); emx_del(&eflo); emx_del(&eunk);
    return rc;
}
