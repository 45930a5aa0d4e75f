// mex_runtime.cpp -- a small stand-in for MATLAB's MEX runtime (libmx/libmex), for TESTS ONLY.
//
// MATLAB is not installed where this repository is built and tested, so the gateways in ../ could only be compile-checked.
// This file implements the subset of the MEX C API declared in stub/mex.h on top of a plain C++ mxArray (class id, column-major
// dims, data, struct fields) and exposes a tiny C harness (mh_*) through which tests/test_mex_gateway.py builds the
// arguments a MATLAB caller would pass, invokes the very same mexFunction bodies, and inspects what they return.
// mexErrMsgIdAndTxt long-jumps out of MATLAB's mexFunction; here it throws, and the harness reports id + message.
// Nothing in the product links against this file.
#include "mex.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    std::vector<mwSize> dims;
    std::vector<unsigned char> data;                       // numeric / logical / char (1 byte per char) payload
    std::vector<std::string> names;                        // struct field order
    std::map<std::string, mxArray *> fields;               // 1x1 structs only (all the gateways use)
    size_t elsize() const {
        switch (cls) { case mxDOUBLE_CLASS: case mxUINT64_CLASS: return 8; case mxSTRUCT_CLASS: return 0; default: return 1; }
    }
    size_t numel() const { size_t n = 1; for (mwSize d : dims) n *= d; return dims.empty() ? 0 : n; }
};

struct MexError { std::string id, msg; };
static std::vector<void (*)(void)> g_atexit;

static mxArray *make(mxClassID cls, mwSize ndim, const mwSize *dims)
{
    mxArray *a = new mxArray_tag();
    a->cls = cls;
    a->dims.assign(dims, dims + ndim);
    if (a->dims.size() < 2) a->dims.resize(2, 1);
    a->data.assign(a->numel() * a->elsize(), 0);
    return a;
}

extern "C" {
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw MexError{id ? id : "", buf};
}
void mexWarnMsgIdAndTxt(const char *, const char *, ...) {}
int mexPrintf(const char *, ...) { return 0; }
int mexAtExit(void (*fn)(void)) { g_atexit.push_back(fn); return 0; }
void mexLock(void) {}
void mexUnlock(void) {}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) { const mwSize d[2] = {m, n}; return make(mxDOUBLE_CLASS, 2, d); }
mxArray *mxCreateDoubleScalar(double v) { mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL); *(double *)a->data.data() = v; return a; }
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity) { return make(cls, ndim, dims); }
mxArray *mxCreateLogicalArray(mwSize ndim, const mwSize *dims) { return make(mxLOGICAL_CLASS, ndim, dims); }
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names)
{
    const mwSize d[2] = {m, n};
    mxArray *a = make(mxSTRUCT_CLASS, 2, d);
    for (int i = 0; i < nfields; ++i) { a->names.push_back(names[i]); a->fields[names[i]] = nullptr; }
    return a;
}
void mxSetField(mxArray *s, mwIndex, const char *name, mxArray *v)
{
    if (!s->fields.count(name)) s->names.push_back(name);
    s->fields[name] = v;
}
mxArray *mxGetField(const mxArray *s, mwIndex, const char *name)
{
    if (!s || s->cls != mxSTRUCT_CLASS) return nullptr;
    auto it = s->fields.find(name);
    return it == s->fields.end() ? nullptr : it->second;
}
double *mxGetPr(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS ? (double *)a->data.data() : nullptr; }
void *mxGetData(const mxArray *a) { return a ? (void *)a->data.data() : nullptr; }
double mxGetScalar(const mxArray *a)
{
    if (!a || a->numel() == 0) throw MexError{"MATLAB:mxGetScalar:empty", "mxGetScalar of an empty array"};
    if (a->cls == mxDOUBLE_CLASS) return *(const double *)a->data.data();
    if (a->cls == mxUINT64_CLASS) return (double)*(const uint64_t *)a->data.data();
    return (double)a->data[0];
}
mxLogical *mxGetLogicals(const mxArray *a) { return a && a->cls == mxLOGICAL_CLASS ? (mxLogical *)a->data.data() : nullptr; }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->dims.size(); }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims.data(); }
size_t mxGetNumberOfElements(const mxArray *a) { return a->numel(); }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) { size_t n = 1; for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i]; return n; }
int mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
int mxIsComplex(const mxArray *) { return 0; }
int mxIsLogical(const mxArray *a) { return a && a->cls == mxLOGICAL_CLASS; }
int mxIsStruct(const mxArray *a) { return a && a->cls == mxSTRUCT_CLASS; }
int mxIsChar(const mxArray *a) { return a && a->cls == mxCHAR_CLASS; }
int mxIsEmpty(const mxArray *a) { return !a || a->numel() == 0; }
int mxIsUint8(const mxArray *a) { return a && a->cls == mxUINT8_CLASS; }
char *mxArrayToString(const mxArray *a)
{
    if (!a || a->cls != mxCHAR_CLASS) return nullptr;
    char *s = (char *)malloc(a->data.size() + 1);
    memcpy(s, a->data.data(), a->data.size());
    s[a->data.size()] = 0;
    return s;
}
void mxFree(void *p) { free(p); }

// the three gateways, compiled with -DmexFunction=<name> (see tests/test_mex_gateway.py)
void gqmap_mex_entry(int, mxArray *[], int, const mxArray *[]);
void get_map_mex_entry(int, mxArray *[], int, const mxArray *[]);
void flowToColor_mex_entry(int, mxArray *[], int, const mxArray *[]);

// ---- harness ------------------------------------------------------------------------------------------------------------------
void *mh_new_double(int ndim, const size_t *dims, const double *data)
{
    std::vector<mwSize> d(dims, dims + ndim);
    mxArray *a = make(mxDOUBLE_CLASS, ndim, d.data());
    if (data) memcpy(a->data.data(), data, a->data.size());
    return a;
}
void *mh_new_logical(int ndim, const size_t *dims, const unsigned char *data)
{
    std::vector<mwSize> d(dims, dims + ndim);
    mxArray *a = make(mxLOGICAL_CLASS, ndim, d.data());
    if (data) memcpy(a->data.data(), data, a->data.size());
    return a;
}
void *mh_new_string(const char *s)
{
    const mwSize d[2] = {1, strlen(s)};
    mxArray *a = make(mxCHAR_CLASS, 2, d);
    memcpy(a->data.data(), s, strlen(s));
    return a;
}
void *mh_new_struct(void) { return mxCreateStructMatrix(1, 1, 0, nullptr); }
void mh_set_field(void *s, const char *name, void *v) { mxSetField((mxArray *)s, 0, name, (mxArray *)v); }
void *mh_get_field(void *s, const char *name) { return mxGetField((const mxArray *)s, 0, name); }
int mh_class(void *a) { return (int)((mxArray *)a)->cls; }
int mh_ndim(void *a) { return (int)((mxArray *)a)->dims.size(); }
void mh_dims(void *a, size_t *out) { mxArray *x = (mxArray *)a; for (size_t i = 0; i < x->dims.size(); ++i) out[i] = x->dims[i]; }
void *mh_data(void *a) { return ((mxArray *)a)->data.data(); }
size_t mh_nbytes(void *a) { return ((mxArray *)a)->data.size(); }
void mh_free(void *a)
{
    mxArray *x = (mxArray *)a;
    if (!x) return;
    for (auto &kv : x->fields) mh_free(kv.second);
    delete x;
}
// returns 0, or 1 after mexErrMsgIdAndTxt (id/msg filled), or 2 for an unknown gateway
int mh_call(const char *gateway, int nlhs, void **plhs, int nrhs, void **prhs, char *errid, char *errmsg, int cap)
{
    void (*fn)(int, mxArray *[], int, const mxArray *[]) = nullptr;
    if (!strcmp(gateway, "gqmap_mex")) fn = gqmap_mex_entry;
    else if (!strcmp(gateway, "get_map_mex")) fn = get_map_mex_entry;
    else if (!strcmp(gateway, "flowToColor_mex")) fn = flowToColor_mex_entry;
    else return 2;
    std::vector<mxArray *> out(nlhs > 0 ? nlhs : 1, nullptr);
    try {
        fn(nlhs, out.data(), nrhs, (const mxArray **)prhs);
    } catch (const MexError &e) {
        snprintf(errid, cap, "%s", e.id.c_str());
        snprintf(errmsg, cap, "%s", e.msg.c_str());
        return 1;
    }
    for (int i = 0; i < (nlhs > 0 ? nlhs : 1); ++i) plhs[i] = out[i];
    return 0;
}
void mh_run_atexit(void) { for (auto fn : g_atexit) fn(); g_atexit.clear(); }
}
