// qgmap_guard.h -- exception barrier of the C ABI (plain C++, no CUDA: also included by the host-only translation unit).
#pragma once
#include "../../include/qgmap.h"
void qgmap_set_last_error(const char *msg);
// include/qgmap.h promises that no C++ exception crosses the C ABI: entry points that allocate host memory (handles, std::vector
// staging, std::string) run inside this guard, which turns std::bad_alloc into QGMAP_ERR_NOMEM and anything else into QGMAP_ERR_ARG.
#include <exception>
#include <new>
template <class F> static inline int qg_guard(F &&body)
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        qgmap_set_last_error("out of host memory");
        return QGMAP_ERR_NOMEM;
    } catch (const std::exception &e) {
        qgmap_set_last_error(e.what());
        return QGMAP_ERR_ARG;
    } catch (...) {
        qgmap_set_last_error("unexpected C++ exception");
        return QGMAP_ERR_ARG;
    }
}
