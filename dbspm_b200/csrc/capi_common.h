// capi_common.h -- error plumbing shared by the translation units of libdbspm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/dbspm_b200.h"

void dbspm_set_error(const char *fmt, ...);

#define DBSPM_CUDA_TRY(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dbspm_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                            cudaGetErrorString(_e));                                           \
            return DBSPM_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

#define DBSPM_REQUIRE(cond, code, ...)                                                         \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            dbspm_set_error(__VA_ARGS__);                                                      \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)
