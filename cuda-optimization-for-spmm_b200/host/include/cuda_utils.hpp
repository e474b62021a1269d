// cuda_utils.hpp -- CUDA error handling and the allclose check of the host layer.
// Same behaviour as the reference's include/cuda_utils.hpp:13-22 (print + exit(code)); the
// libtorch dependency (toTorch + torch::allclose, :27-41) is replaced by allClose() below, which
// applies torch's formula |a - b| <= atol + rtol * |b| element-wise.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#include "cuspmm_b200.h"

#define cudaCheckError(ans) cudaAssert((ans), __FILE__, __LINE__)
inline void cudaAssert(cudaError_t code, const char *file, int line, bool abort = true) {
    if (code != cudaSuccess) {
        fprintf(stderr, "CUDA Error: %s at %s:%d\n", cudaGetErrorString(code), file, line);
        if (abort) exit(code);
    }
}

// C-ABI status check: the engine has no CPU fallback, a failing kernel call ends the process
#define cuspmmCheck(ans) cuspmmAssert((ans), __FILE__, __LINE__)
inline void cuspmmAssert(int status, const char *file, int line) {
    if (status != CUSPMM_OK) {
        fprintf(stderr, "cuspmm error %d: %s at %s:%d\n", status, cuspmm_last_error(), file, line);
        exit(status);
    }
}

namespace cuspmm {
template <typename DT>
inline bool allClose(const DT *a, const DT *b, size_t n, double rtol, double atol) {
    for (size_t i = 0; i < n; ++i) {
        const double d = std::fabs((double)a[i] - (double)b[i]);
        if (!(d <= atol + rtol * std::fabs((double)b[i]))) return false;
    }
    return true;
}
}  // namespace cuspmm
