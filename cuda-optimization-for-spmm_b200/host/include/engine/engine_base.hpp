// engine/engine_base.hpp -- kernel-number dispatch base (reference: include/engine/engine_base.hpp:5-10)
// plus the run options the B200 engine adds (device-side timing, device ordinal, GPU count).
#pragma once

#include <string>

namespace cuspmm {

class EngineBase {
  public:
    int numKernels = 0;
    virtual ~EngineBase() = default;
    // num == 0: the reference's CPU function (in-process checker); 1..numKernels: GPU variants;
    // -1: the engine's default (selector).  Anything else throws std::runtime_error("Not implemented").
    virtual void *runKernel(int num, void *_ma, void *_mb, void *_mc) = 0;
};

// Additive knobs (the reference hard-codes device 7, one launch, host chrono: src/main.cu:176,
// src/spmm/csr/spmm_csr_k1.cu:45-73).  Defined in engine.cpp.
struct RunOptions {
    int device = 0;          // --device
    int nGpus = 1;           // --gpus: > 1 adds the row-panel multi-GPU run (every format)
    int warmup = 1;          // --warmup: untimed launches before timing
    int iters = 5;           // --iters: launches timed with CUDA events (cudaKernelTimeMs = average)
    int onlyKernel = 0;      // --variant: run just this kernel number (0 = all)
    int bsrBlock = 0;        // --bsr-block: convert the .csr file to BSR(b x b) on the device instead of reading .bsr
    bool gather = true;      // multi-GPU: store C panels straight into GPU 0's C over NVLink
    double hbmPeakGBs = 0;   // --hbm-peak: roofline denominator in GB/s (0 = MEASURED_PEAKS.json / environment / fallback)
};
extern RunOptions g_opts;

}  // namespace cuspmm
