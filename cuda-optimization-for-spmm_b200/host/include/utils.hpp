// utils.hpp -- tolerances, suffix test and the record printer.
// Same contract as the reference's include/utils.hpp:10-49: REL_TOL / ABS_TOL, endsWith(), and
// reportTime() printing one `{...},` fragment per kernel with the keys
//   testcase, sparsity, format, kernelType, denseOrdering, correct, cudaPrologTimeMs,
//   cudaKernelTimeMs, cudaEpilogTimeMs, cudaTotalTimeMs, sequentialTimeMs
// (all values quoted strings, as the reference prints them).  New keys are additive and come
// after the reference's: kernelName, nGpus, gflops, algBytes, hbmGBs, hbmFrac (of hbmPeakGBs), maxAbsErr (max |C - Cref|).
#pragma once

#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>

// global declared in main (reference: src/main.cu:17, include/utils.hpp:8)
extern std::string testcase;

#define REL_TOL 1e-2f
#define ABS_TOL 1e-3f

inline bool endsWith(const std::string &s, const std::string &suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

namespace cuspmm {
// optional extension of a record; negative numbers mean "not measured"
struct RecordExtra {
    std::string kernelName;
    int nGpus = 1;
    double gflops = -1, algBytes = -1, hbmGBs = -1, hbmFrac = -1, maxAbsErr = -1, e2eMs = -1;
    double imbalance = -1;      // multi-GPU: max panel share / mean panel share (non-zeros, slots or blocks)
};
// HBM copy bandwidth used as the roofline denominator, GB/s: --hbm-peak, else $CUSPMM_HBM_PEAK_GBS, else "hbm_gbs" of
// MEASURED_PEAKS.json ($CUSPMM_MEASURED_PEAKS, ./, ../, ../../, ../../../), else the profiling guide's fallback 6650.
double hbmPeakGBs();
const char *hbmPeakSource();
}  // namespace cuspmm

inline void reportTime(std::string tc, uint32_t aNumRows, uint32_t aNumCols, uint32_t aNumNonZero, std::string format,
                       int ordering, int kernelNum, double pro, double kernel, double epilog, bool correct,
                       const cuspmm::RecordExtra *extra = nullptr) {
    const char *ord = ordering == 0 ? "ROW_MAJOR" : "COL_MAJOR";
    const double total = pro + kernel + epilog;
    // the reference divides by a uint32 product (utils.hpp:39), which wraps above 65535^2; use doubles
    const double density = (double)aNumNonZero / ((double)aNumRows * (double)aNumCols);
    std::cout << "{\n\"testcase\":\"" << tc << "\",\n"
              << "\"sparsity\":\"" << density << "\",\n"
              << "\"format\":\"" << format << "\",\n"
              << "\"kernelType\":\"" << kernelNum << "\",\n"
              << "\"denseOrdering\":\"" << ord << "\",\n"
              << "\"correct\":\"" << correct << "\",\n";
    std::cout.flush();
    printf("\"cudaPrologTimeMs\":\"%lf\",\n\"cudaKernelTimeMs\":\"%lf\",\n\"cudaEpilogTimeMs\":\"%lf\",\n"
           "\"cudaTotalTimeMs\":\"%lf\",\n\"sequentialTimeMs\":\"%lf\"",
           pro, kernel, epilog, total, 0.0);
    if (extra) {
        printf(",\n\"kernelName\":\"%s\",\n\"nGpus\":\"%d\"", extra->kernelName.c_str(), extra->nGpus);
        if (extra->gflops >= 0) printf(",\n\"gflops\":\"%.3f\"", extra->gflops);
        if (extra->algBytes >= 0) printf(",\n\"algBytes\":\"%.0f\"", extra->algBytes);
        if (extra->hbmGBs >= 0)
            printf(",\n\"hbmGBs\":\"%.3f\",\n\"hbmFrac\":\"%.5f\",\n\"hbmPeakGBs\":\"%.1f\",\n\"hbmPeakSource\":\"%s\"", extra->hbmGBs,
                   extra->hbmFrac, cuspmm::hbmPeakGBs(), cuspmm::hbmPeakSource());
        if (extra->maxAbsErr >= 0) printf(",\n\"maxAbsErr\":\"%.3e\"", extra->maxAbsErr);
        if (extra->e2eMs >= 0) printf(",\n\"e2eTotalTimeMs\":\"%.4f\"", extra->e2eMs);
        if (extra->imbalance >= 0) printf(",\n\"panelImbalance\":\"%.4f\"", extra->imbalance);
    }
    printf("\n},\n");
    fflush(stdout);
}
