// main.cpp -- the `cuspmm` command line.
// Flags, file discovery by suffix, error messages and exit codes follow the reference's src/main.cu:
//   cuspmm --bsr --coo --csr --ell [--cuda] -d <dir> [-h]                       (main.cu:19-29,36-82)
//   files: *.coo *.csr *.bsr *_colind.ell *_values.ell *_rowind.ell *_values_colmajor.ell dense.in  (:98-144)
// Additive flags (defaults keep the reference behaviour except the device ordinal, which the
// reference hard-codes to 7, main.cu:176):
//   --device D  --gpus N  --iters I  --warmup W  --variant K  --skip-cpu  --bsr-block B  --no-gather  --hbm-peak GB/s
#include "engine.hpp"
#include "format.hpp"

#include <dirent.h>
#include <getopt.h>
#include <sys/stat.h>

#include <algorithm>
#include <vector>

std::string testcase;

static void printHelp(const char *prog) {
    std::cout << "Usage: " << prog << " [OPTIONS]\n"
              << "Options:\n"
              << "  --bsr           Process data in Block Sparse Row format\n"
              << "  --coo           Process data in Coordinate format\n"
              << "  --csr           Process data in Compressed Sparse Row format\n"
              << "  --ell           Process data in ELLPACK format\n"
              << "  --cuda          Enable CUDA processing\n"
              << "  -d <directory>  Data directory\n"
              << "  -h, --help      Display this help message\n"
              << "B200 engine additions:\n"
              << "  --device <n>    CUDA device ordinal (default 0)\n"
              << "  --gpus <n>      also run every selected format sharded by balanced row panels over n GPUs\n"
              << "  --no-gather     multi-GPU: leave C sharded instead of storing it into GPU 0 over NVLink\n"
              << "  --iters <n>     timed launches per kernel (CUDA events, default 5)\n"
              << "  --warmup <n>    untimed launches per kernel (default 1)\n"
              << "  --variant <k>   run only GPU kernel number k of each engine\n"
              << "  --skip-cpu      skip kernel 0 (records then compare against zeros, as the reference's skipSeq)\n"
              << "  --bsr-block <b> build BSR(b x b) from the .csr file on the device instead of reading .bsr\n"
              << "  --no-gather     multi-GPU: leave C sharded instead of storing the panels into GPU 0's C over NVLink\n"
              << "  --hbm-peak <x>  HBM bandwidth in GB/s for the roofline fraction (default: MEASURED_PEAKS.json, else 6650)\n";
}

int main(int argc, char *argv[]) {
    std::string dir;
    bool TEST_COO = false, TEST_CSR = false, TEST_BSR = false, TEST_ELL = false, skipCpu = false;
    const struct option longOpts[] = {
        {"bsr", no_argument, nullptr, 1},         {"coo", no_argument, nullptr, 2},
        {"csr", no_argument, nullptr, 3},         {"ell", no_argument, nullptr, 4},
        {"cuda", no_argument, nullptr, 5},        {"help", no_argument, nullptr, 'h'},
        {"device", required_argument, nullptr, 10}, {"gpus", required_argument, nullptr, 11},
        {"iters", required_argument, nullptr, 12},  {"warmup", required_argument, nullptr, 13},
        {"variant", required_argument, nullptr, 14}, {"skip-cpu", no_argument, nullptr, 15},
        {"bsr-block", required_argument, nullptr, 16}, {"no-gather", no_argument, nullptr, 17},
        {"hbm-peak", required_argument, nullptr, 18},
        {nullptr, 0, nullptr, 0}};
    int opt;
    while ((opt = getopt_long(argc, argv, "hd:", longOpts, nullptr)) != -1) {
        switch (opt) {
        case 1: TEST_BSR = true; break;
        case 2: TEST_COO = true; break;
        case 3: TEST_CSR = true; break;
        case 4: TEST_ELL = true; break;
        case 5: break;   // --cuda is accepted and ignored, as in the reference (main.cu:64-67)
        case 10: cuspmm::g_opts.device = atoi(optarg); break;
        case 11: cuspmm::g_opts.nGpus = std::max(1, atoi(optarg)); break;
        case 12: cuspmm::g_opts.iters = std::max(1, atoi(optarg)); break;
        case 13: cuspmm::g_opts.warmup = std::max(0, atoi(optarg)); break;
        case 14: cuspmm::g_opts.onlyKernel = atoi(optarg); break;
        case 15: skipCpu = true; break;
        case 16: cuspmm::g_opts.bsrBlock = atoi(optarg); break;
        case 17: cuspmm::g_opts.gather = false; break;
        case 18: cuspmm::g_opts.hbmPeakGBs = atof(optarg); break;
        case 'h': printHelp(argv[0]); return 0;
        case 'd': dir = optarg; break;
        case '?': return 1;
        default: break;
        }
    }
    if (dir.empty() || (!TEST_COO && !TEST_CSR && !TEST_BSR && !TEST_ELL)) {
        printHelp(argv[0]);
        exit(EXIT_FAILURE);
    }

    // file discovery by suffix (first match wins)
    std::string coo, csr, bsr, dense, ellColind, ellValues, ellRowind, ellValuesCol;
    DIR *d = opendir(dir.c_str());
    if (!d) {
        std::cerr << "Error: cannot open directory " << dir << "\n";
        exit(EXIT_FAILURE);
    }
    std::vector<std::string> names;
    while (dirent *e = readdir(d)) names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string &n : names) {
        const std::string path = dir + (dir.back() == '/' ? "" : "/") + n;
        struct stat st;
        if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        if (endsWith(n, ".coo") && coo.empty()) coo = path;
        else if (endsWith(n, ".csr") && csr.empty()) csr = path;
        else if (endsWith(n, ".bsr") && bsr.empty()) bsr = path;
        else if (endsWith(n, "_colind.ell") && ellColind.empty()) ellColind = path;
        else if (endsWith(n, "_values_colmajor.ell") && ellValuesCol.empty()) ellValuesCol = path;
        else if (endsWith(n, "_values.ell") && ellValues.empty()) ellValues = path;
        else if (endsWith(n, "_rowind.ell") && ellRowind.empty()) ellRowind = path;
        else if (endsWith(n, "dense.in") && dense.empty()) dense = path;
    }
    const bool bsrFromCsr = TEST_BSR && cuspmm::g_opts.bsrBlock > 0;
    if (TEST_COO && coo.empty()) { std::cerr << "Error: Missing required files *.coo in " << dir << "\n"; exit(EXIT_FAILURE); }
    if ((TEST_CSR || bsrFromCsr) && csr.empty()) { std::cerr << "Error: Missing required files *.csr in " << dir << "\n"; exit(EXIT_FAILURE); }
    if (TEST_BSR && !bsrFromCsr && bsr.empty()) { std::cerr << "Error: Missing required files *.bsr in " << dir << "\n"; exit(EXIT_FAILURE); }
    if (TEST_ELL && (ellColind.empty() || ellValues.empty())) {
        std::cerr << "Error: Missing required files *_colind.ell and/or *_values.ell in " << dir << "\n";
        exit(EXIT_FAILURE);
    }
    if (TEST_ELL && (ellRowind.empty() || ellValuesCol.empty())) {
        std::cerr << "Error: Missing required files *_rowind.ell and/or *_values_colmajor.ell in " << dir << "\n";
        exit(EXIT_FAILURE);
    }
    if (dense.empty()) { std::cerr << "Error: Missing required file dense.in in " << dir << "\n"; exit(EXIT_FAILURE); }

    cudaCheckError(cudaSetDevice(cuspmm::g_opts.device));
    testcase = dir;
    const float abs_tol = 1.0e-3f, rel_tol = 1.0e-2f;
    using namespace cuspmm;
    auto *b = new DenseMatrix<float, uint32_t>(dense);

    if (TEST_COO) {
        auto *a = new SparseMatrixCOO<float, uint32_t>(coo);
        auto *engine = new EngineCOO<float, uint32_t, double>(dir);
        runEngine(engine, a, b, abs_tol, rel_tol, skipCpu);
        delete engine; delete a;
    }
    if (TEST_CSR) {
        auto *a = new SparseMatrixCSR<float, uint32_t>(csr);
        auto *engine = new EngineCSR<float, uint32_t, double>(dir);
        runEngine(engine, a, b, abs_tol, rel_tol, skipCpu);
        delete engine; delete a;
    }
    if (TEST_BSR) {
        SparseMatrixBSR<float, uint32_t> *a = nullptr;
        DenseMatrix<float, uint32_t> *bb = b;
        if (bsrFromCsr) {   // device conversion, with zero padding of M and K (and of B's rows) to block multiples
            auto *hc = new SparseMatrixCSR<float, uint32_t>(csr);
            auto *dc = hc->copy2Device();
            auto *dbsr = SparseMatrixBSR<float, uint32_t>::fromCSR(dc, g_opts.bsrBlock, g_opts.bsrBlock);
            a = dbsr->copy2Host();
            a->numNonZero = hc->numNonZero;
            if (a->numCols != b->numRows) {
                bb = new DenseMatrix<float, uint32_t>(a->numCols, b->numCols, false);
                std::memcpy(bb->data, b->data, b->numElements() * sizeof(float));
            }
            delete dbsr; delete dc; delete hc;
        } else {
            a = new SparseMatrixBSR<float, uint32_t>(bsr);
        }
        auto *engine = new EngineBSR<float, uint32_t, double>(dir);
        runEngine(engine, a, bb, abs_tol, rel_tol, skipCpu);
        if (bb != b) delete bb;
        delete engine; delete a;
    }
    if (TEST_ELL) {
        auto *a = new SparseMatrixELL<float, uint32_t>(ellRowind, ellValuesCol);
        auto *engine = new EngineELL<float, uint32_t, double>(dir);
        runEngine(engine, a, b, abs_tol, rel_tol, skipCpu);
        delete engine; delete a;
    }
    delete b;
    return 0;
}
