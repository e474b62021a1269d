// engine/cusparse.hpp -- same-run vendor baseline (reference: include/engine/cusparse.hpp:4-5).
#pragma once

#include "formats/dense.hpp"
#include "formats/matrix.hpp"

namespace cuspmm {
// pro / kernel / epi in microseconds, as in the reference (src/engine/cusparse.cu:44-54); `kernel` is
// the CUDA-event average of g_opts.iters launches, handle/descriptor/buffer/preprocess land in `pro`.
template <typename DT, typename MT>
DenseMatrix<DT, MT> *cusparseTest(SparseMatrix<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *c, long &pro,
                                  long &kernel, long &epi);
}  // namespace cuspmm
