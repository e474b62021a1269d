// engine.hpp -- runEngine (reference: include/engine.hpp:12-13, src/engine/engine.cpp:17-61).
#pragma once

#include "engine/cusparse.hpp"
#include "engine/engine_base.hpp"
#include "engine/engine_bsr.hpp"
#include "engine/engine_coo.hpp"
#include "engine/engine_csr.hpp"
#include "engine/engine_ell.hpp"

namespace cuspmm {

// H2D of A and B, the CPU function (unless skipSeq), every GPU kernel of the engine checked against
// it, cuSPARSE when the format supports it, and (CSR, --gpus N) the multi-GPU row-panel run.
template <typename EngT>
void runEngine(EngT *engine, typename EngT::MataT *a, typename EngT::MatbT *b, float abs_tol, float rel_tol, bool skipSeq);

}  // namespace cuspmm
