// format.hpp -- umbrella include of the storage classes (reference: include/format.hpp:1-9).
#pragma once

#include "formats/matrix.hpp"
#include "formats/dense.hpp"
#include "formats/sparse_csr.hpp"
#include "formats/sparse_coo.hpp"
#include "formats/sparse_ell.hpp"
#include "formats/sparse_bsr.hpp"
