// spmm_cusparse.hpp -- cuSPARSE include + status check (reference: include/spmm_cusparse.hpp:1-12).
#pragma once

#include <cstdio>
#include <cstdlib>

#include <cusparse.h>

#define CHECK_CUSPARSE(func)                                                                           \
    {                                                                                                  \
        cusparseStatus_t status_ = (func);                                                             \
        if (status_ != CUSPARSE_STATUS_SUCCESS) {                                                      \
            std::printf("CUSPARSE API failed at line %d with error: %s (%d)\n", __LINE__,             \
                        cusparseGetErrorString(status_), status_);                                     \
            std::exit(EXIT_FAILURE);                                                                   \
        }                                                                                              \
    }
