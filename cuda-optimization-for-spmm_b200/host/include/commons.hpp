// commons.hpp -- shared includes and index helpers of the host layer.
// Mirrors the role of the reference's include/commons.hpp:1-21 (RowMjIdx / ColMjIdx / assertTypes3).
#pragma once

#include <cassert>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "utils.hpp"

#define RowMjIdx(r, c, ncols) ((size_t)(r) * (size_t)(ncols) + (size_t)(c))
#define ColMjIdx(r, c, nrows) ((size_t)(c) * (size_t)(nrows) + (size_t)(r))

#define assertTypes3(DT, ta, tb, tc) \
    static_assert(std::is_same_v<DT, ta> || std::is_same_v<DT, tb> || std::is_same_v<DT, tc>, "Unsupported type")
