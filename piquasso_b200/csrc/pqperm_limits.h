// pqperm_limits.h -- compile-time limits shared by host planning and kernels.
#pragma once

#include <cstdint>

namespace pqperm {

constexpr int kMaxCols = 64;            // register-resident row sums per thread
constexpr int kMaxDigits = 64;          // Gray digits (active rows after the split)
constexpr int kMaxLowDigits = 24;       // digits walked inside one segment
constexpr int kMaxMultiplicity = 254;   // radix r+1 is stored in a byte
constexpr int kBinMinCols = 8;          // narrowest binary constant-bank kernel
constexpr int kBinMaxCols = 48;         // widest binary constant-bank kernel
constexpr int kBinDefaultUnroll = 2;    // log2 of the unrolled inner block
constexpr int kBinDefaultChains = 2;    // independent product chains
constexpr int64_t kMaxSegLenBinary = INT64_C(1) << 14;
constexpr int64_t kMaxSegLenNary = INT64_C(1) << 12;

} // namespace pqperm
