// pqperm_limits.h -- compile-time limits shared by host planning and kernels.
#pragma once

#include <cstdint>
#include <cstdlib>

namespace pqperm {

constexpr int kMaxCols = 64;            // register-resident row sums per thread
constexpr int kLapMaxCols = 256;        // lane-split walk: 32 lanes x 8 columns (wide problems)
constexpr int kMaxDigits = 64;          // Gray digits (active rows after the split)
constexpr int kMaxLowDigits = 24;       // digits walked inside one segment
constexpr int kMaxMultiplicity = 254;   // radix r+1 is stored in a byte
constexpr int kBinMinCols = 8;          // narrowest binary constant-bank kernel
constexpr int kBinMaxCols = 48;         // widest binary constant-bank kernel
constexpr int kBinMaxParamCols = 44;     // widest binary matrix that rides in the kernel parameter block
constexpr int kBinMaxBlockExp = 4;      // kernel 2: at most 2^4 terms per block (instantiated for C <= 32)
// kernel 2: log2 of the terms per block for nc columns (register budget:
// 4*nc for the row sums + 4 * 2^B for the running products)
inline int binary_block_exponent(int /*nc*/) { return 3; }  // measured best for 20 <= nc <= 48
// Below 2^kBinMinDigitsAuto terms the generic walk is used (B200 sweep over the
// segment lengths, tools/r2_sweep.py, best kernel times generic / binary:
// n = 20: 20.5 / 23.3 us, n = 22: 43.5 / 44.0, n = 24: 114.6 / 110.7, n = 25: 244 / 209,
// n = 26: 450 / 364).
constexpr int kBinMinDigitsAuto = 23;
constexpr int64_t kMaxSegLenBinary = INT64_C(1) << 14;
constexpr int64_t kMaxSegLenNary = INT64_C(1) << 10;  // step tables live in shared memory (9 KB)

// Laplace: terms per segment (step tables in shared memory, 16 bytes per term).  B200,
// batches of 2000 k-column problems, 256 -> 512 -> 1024 terms: k = 24 197.7 -> 188.0 -> 186.6 ms,
// k = 25 409.0 -> 402.2 -> 398.2 ms; config 4 (10^4 shots) kernels 1.285 -> 1.257 s at 512.
#ifndef PQ_LAP_MAXSEG
#define PQ_LAP_MAXSEG 512
#endif
constexpr int kLapMaxSegLen = PQ_LAP_MAXSEG;
// static shared memory of the batched walks: the two step tables and a few words
constexpr size_t kLapStaticSmem = (size_t)kLapMaxSegLen * 16 + 64;
#ifndef PQ_LAP_THREADS
#define PQ_LAP_THREADS 128
#endif
constexpr int kLapThreads = PQ_LAP_THREADS;
// Dynamic shared memory of one CTA of the Laplace walk: the (D+1) x NCP complex
// matrix, then every thread's double-double totals (4 doubles per column of the
// lane plus 4 for the full product).  Must fit the 227 KB of a B200 CTA beside the
// static tables (and 1 KB the driver reserves).
constexpr size_t kLapSmemLimit = 227 * 1024 - kLapStaticSmem - 1024;
// (perm_only: batched permanents keep only the full product's four slots)
inline size_t lap_smem_bytes(int D, int S, int NCL, bool perm_only = false)
{
    return (size_t)(D + 1) * S * NCL * 16 +
           (size_t)(perm_only ? 4 : 4 * NCL + 4) * kLapThreads * 8;
}

// Lane split of the Laplace walk for nc active columns: S lanes per Gray
// segment, NCL columns per lane (S * NCL >= nc).
struct LapVariant {
    int S;
    int NCL;
};
// experiments: PQ_LAP_S4_FROM=n uses four lanes per segment from n columns on
inline int lap_s4_from()
{
    static const int v = [] {
        const char *e = std::getenv("PQ_LAP_S4_FROM");
        return e ? std::atoi(e) : 27;
    }();
    return v < 17 ? 17 : v;
}
inline LapVariant laplace_variant(int nc)
{
    if (nc <= 8)
        return {1, nc < 1 ? 1 : nc};
    if (nc <= 26 && nc < lap_s4_from())
        return {2, (nc + 1) / 2};
    if (nc <= kMaxCols)
        return {4, (nc + 3) / 4};
    // wide problems (few rows, many columns): a whole warp per Gray segment
    return {32, (nc + 31) / 32};
}
// Batched permanents (full product only): no per-column accumulators, so one lane
// holds all the row sums of up to kPermS1MaxCols columns -- no lane exchange, no
// duplicated step bookkeeping (B200: n = 20 batches 11.0 -> see DESIGN.md).
constexpr int kPermS1MaxCols = 32;
// ... and, with unit columns and at least kHyperDigits rows of multiplicity 1, the
// hypercube flavour (pqperm_permhyper.cuh) from kHyperMinCols columns on
constexpr int kHyperDigits = 3;
constexpr int kHyperMinCols = 4;
inline LapVariant perm_variant(int nc)
{
    if (nc <= kPermS1MaxCols)
        return {1, nc < 1 ? 1 : nc};
    return laplace_variant(nc);
}

} // namespace pqperm
