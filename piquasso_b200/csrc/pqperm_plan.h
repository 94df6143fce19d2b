// pqperm_plan.h -- host-side preprocessing and partition planning (no CUDA).
//
// Restates src/permanent.cpp:54-142 of the reference (row split, sum check,
// trivial cases, limits, idx_max) and then compacts the problem for the GPU:
// rows and columns with multiplicity 0 are dropped (they contribute a radix-1
// digit that never moves, resp. a factor s^0 = 1), which leaves the offset ->
// Gray-code map of the remaining digits unchanged.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pqperm.h"

namespace pqperm {

struct Plan {
    // reference-level facts
    int trivial = 0;          // 1: reference early-out, value in triv
    double triv[2] = {1.0, 0.0};
    int sum_rows = 0;
    int ref_digits = 0;       // digits of the reference counter (= rows after the split - 1)
    int64_t idx_max = 1;

    // compacted problem
    int D = 0;                // digits with radix >= 2
    int NC = 0;               // columns the kernels see (after the expansion below)
    int NC_active = 0;        // columns of the caller's matrix with multiplicity > 0
    int NCP = 0;              // columns the chosen kernel is instantiated for
    int M = 0;                // sum of column multiplicities
    bool binary = false;      // every radix is 2
    bool unitcols = false;    // every column multiplicity is 1
    std::vector<int> ref_digit_of;  // compact digit -> reference digit index
    std::vector<int> mult;          // r_d
    std::vector<int> colmult;       // c_j, padded with 1 up to NCP
    std::vector<int> src_row;       // matrix row feeding digit d (index into A)
    std::vector<int> src_col;       // matrix column feeding compact column j
    int pinned_row = 0;             // matrix row used as the delta=+1 row
    std::vector<double> A2;         // (D+1) x NCP interleaved, rows 1..D doubled

    // partition
    int q = 0;                // low digits walked inside a segment
    int64_t W = 1;            // terms per segment
    int64_t nseg = 1;         // segments
    int kernel = 1;           // 1 generic, 2 binary constant-bank
    int B = 0;                // kernel 2: 2^B terms (one hypercube of the low digits) per block
    std::vector<double> binom;      // flattened C(r_d, g)
    std::vector<int> binom_off;     // [D]
};

struct PlanOptions {
    int kernel_choice = 0;    // 0 auto, 1 generic, 2 binary; 2 + 10*B forces the block exponent
    int64_t seg_len_hint = 0; // 0 auto
    int num_sms = 148;
};

// Builds the plan.  `A` may be null (structure only: no A2).  Returns PQ_OK or
// an error code with `err` filled.
int make_plan(const double *A, int R, int C, const int32_t *rows, const int32_t *cols,
              const PlanOptions &opt, Plan &plan, std::string &err);

// Gray digits, in REFERENCE digit order (plan.ref_digits entries), of `offset`
// as the device seeding + walk assigns them.
void plan_gray_of_offset(const Plan &plan, int64_t offset, int32_t *gray);

} // namespace pqperm
