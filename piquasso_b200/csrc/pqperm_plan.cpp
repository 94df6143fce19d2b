// pqperm_plan.cpp -- see pqperm_plan.h.
#include "pqperm_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "pqperm_limits.h"

namespace pqperm {

static double binom_d(int n, int k)
{
    if (k < 0 || k > n)
        return 0.0;
    if (k > n - k)
        k = n - k;
    double r = 1.0;
    for (int i = 1; i <= k; i++)
        r = r * (double)(n - k + i) / (double)i; // exact while the value < 2^53
    return std::nearbyint(r);
}

int make_plan(const double *A, int R, int C, const int32_t *rows, const int32_t *cols,
              const PlanOptions &opt, Plan &plan, std::string &err)
{
    plan = Plan();
    if (R < 0 || C < 0 || (R > 0 && !rows) || (C > 0 && !cols)) {
        err = "negative shape or null multiplicity vector";
        return PQ_ERR_BAD_ARG;
    }
    int64_t sum_rows = 0, sum_cols = 0;
    for (int i = 0; i < R; i++) {
        if (rows[i] < 0) {
            err = "negative row multiplicity";
            return PQ_ERR_BAD_ARG;
        }
        sum_rows += rows[i];
    }
    for (int j = 0; j < C; j++) {
        if (cols[j] < 0) {
            err = "negative column multiplicity";
            return PQ_ERR_BAD_ARG;
        }
        sum_cols += cols[j];
    }
    plan.sum_rows = (int)sum_rows;

    // src/permanent.cpp:97-108 (the split does not change sum(rows))
    if (sum_rows != sum_cols) {
        err = "Number of input and output states should be equal";
        return PQ_ERR_SUM_MISMATCH;
    }
    if (R == 0 || C == 0 || sum_rows == 0 || sum_cols == 0) {
        plan.trivial = 1;
        return PQ_OK;
    }

    // src/permanent.cpp:54-64: first index attaining the smallest non-zero
    // multiplicity.
    int min_idx = 0, minelem = 0;
    for (int i = 0; i < R; i++) {
        if (minelem == 0 || (rows[i] < minelem && rows[i] != 0)) {
            minelem = rows[i];
            min_idx = i;
        }
    }
    // sum_rows > 0 here, so minelem != 0 and the split always happens (:66-95)
    plan.pinned_row = min_idx;
    plan.ref_digits = R; // rows after the split = R + 1, digits = R

    plan.idx_max = 1;
    for (int i = 0; i < R; i++) {
        const int r = rows[i] - (i == min_idx ? 1 : 0);
        if (r == 0)
            continue;
        if (r > kMaxMultiplicity) {
            err = "row multiplicity above " + std::to_string(kMaxMultiplicity);
            return PQ_ERR_TOO_LARGE;
        }
        plan.ref_digit_of.push_back(i);
        plan.mult.push_back(r);
        plan.src_row.push_back(i);
        if (plan.idx_max > (INT64_C(1) << 62) / (r + 1)) {
            err = "term space exceeds 2^62";
            return PQ_ERR_TOO_LARGE;
        }
        plan.idx_max *= (r + 1);
    }
    plan.D = (int)plan.mult.size();
    if (plan.D > PQ_MAX_DIGITS) {
        err = "more than 64 rows with non-zero multiplicity";
        return PQ_ERR_TOO_LARGE;
    }
    plan.binary = true;
    for (int r : plan.mult)
        if (r != 1)
            plan.binary = false;

    plan.unitcols = true;
    for (int j = 0; j < C; j++) {
        if (cols[j] == 0)
            continue;
        if (cols[j] > kMaxMultiplicity) {
            err = "column multiplicity above " + std::to_string(kMaxMultiplicity);
            return PQ_ERR_TOO_LARGE;
        }
        plan.src_col.push_back(j);
        plan.colmult.push_back(cols[j]);
        plan.M += cols[j];
        if (cols[j] != 1)
            plan.unitcols = false;
    }
    plan.NC = (int)plan.colmult.size();
    if (plan.NC > PQ_MAX_COLS) {
        err = "more than 64 columns with non-zero multiplicity";
        return PQ_ERR_TOO_LARGE;
    }
    plan.NC_active = plan.NC;
    // A column of multiplicity c is written out as c unit columns while the
    // expanded width still fits the registers: s_j^c becomes c factors of an
    // unrolled product with independent chains instead of a run-time loop on one
    // dependent chain, at the price of 2(c-1) more row-sum FMAs per term.  The
    // terms -- hence the enumeration and the partition -- do not change.
    if (!plan.unitcols && plan.M <= PQ_MAX_COLS) {
        std::vector<int> src;
        src.reserve((size_t)plan.M);
        for (int j = 0; j < plan.NC; j++)
            for (int k = 0; k < plan.colmult[j]; k++)
                src.push_back(plan.src_col[j]);
        plan.src_col.swap(src);
        plan.colmult.assign((size_t)plan.M, 1);
        plan.NC = plan.M;
        plan.unitcols = true;
    }

    // ---- kernel choice ------------------------------------------------------
    int choice = opt.kernel_choice;
    int forced_B = 0;
    if (choice >= 10) {
        forced_B = (choice / 10) % 10;
        choice = choice % 10;
    }
    const bool bin_ok = plan.binary && plan.unitcols &&
                        plan.NC >= kBinMinCols && plan.NC <= kBinMaxCols;
    plan.kernel = 1;
    if (bin_ok && choice != 1) {
        plan.B = forced_B ? std::min(std::max(forced_B, 2), plan.NC <= 32 ? kBinMaxBlockExp : 3)
                          : binary_block_exponent(plan.NC);
        // automatic choice: only where the term space is large enough to repay the
        // constant-bank upload; forced (tests, tuning): wherever a block fits
        if (plan.D >= (choice == 2 ? plan.B : kBinMinDigitsAuto))
            plan.kernel = 2;
    }
    plan.NCP = plan.kernel == 2 ? plan.NC : (plan.NC <= 4 ? 4 : (plan.NC + 3) / 4 * 4);
    plan.colmult.resize(plan.NCP, 1);

    // ---- where to cut the digits --------------------------------------------
    // cost(q) ~ waves(q) * (seed + W(q) * step); see DESIGN.md "partition".
    // Threads resident on the device: registers per thread as ptxas reports them
    // (profiles/r02_ptxas_v.txt), CTA sizes as the launchers use them.
    double regs, nt;
    if (plan.kernel == 2) {
        nt = 64.0;
        regs = plan.NC <= 16 ? 128.0 : (plan.NC <= 26 ? 168.0 : 255.0);
    } else {
        nt = plan.NCP <= 16 ? 128.0 : 64.0;
        regs = plan.NCP <= 4 ? 68.0
               : plan.NCP <= 8 ? 102.0
               : plan.NCP <= 16 ? 128.0
               : plan.NCP <= 24 ? 168.0
               : (plan.NCP <= 28 && plan.unitcols) ? 168.0 : 255.0;
    }
    const double ctas = std::max(1.0, std::min(32.0, std::floor(65536.0 / (nt * regs))));
    const double resident = std::min(2048.0, ctas * nt) * opt.num_sms;
    const double step = 2.0 * plan.NCP + 4.0 * plan.M + 2.0;
    // a seed is latency-bound (row after row through the constant bank / shared
    // memory), not FMA-bound: weight from the B200 sweep tools/r2_sweep.py
    const double seed = (plan.kernel == 2 ? 2.5 : 1.0) *
                        (2.0 * plan.NCP * (plan.D + 1) + 40.0 * plan.D + 100.0);
    const int64_t wmax = (plan.binary && plan.unitcols) ? kMaxSegLenBinary : kMaxSegLenNary;
    const int qmin = plan.kernel == 2 ? plan.B : 0;
    int best_q = -1;
    double best_cost = 0.0;
    {
        int64_t W = 1;
        for (int q = 0; q <= plan.D; q++) {
            if (q > 0)
                W *= (plan.mult[q - 1] + 1);
            if (q > kMaxLowDigits || W > wmax)
                break;
            if (q < qmin)
                continue;
            if (opt.seg_len_hint > 0) {
                // tests: the longest admissible segment not above the hint
                if (best_q < 0 || W <= opt.seg_len_hint)
                    best_q = q;
                continue;
            }
            // a partly filled last wave costs a whole one: the dispenser hands out
            // 32 segments per warp and every warp ends up with ceil(.) batches
            const double nseg = (double)(plan.idx_max / W);
            const double waves = std::ceil(nseg / resident);
            const double cost = waves * (seed + (double)W * step);
            if (best_q < 0 || cost < best_cost * 0.999) {
                best_q = q;
                best_cost = cost;
            }
        }
    }
    if (best_q < 0) {
        // cannot honour qmin (tiny problem): fall back to the generic walk
        plan.kernel = 1;
        plan.NCP = plan.NC <= 4 ? 4 : (plan.NC + 3) / 4 * 4;
        plan.colmult.resize(plan.NCP, 1);
        best_q = 0;
    }
    plan.q = best_q;
    plan.W = 1;
    for (int d = 0; d < plan.q; d++)
        plan.W *= (plan.mult[d] + 1);
    plan.nseg = plan.idx_max / plan.W;

    // ---- tables --------------------------------------------------------------
    plan.binom_off.resize(plan.D);
    for (int d = 0; d < plan.D; d++) {
        plan.binom_off[d] = (int)plan.binom.size();
        for (int g = 0; g <= plan.mult[d]; g++)
            plan.binom.push_back(binom_d(plan.mult[d], g));
    }
    // (the step tables of the low counter -- which digit moves into local index m
    // and the weight (-1)^m prod_{d<q} C(r_d, c_d(m)) -- are built on the device by
    // every CTA, pqperm_walk.cuh)

    // ---- compacted, pre-doubled matrix ----------------------------------------
    if (A) {
        plan.A2.assign((size_t)(plan.D + 1) * plan.NCP * 2, 0.0);
        for (int j = 0; j < plan.NCP; j++) {
            double re = 1.0, im = 0.0; // padding column: s_j == 1 for every term
            if (j < plan.NC) {
                const size_t src = ((size_t)plan.pinned_row * C + plan.src_col[j]) * 2;
                re = A[src];
                im = A[src + 1];
            }
            plan.A2[(size_t)j * 2] = re;
            plan.A2[(size_t)j * 2 + 1] = im;
        }
        for (int d = 0; d < plan.D; d++)
            for (int j = 0; j < plan.NC; j++) {
                const size_t src = ((size_t)plan.src_row[d] * C + plan.src_col[j]) * 2;
                const size_t dst = ((size_t)(d + 1) * plan.NCP + j) * 2;
                plan.A2[dst] = 2.0 * A[src]; // src/permanent.cpp:124-128
                plan.A2[dst + 1] = 2.0 * A[src + 1];
            }
    }
    return PQ_OK;
}

// Host mirror of the device bookkeeping (seed_segment + dirmask walk in
// pqperm_walk.cuh): offset = seg * W + m.
void plan_gray_of_offset(const Plan &plan, int64_t offset, int32_t *gray)
{
    for (int i = 0; i < plan.ref_digits; i++)
        gray[i] = 0;
    if (plan.trivial || plan.D == 0)
        return;
    const int64_t seg = offset / plan.W;
    const int64_t mloc = offset % plan.W;
    std::vector<int> g(plan.D, 0), chain(plan.D, 0);
    int64_t rest = seg;
    for (int d = plan.q; d < plan.D; d++) {
        chain[d] = (int)(rest % (plan.mult[d] + 1));
        rest /= (plan.mult[d] + 1);
    }
    int odd = 0;
    for (int d = plan.D - 1; d >= plan.q; d--) {
        g[d] = odd ? plan.mult[d] - chain[d] : chain[d];
        odd ^= (g[d] & 1);
    }
    unsigned dirmask = 0;
    for (int d = plan.q - 1; d >= 0; d--) {
        dirmask |= (unsigned)odd << d;
        g[d] = odd ? plan.mult[d] : 0;
        if (plan.mult[d] & 1)
            odd = 0;
    }
    std::vector<int> low(std::max(plan.q, 1), 0);
    for (int64_t m = 1; m <= mloc; m++) {
        int p = 0;
        while (low[p] == plan.mult[p]) {
            low[p] = 0;
            p++;
        }
        low[p]++;
        g[p] += ((dirmask >> p) & 1u) ? -1 : 1;
        dirmask ^= (1u << p) - 1u;
    }
    for (int d = 0; d < plan.D; d++)
        gray[plan.ref_digit_of[d]] = g[d];
}

} // namespace pqperm
