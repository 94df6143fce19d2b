// pqperm_walk.cuh -- the single-permanent Gray-code walks (sm_100a).
//
// What is computed (canonical statement):
//   perm * 2^(N-1) = sum_{offset < idx_max} term(offset)
//   term(offset)   = (-1)^{sum g} * prod_d C(r_d, g_d) * prod_j s_j^{c_j}
//   s_j            = a_0j + sum_d a_{d+1,j} (r_d - 2 g_d),  g = gray(offset)
// with gray() the mixed-radix reflected code of
// src/n_aryGrayCodeCounter.hpp:170-194 (reference tree).  The hot loop being
// replaced is src/permanent.cpp:218-250.
//
// Mapping onto the GPU: the digits are cut at q.  Digits q..D-1 of the
// offset ARE the segment (= thread) index; digits 0..q-1 are walked by every
// thread in lock step, W = prod_{d<q} radix[d] terms.  Because all segments
// start at a multiple of W, the digit that moves on step m is the same for
// every thread of the grid (uniform shared-memory / constant-bank operand);
// only the direction of the move differs per thread and is kept as one bit
// per digit (`dirmask`, bit d = parity of the Gray digits above d).
#pragma once

#include "pqperm_device.cuh"

#ifndef PQ_CHAINS
#define PQ_CHAINS(NC) ((NC) >= 16 ? 4 : ((NC) >= 6 ? 2 : 1))
#endif

namespace pqperm {

// ---- dynamic segment distribution -------------------------------------------
// The warp schedulers do not share issue slots fairly between resident warps:
// with a static split the favoured warp of each SM sub-partition finishes at
// ~2/3 of the run and its partner runs alone (profiles/round1: 6.6 of 8 warps
// active on average, FP64 pipe 72%).  Instead every warp draws batches of 32
// consecutive segments from one global counter, so all warps stay busy until
// the segment range is exhausted.  Returns this lane's segment, or -1 when the
// range is exhausted (warp-uniform); the value may be >= P.seg_end in the last
// batch.
__device__ __forceinline__ long long next_batch(const WalkParams &P)
{
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0)
        base = atomicAdd(P.counter, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base >= P.seg_end - P.seg_begin)
        return -1;
    return P.seg_begin + (long long)base + (threadIdx.x & 31);
}

// ---- seeding: s_j, direction mask and weight of term(seg * W) ---------------
// Restates the seed of src/permanent.cpp:174-202 for offset = seg * W.
template <int NC, bool BINARY>
__device__ __forceinline__ void seed_segment(const WalkParams &P, const double2 *smA,
                                             long long seg, double (&sr)[NC],
                                             double (&si)[NC], unsigned &dirmask,
                                             double &factor)
{
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const double2 a = smA[j];
        sr[j] = a.x;
        si[j] = a.y;
    }
    int odd = 0;
    double bin = 1.0;
    if (BINARY) {
        // digits q..D-1 are the bits of seg; gray = seg ^ (seg >> 1)
        const unsigned long long gh =
            (unsigned long long)seg ^ ((unsigned long long)seg >> 1);
        for (int d = P.D - 1; d >= P.q; --d) {
            const int g = (int)((gh >> (d - P.q)) & 1ull);
            const double w = g ? -0.5 : 0.5; // rows are stored doubled
            const double2 *row = smA + (d + 1) * NC;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = row[j];
                sr[j] = __fma_rn(w, a.x, sr[j]);
                si[j] = __fma_rn(w, a.y, si[j]);
            }
        }
        odd = __popcll(gh) & 1;
    } else {
        uint8_t chain[kMaxDigits];
        unsigned long long rest = (unsigned long long)seg;
        for (int d = P.q; d < P.D; ++d) {
            const unsigned L = P.radix[d];
            if (rest >> 32) {
                chain[d] = (uint8_t)(rest % L);
                rest /= L;
            } else {
                const unsigned r32 = (unsigned)rest;
                chain[d] = (uint8_t)(r32 % L);
                rest = r32 / L;
            }
        }
        for (int d = P.D - 1; d >= P.q; --d) {
            const int r = P.mult[d];
            const int g = odd ? r - chain[d] : chain[d];
            odd ^= (g & 1);
            bin *= P.binom[P.binom_off[d] + g];
            const double w = 0.5 * (double)(r - 2 * g);
            const double2 *row = smA + (d + 1) * NC;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = row[j];
                sr[j] = __fma_rn(w, a.x, sr[j]);
                si[j] = __fma_rn(w, a.y, si[j]);
            }
        }
    }
    // low digits: counter digits are 0, so g_d = 0 (even prefix) or r_d (odd)
    dirmask = 0;
    for (int d = P.q - 1; d >= 0; --d) {
        const int r = BINARY ? 1 : P.mult[d];
        dirmask |= (unsigned)odd << d;
        const double w = odd ? -0.5 * (double)r : 0.5 * (double)r;
        const double2 *row = smA + (d + 1) * NC;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double2 a = row[j];
            sr[j] = __fma_rn(w, a.x, sr[j]);
            si[j] = __fma_rn(w, a.y, si[j]);
        }
        if (r & 1)
            odd = 0; // g_d = r_d odd flips the parity seen by the digits below
    }
    factor = odd ? -bin : bin;
}

// ---- prod_j s_j^{c_j} ------------------------------------------------------
template <int NC, bool UNITCOLS, int CHAINS = PQ_CHAINS(NC)>
__device__ __forceinline__ void column_product(const WalkParams &P, const double (&sr)[NC],
                                               const double (&si)[NC], double &pr,
                                               double &pi)
{
    if (UNITCOLS) {
        // independent partial chains so that the FP64 pipe always has work
        constexpr int CH = CHAINS;
        double cr[CH], ci[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int j0 = (NC * c) / CH;
            const int j1 = (NC * (c + 1)) / CH;
            cr[c] = sr[j0];
            ci[c] = si[j0];
#pragma unroll
            for (int j = j0 + 1; j < j1; j++)
                cmul(cr[c], ci[c], sr[j], si[j]);
        }
        if (CH == 4) {
            cmul(cr[0], ci[0], cr[1], ci[1]);
            cmul(cr[2], ci[2], cr[3], ci[3]);
            cmul(cr[0], ci[0], cr[2], ci[2]);
        } else if (CH == 2) {
            cmul(cr[0], ci[0], cr[1], ci[1]);
        }
        pr = cr[0];
        pi = ci[0];
    } else {
        pr = 1.0;
        pi = 0.0;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int c = P.colmult[j];
            for (int k = 0; k < c; k++)
                cmul(pr, pi, sr[j], si[j]);
        }
    }
}

// multiply a double-double by a double (two-product through FMA)
__device__ __forceinline__ dd dd_scale(const dd &a, double f)
{
    const double ph = a.hi * f;
    const double pe = __fma_rn(a.hi, f, -ph);
    dd r;
    r.hi = ph;
    r.lo = __fma_rn(a.lo, f, pe);
    return r;
}

// ---- kernel 1: generic n-ary walk ------------------------------------------
// One Gray step = NC uniform LDS.128 + 2*NC DFMA (row-sum update) + product +
// 2 DFMA (weighted accumulate).  Per-thread integer work: one bit test, one
// mask update.  The binomial weight of the low digits is the same for every
// thread (C(r,g) = C(r,r-g)), so it comes from the host-built wtab[m]; the
// thread's own high-digit weight is applied once per segment.
template <int NC, bool BINARY, bool UNITCOLS, int NT>
__global__ void __launch_bounds__(NT) perm_walk_generic(const __grid_constant__ WalkParams P)
{
    extern __shared__ double2 smA[];
    {
        const int nelem = (P.D + 1) * NC;
        for (int i = threadIdx.x; i < nelem; i += NT)
            smA[i] = P.A2[i];
    }
    __syncthreads();

    dd totre{0.0, 0.0}, totim{0.0, 0.0};
    const int W = (int)P.W;
    for (;;) {
        // dynamic distribution: a warp takes the next 32 segments (see next_batch)
        const long long seg = next_batch(P);
        if (seg < 0)
            break;
        if (seg >= P.seg_end)
            continue;
        double sr[NC], si[NC];
        unsigned dirmask;
        double factor;
        seed_segment<NC, BINARY>(P, smA, seg, sr, si, dirmask, factor);

        dd segre{0.0, 0.0}, segim{0.0, 0.0};
        constexpr int CHUNK = 64;
        for (int m0 = 0; m0 < W; m0 += CHUNK) {
            const int m1 = min(W, m0 + CHUNK);
            double accr = 0.0, acci = 0.0;
            for (int m = m0; m < m1; ++m) {
                double w = 1.0;
                if (m != 0) {
                    int p;
                    if (BINARY) {
                        p = __ffs(m) - 1;
                        w = (m & 1) ? -1.0 : 1.0;
                    } else {
                        p = P.sched[m];
                        w = P.wtab[m];
                    }
                    const double sg = ((dirmask >> p) & 1u) ? 1.0 : -1.0;
                    dirmask ^= (1u << p) - 1u;
                    const double2 *row = smA + (p + 1) * NC;
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const double2 a = row[j];
                        sr[j] = __fma_rn(sg, a.x, sr[j]);
                        si[j] = __fma_rn(sg, a.y, si[j]);
                    }
                }
                double pr, pi;
                column_product<NC, UNITCOLS>(P, sr, si, pr, pi);
                accr = __fma_rn(w, pr, accr);
                acci = __fma_rn(w, pi, acci);
            }
            dd_add(segre, accr);
            dd_add(segim, acci);
        }
        segre = dd_scale(segre, factor);
        segim = dd_scale(segim, factor);
        if (P.segsums) {
            P.segsums[2 * (seg - P.seg_begin)] = segre.hi + segre.lo;
            P.segsums[2 * (seg - P.seg_begin) + 1] = segim.hi + segim.lo;
        }
        dd_add(totre, segre);
        dd_add(totim, segim);
    }
    block_reduce_store<NT>(totre, totim, P.partials + 4 * (size_t)blockIdx.x);
}

#ifdef PQ_BINARY_CONST_MATRIX
// ---- kernel 2: binary constant-bank walk -----------------------------------
// All multiplicities 1 (radix 2 everywhere): the step sequence inside an
// aligned block of 2^B offsets is the ruler sequence, known at compile time.
// The rows of digits 0..B-1 live in the kernel parameter block (constant
// bank), so the 2^B - 1 inner steps are pure FP64: per column 2 DFMA with an
// immediate-offset constant operand + 2 DMUL + 2 DFMA, no loads and no integer
// bookkeeping.  Only the last step of each block moves a run-time digit
// (B + ctz(block index + 1)) whose row comes from shared memory.
template <int NC, int B, int CHAINS, int NT>
__global__ void __launch_bounds__(NT) perm_walk_binary(const __grid_constant__ WalkParams P)
{
    const double2 *smA = PQ_BINARY_CONST_MATRIX;

    dd totre{0.0, 0.0}, totim{0.0, 0.0};
    const int nblk = (int)(P.W >> B);
    for (;;) {
        const long long seg = next_batch(P);
        if (seg < 0)
            break;
        if (seg >= P.seg_end)
            continue;
        double sr[NC], si[NC];
        unsigned dirmask;
        double factor;
        seed_segment<NC, true>(P, smA, seg, sr, si, dirmask, factor);

        // direction of the next move of digits 0..B-1 as +-1.0
        double sg[B];
#pragma unroll
        for (int d = 0; d < B; d++)
            sg[d] = ((dirmask >> d) & 1u) ? 1.0 : -1.0;

        dd segre{0.0, 0.0}, segim{0.0, 0.0};
        double accr, acci;
        column_product<NC, true, CHAINS>(P, sr, si, accr, acci); // term m = 0

        for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll
            for (int i = 1; i < (1 << B); ++i) {
                // digit ctz(i): compile-time after unrolling
                const int p = (i & 1) ? 0 : ((i & 2) ? 1 : ((i & 4) ? 2 : 3));
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double2 a = smA[(p + 1) * NC + j];
                    sr[j] = __fma_rn(sg[p], a.x, sr[j]);
                    si[j] = __fma_rn(sg[p], a.y, si[j]);
                }
#pragma unroll
                for (int d = 0; d < B; d++)
                    if (d < p)
                        sg[d] = -sg[d];
                double pr, pi;
                column_product<NC, true, CHAINS>(P, sr, si, pr, pi);
                if (i & 1) {
                    accr -= pr;
                    acci -= pi;
                } else {
                    accr += pr;
                    acci += pi;
                }
            }
            if (blk + 1 < nblk) {
                const int p = B + __ffs(blk + 1) - 1;
                const double sgh = ((dirmask >> p) & 1u) ? 1.0 : -1.0;
                dirmask ^= (1u << p) - 1u;
#pragma unroll
                for (int d = 0; d < B; d++)
                    sg[d] = -sg[d];
                const double2 *row = smA + (p + 1) * NC;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double2 a = row[j];
                    sr[j] = __fma_rn(sgh, a.x, sr[j]);
                    si[j] = __fma_rn(sgh, a.y, si[j]);
                }
                double pr, pi;
                column_product<NC, true, CHAINS>(P, sr, si, pr, pi);
                dd_add(segre, accr);
                dd_add(segim, acci);
                accr = pr;
                acci = pi;
            }
        }
        dd_add(segre, accr);
        dd_add(segim, acci);
        segre = dd_scale(segre, factor);
        segim = dd_scale(segim, factor);
        if (P.segsums) {
            P.segsums[2 * (seg - P.seg_begin)] = segre.hi + segre.lo;
            P.segsums[2 * (seg - P.seg_begin) + 1] = segim.hi + segim.lo;
        }
        dd_add(totre, segre);
        dd_add(totim, segim);
    }
    block_reduce_store<NT>(totre, totim, P.partials + 4 * (size_t)blockIdx.x);
}

#endif // PQ_BINARY_CONST_MATRIX

} // namespace pqperm
