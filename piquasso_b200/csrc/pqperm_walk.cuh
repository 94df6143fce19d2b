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

// independent product chains of the generic walk for NC register-resident columns
#define PQ_CHAINS(NC) ((NC) >= 16 ? 4 : ((NC) >= 6 ? 2 : 1))

namespace pqperm {

#ifdef PQ_TRACE
// Experiment builds only (tools/trace_small.py): thread 0 of every CTA stamps
// %globaltimer at the phase boundaries of the generic walk.
static __device__ unsigned long long pq_trace_buf[8 * 8192];
__device__ __forceinline__ void trace_stamp(int phase)
{
    if (threadIdx.x == 0 && blockIdx.x < 8192) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        pq_trace_buf[blockIdx.x * 8 + phase] = t;
    }
}
#define PQ_STAMP(p) trace_stamp(p)
#else
#define PQ_STAMP(p)
#endif

// ---- dynamic segment distribution -------------------------------------------
// The warp schedulers do not share issue slots fairly between resident warps:
// with a static split the favoured warp of each SM sub-partition finishes at
// ~2/3 of the run and its partner runs alone (profiles/round1: 6.6 of 8 warps
// active on average, FP64 pipe 72%).  Instead every warp draws batches of 32
// consecutive segments from one global counter, so all warps stay busy until
// the segment range is exhausted.  Returns this lane's segment, or -1 when the
// range is exhausted (warp-uniform); the value may be >= P.seg_end in the last
// batch.
//
// A launch whose grid already holds one thread per segment (small problems: a
// single partly filled wave) needs no dispenser at all: `round` 0 hands thread t of
// the grid segment t, round 1 ends the loop -- two global atomic round trips less
// on a path that is a few microseconds long.
__device__ __forceinline__ long long next_batch(const WalkParams &P, int &round)
{
    const long long nseg = P.seg_end - P.seg_begin;
    if ((long long)gridDim.x * blockDim.x >= nseg) {
        if (round++ > 0)
            return -1;
        return P.seg_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    }
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
// ---- warp-cooperative seeding (generic walk) -----------------------------------
// The 32 lanes of a warp hold 32 CONSECUTIVE segments: their Gray digits agree on
// every high digit above the lowest few.  Instead of every lane summing all D rows
// (D*NC loads and 2*D*NC FMAs per lane -- as much work as 8-16 terms, and what
// small problems spend most of their time on), the warp evaluates the common
// part once:
//   * top-down over the high digits while all lanes carry the same weight: lane c
//     accumulates column c (and c + 32) of  a_0 + sum_d w_d * row_d;
//   * the remaining (varying) high digits are added per lane, row by row;
//   * the common vector goes through a per-warp shared buffer to every lane;
//   * the low digits (counter 0: g_d = 0 or r_d) contribute one of TWO vectors,
//     selected by the lane's parity, tabulated once per CTA (`lowtab`).
// Same terms, same partition; only the order of the additions inside a row sum
// changes.  All 32 lanes must call it (votes, shuffles): lanes without a segment
// pass a valid one and get their factor zeroed by the caller.
template <int NC, bool BINARY>
__device__ __forceinline__ void seed_segment_coop(const WalkParams &P, const double2 *smA,
                                                  const double2 *lowtab, double2 *cbuf,
                                                  const double *binom, long long seg,
                                                  double (&sr)[NC],
                                                  double (&si)[NC], unsigned &dirmask,
                                                  double &factor)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NCW = (NC + 31) / 32;
    const int lane = threadIdx.x & 31;
    double cx[NCW], cy[NCW];
#pragma unroll
    for (int c = 0; c < NCW; c++) {
        const int col = lane + 32 * c;
        const double2 a = col < NC ? smA[col] : make_double2(0.0, 0.0);
        cx[c] = a.x;
        cy[c] = a.y;
    }
#pragma unroll
    for (int j = 0; j < NC; j++)
        sr[j] = si[j] = 0.0;

    uint8_t chain[BINARY ? 1 : kMaxDigits];
    unsigned long long gh = 0;
    if (BINARY) {
        gh = (unsigned long long)seg ^ ((unsigned long long)seg >> 1);
    } else {
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
    }
    int odd = 0;
    double bin = 1.0;
    bool common = true; // warp-uniform: still inside the prefix of digits all lanes share
    for (int d = P.D - 1; d >= P.q; --d) {
        double w;
        if (BINARY) {
            const int g = (int)((gh >> (d - P.q)) & 1ull);
            odd ^= g;
            w = g ? -0.5 : 0.5; // rows are stored doubled
        } else {
            const int r = P.mult[d];
            const int g = odd ? r - chain[d] : chain[d];
            odd ^= (g & 1);
            bin *= binom[P.binom_off[d] + g];
            w = 0.5 * (double)(r - 2 * g);
        }
        if (common)
            common = __all_sync(FULL, w == __shfl_sync(FULL, w, 0));
        const double2 *row = smA + (d + 1) * NC;
        if (common) {
#pragma unroll
            for (int c = 0; c < NCW; c++) {
                const int col = lane + 32 * c;
                if (col < NC) {
                    const double2 a = row[col];
                    cx[c] = __fma_rn(w, a.x, cx[c]);
                    cy[c] = __fma_rn(w, a.y, cy[c]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = row[j];
                sr[j] = __fma_rn(w, a.x, sr[j]);
                si[j] = __fma_rn(w, a.y, si[j]);
            }
        }
    }
    // the common vector: through this warp's buffer to every lane
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NCW; c++) {
        const int col = lane + 32 * c;
        if (col < NC)
            cbuf[col] = make_double2(cx[c], cy[c]);
    }
    __syncwarp();
    const double2 *low = lowtab + (odd ? NC : 0);
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const double2 a = cbuf[j], l = low[j];
        sr[j] += a.x + l.x;
        si[j] += a.y + l.y;
    }
    // low digits: counter digits are 0, so g_d = 0 (even prefix) or r_d (odd)
    dirmask = 0;
    for (int d = P.q - 1; d >= 0; --d) {
        const int r = BINARY ? 1 : P.mult[d];
        dirmask |= (unsigned)odd << d;
        if (r & 1)
            odd = 0; // g_d = r_d odd flips the parity seen by the digits below
    }
    factor = odd ? -bin : bin;
}

// The two low-digit vectors of seed_segment_coop: lowtab[o * NC + j] =
// sum_{d < q} w_d(o) * row_d[j] for a segment entered with parity o.
template <int NC, bool BINARY, int NT>
__device__ __forceinline__ void build_lowtab(const WalkParams &P, const double2 *smA,
                                             double2 *lowtab)
{
    for (int t = threadIdx.x; t < 2 * NC; t += NT) {
        int odd = t / NC;
        const int j = t - odd * NC;
        double x = 0.0, y = 0.0;
        for (int d = P.q - 1; d >= 0; --d) {
            const int r = BINARY ? 1 : P.mult[d];
            const double w = odd ? -0.5 * (double)r : 0.5 * (double)r;
            const double2 a = smA[(d + 1) * NC + j];
            x = __fma_rn(w, a.x, x);
            y = __fma_rn(w, a.y, y);
            if (r & 1)
                odd = 0;
        }
        lowtab[t] = make_double2(x, y);
    }
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
        // pairwise combination of the partial chains (no extra multiplies:
        // (NC - CH) + (CH - 1) = NC - 1 complex products whatever CH is)
#pragma unroll
        for (int stride = 1; stride < CH; stride *= 2) {
#pragma unroll
            for (int c = 0; c + stride < CH; c += 2 * stride)
                cmul(cr[c], ci[c], cr[c + stride], ci[c + stride]);
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
//
// The matrix is staged in shared memory from the uploaded blob.  (Measured on B200
// and rejected: the matrix in the kernel's parameter block.  Walking it there costs
// register-indexed LDC.64 loads, twice as slow as uniform LDS.128 -- n = 24: 127 ->
// 213 us; staging it from there costs divergent constant loads -- config 3: 86 ->
// 108 us.  One 6 KB H2D copy is cheaper than either.)
template <int NC, bool BINARY, bool UNITCOLS, int NT>
__device__ __forceinline__ void generic_walk_body(const WalkParams &P, const double2 *src,
                                                  const double *binom)
{
    extern __shared__ double2 smA[]; // (D+1) x NC matrix, 2 x NC low vectors, NC per warp
    double2 *lowtab = smA + (P.D + 1) * NC;
    double2 *cbuf = lowtab + 2 * NC + (threadIdx.x >> 5) * NC;
    // step tables of the low counter (n-ary only), built by the CTA itself: the
    // digit moved on the step into m and (-1)^m prod_{d<q} C(r_d, c_d(m)).  In
    // shared memory they cost one LDS per step; nothing is uploaded for them.
    __shared__ double s_wtab[BINARY ? 1 : kMaxSegLenNary];
    __shared__ uint8_t s_sched[BINARY ? 1 : kMaxSegLenNary];
    const int W = (int)P.W;
    PQ_STAMP(0);
    {
        const int nelem = (P.D + 1) * NC;
        for (int i = threadIdx.x; i < nelem; i += NT)
            smA[i] = src[i];
        __syncthreads();
        PQ_STAMP(1);
        build_lowtab<NC, BINARY, NT>(P, smA, smA + (P.D + 1) * NC);
        if (!BINARY) {
            for (int m = threadIdx.x; m < W; m += NT) {
                int rest = m, p = -1;
                double w = (m & 1) ? -1.0 : 1.0;
                for (int d = 0; d < P.q; d++) {
                    const int L = P.radix[d];
                    const int c = rest % L;
                    rest /= L;
                    if (p < 0 && c != 0)
                        p = d;
                    w *= binom[P.binom_off[d] + c];
                }
                s_sched[m] = (uint8_t)(p < 0 ? 0 : p);
                s_wtab[m] = w;
            }
        }
    }
    __syncthreads();
    PQ_STAMP(2);
#ifdef PQ_TRACE
    bool trace_first = true;
#endif

    dd totre{0.0, 0.0}, totim{0.0, 0.0};
    int round = 0;
    for (;;) {
        // dynamic distribution: a warp takes the next 32 segments (see next_batch)
        const long long seg_drawn = next_batch(P, round);
        if (seg_drawn < 0)
            break;
        // the whole warp seeds together: lanes past the range redo the last segment
        // with weight 0
        const bool valid = seg_drawn < P.seg_end;
        if (__all_sync(0xffffffffu, !valid))
            continue;
        const long long seg = valid ? seg_drawn : P.seg_end - 1;
        double sr[NC], si[NC];
        unsigned dirmask;
        double factor;
        seed_segment_coop<NC, BINARY>(P, smA, lowtab, cbuf, binom, seg, sr, si, dirmask,
                                      factor);
        if (!valid)
            factor = 0.0;
#ifdef PQ_TRACE
        if (trace_first) {
            PQ_STAMP(3);
            trace_first = false;
        }
#endif

        dd segre{0.0, 0.0}, segim{0.0, 0.0};
        constexpr int CHUNK = 64;
        int p_next = (!BINARY && W > 1) ? s_sched[1] : 0; // fetched one step ahead
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
                        p = p_next;
                        p_next = s_sched[m + 1 < W ? m + 1 : m];
                        w = s_wtab[m];
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
        if (P.segsums && valid) {
            P.segsums[2 * (seg - P.seg_begin)] = segre.hi + segre.lo;
            P.segsums[2 * (seg - P.seg_begin) + 1] = segim.hi + segim.lo;
        }
        dd_add(totre, segre);
        dd_add(totim, segim);
    }
    PQ_STAMP(4);
    finish_grid<NT>(P, totre, totim);
    PQ_STAMP(5);
}

template <int NC, bool BINARY, bool UNITCOLS, int NT>
__global__ void __launch_bounds__(NT) perm_walk_generic(const __grid_constant__ WalkParams P)
{
    generic_walk_body<NC, BINARY, UNITCOLS, NT>(P, P.A2, P.binom);
}

#ifndef PQ_COL_CHUNK
#define PQ_COL_CHUNK 8
#endif

// Scheduling fence: the unrolled column loop is one huge basic block and ptxas
// tends to hoist the operand loads of all columns to its top.  Routing the
// running products and the row pointer through an empty volatile asm keeps a
// window of PQ_COL_CHUNK columns in flight.
template <int NP>
__device__ __forceinline__ void sched_fence(double (&pr)[NP], double (&pi)[NP],
                                            const double2 *&ptr)
{
#pragma unroll
    for (int t = 0; t < NP; t++)
        asm volatile("" : "+d"(pr[t]), "+d"(pi[t]));
    asm volatile("" : "+l"(ptr));
}

// ---- kernel 2: binary walk, one hypercube of the low digits per block ---------
// All multiplicities 1 (radix 2 everywhere).  The 2^B offsets of an aligned
// block share the Gray digits >= B and run through ALL 2^B values of the digits
// 0..B-1 -- a hypercube whose corner (all low digits 0) is a per-thread vector
// c_j and whose other vertices are c_j - sum_{d in e} 2 a_dj.  The term sign is
// (-1)^{sum of the Gray digits >= B} * (-1)^{|e|}.
//
// A block is evaluated column-major: for column j the 2^B - 1 other vertices
// are one add each (a binary tree of depth B from the corner; the B row
// operands come from the constant bank through the uniform datapath,
// LDCU.128 -> DADD R, R, -UR) and are multiplied straight into 2^B independent
// running products, one per term; the same pass applies the block-level Gray
// move (digit B + ctz(block+1), warp-uniform row) to the corner.  Per column
// and block that is B + 1 operand fetches for 2^B terms, the products give the
// FP64 pipe 2^B-way ILP with two warps per SM sub-partition resident, and
// there is no per-step integer bookkeeping at all.
// FP64 instructions per term: 2C (vertex adds + corner move) + 4(C-1)
// (products) + 2 (signed sum)  =  6C - 2, as in the reference's hot loop
// (src/permanent.cpp:218-250) but with every operand warp-uniform.
//
// `cA` is the (D+1) x NC doubled matrix in a CONSTANT bank: either the kernel's own
// parameter block (perm_walk_binary_pm: the matrix rides in the launch, nothing is
// uploaded beforehand) or the translation unit's __constant__ array (perm_walk_binary,
// for the widths whose matrix does not fit the 32 KB parameter space).
template <int NC, int B, int NT>
__device__ __forceinline__ void binary_walk_body(const WalkParams &P, const double2 *cA)
{
    constexpr int NP = 1 << B;

    dd totre{0.0, 0.0}, totim{0.0, 0.0};
    const int nblk = (int)(P.W >> B);
    int round = 0;
    for (;;) {
        const long long seg = next_batch(P, round);
        if (seg < 0)
            break;
        if (seg >= P.seg_end)
            continue;

        // ---- corner of the segment's first block (cf. seed_segment): digits
        // >= q from the segment index, digits B..q-1 at counter 0, digits < B at 0
        double cr[NC], ci[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double2 a = cA[j];
            cr[j] = a.x;
            ci[j] = a.y;
        }
        const unsigned long long gh = (unsigned long long)seg ^ ((unsigned long long)seg >> 1);
        int odd = __popcll(gh) & 1;
        unsigned dirmask = 0;
        for (int d = P.D - 1; d >= 0; --d) {
            double w;
            if (d >= P.q) {
                w = ((gh >> (d - P.q)) & 1ull) ? -0.5 : 0.5; // rows are stored doubled
            } else if (d >= B) {
                dirmask |= (unsigned)odd << d;
                w = odd ? -0.5 : 0.5;
                odd = 0;
            } else {
                w = 0.5;
            }
            const double2 *row = cA + (d + 1) * NC;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = row[j];
                cr[j] = __fma_rn(w, a.x, cr[j]);
                ci[j] = __fma_rn(w, a.y, ci[j]);
            }
        }
        const double factor = odd ? -1.0 : 1.0; // (-1)^{sum of the Gray digits >= B}

        // blocks are summed in plain FP64 in groups of 8 and the group sums folded
        // into the thread's double-double total
        for (int blk0 = 0; blk0 < nblk; blk0 += 8) {
            double accr = 0.0, acci = 0.0;
            const int blk1 = min(nblk, blk0 + 8);
            for (int blk = blk0; blk < blk1; ++blk) {
                // the Gray move that leads from this block to the next one
                const bool more = blk + 1 < nblk;
                const int p = B + __ffs(blk + 1) - 1;
                const double sgh = more ? (((dirmask >> p) & 1u) ? 1.0 : -1.0) : 0.0;
                dirmask ^= (1u << p) - 1u;
                const double2 *rowh = cA + (more ? (p + 1) * NC : 0);

                double pr[NP], pi[NP];
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double c0r = cr[j], c0i = ci[j];
                    if (j == 0) {
                        pr[0] = c0r;
                        pi[0] = c0i;
                    } else {
                        cmul(pr[0], pi[0], c0r, c0i);
                    }
                    // vertex e = vertex (e with its lowest set bit cleared) minus the
                    // doubled row of that bit: one add per vertex, every operand an
                    // exact matrix entry (a table of pre-rounded subset sums would put
                    // the SAME rounding error into one vertex of every block, which
                    // the cancellation of the Glynn sum amplifies)
                    double vr[NP], vi[NP];
                    vr[0] = c0r;
                    vi[0] = c0i;
#pragma unroll
                    for (int e = 1; e < NP; e++) {
                        const int d = (e & 1) ? 0 : ((e & 2) ? 1 : ((e & 4) ? 2 : 3));
                        const double2 a = cA[(d + 1) * NC + j];
                        vr[e] = vr[e & (e - 1)] - a.x;
                        vi[e] = vi[e & (e - 1)] - a.y;
                        if (j == 0) {
                            pr[e] = vr[e];
                            pi[e] = vi[e];
                        } else {
                            cmul(pr[e], pi[e], vr[e], vi[e]);
                        }
                    }
                    const double2 ah = rowh[j];
                    cr[j] = __fma_rn(sgh, ah.x, c0r);
                    ci[j] = __fma_rn(sgh, ah.y, c0i);
                    if (PQ_COL_CHUNK > 0 && (j % (PQ_COL_CHUNK > 0 ? PQ_COL_CHUNK : 1)) ==
                            PQ_COL_CHUNK - 1 && j + 1 < NC)
                        sched_fence<NP>(pr, pi, rowh);
                }
                // signed sum over the vertices: (-1)^{|e|}
                double br = 0.0, bi = 0.0;
#pragma unroll
                for (int e = 0; e < NP; e++) {
                    if (__builtin_popcount(e) & 1) {
                        br -= pr[e];
                        bi -= pi[e];
                    } else {
                        br += pr[e];
                        bi += pi[e];
                    }
                }
                // every block move flips one Gray digit >= B
                if (blk & 1) {
                    accr -= br;
                    acci -= bi;
                } else {
                    accr += br;
                    acci += bi;
                }
            }
            dd_add(totre, factor * accr);
            dd_add(totim, factor * acci);
            if (P.segsums) {
                P.segsums[2 * (seg - P.seg_begin)] += factor * accr;
                P.segsums[2 * (seg - P.seg_begin) + 1] += factor * acci;
            }
        }
    }
    finish_grid<NT>(P, totre, totim);
}

// Binary problems with unit columns have D + 1 = NC rows: an NC x NC matrix.
template <int NC>
struct WalkParamsM {
    WalkParams P;
    double2 m[NC * NC];
};
// the widest such matrix must fit the kernel parameter space (32764 bytes, CUDA >= 12.1)
static_assert(sizeof(WalkParamsM<kBinMaxParamCols>) <= 32764, "parameter block too large");

template <int NC, int B, int NT>
__global__ void __launch_bounds__(NT)
perm_walk_binary_pm(const __grid_constant__ WalkParamsM<NC> Q)
{
    binary_walk_body<NC, B, NT>(Q.P, Q.m);
}

#ifdef PQ_BINARY_CONST_MATRIX
template <int NC, int B, int NT>
__global__ void __launch_bounds__(NT) perm_walk_binary(const __grid_constant__ WalkParams P)
{
    binary_walk_body<NC, B, NT>(P, PQ_BINARY_CONST_MATRIX);
}
#endif // PQ_BINARY_CONST_MATRIX

} // namespace pqperm
