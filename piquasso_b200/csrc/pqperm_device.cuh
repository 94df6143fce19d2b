// pqperm_device.cuh -- parameter blocks and device helpers shared by the
// permanent kernels (sm_100a).  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pqperm_limits.h"

namespace pqperm {

// Parameter block of the single-permanent walks.  Passed by value as a
// __grid_constant__ so that the small per-digit tables sit in the constant
// bank (warp-uniform reads through the uniform datapath).
struct WalkParams {
    const double2 *A2;     // (D+1) x NC: row 0 = pinned row a_0, rows 1..D = 2*a_d
    const double *binom;   // flattened C(r_d, g) tables, binom_off[d] + g
    double *partials;      // [gridDim.x][4] re_hi, re_lo, im_hi, im_lo
    double *out4;          // the launch's sum (double-double), written by the last CTA to finish
    double *segsums;       // optional [seg_end - seg_begin][2] per-segment sums
    unsigned long long *counter;  // next undistributed segment (relative); 0 between launches
    unsigned int *done;    // CTAs that have stored their partial; 0 between launches
    long long seg_begin;   // segments [seg_begin, seg_end) belong to this launch
    long long seg_end;
    long long W;           // terms per segment = prod_{d<q} radix[d]
    int D;                 // Gray digits
    int q;                 // digits 0..q-1 are walked inside the segment
    uint8_t radix[kMaxDigits];    // r_d + 1
    uint8_t mult[kMaxDigits];     // r_d
    uint8_t colmult[kMaxCols];    // c_j (1 for padding columns)
    uint16_t binom_off[kMaxDigits];
};

// ---- batched Laplace walk (pqperm_laplace.cuh) ------------------------------

struct LapProblem {
    long long a_off;     // first double2 of this problem's (D+1) x NCP matrix in the pack
    long long nseg;      // Gray segments walked by this launch ...
    long long seg_begin; // ... starting at this one (term-space split over ranks)
    int first_block;     // CTAs [first_block, first_block + nblocks) work on this problem
    int nblocks;
    int D;               // Gray digits
    int q;               // low digits walked inside a segment
    int W;               // terms per segment
    int exp2;            // results are scaled by 2^-exp2 (= sum_rows - 1)
    int nc;              // active columns (the rest of NCP is padding)
    int tag;             // caller's index (sampler: row of the pmf output)
    uint8_t mult[kMaxDigits];    // r_d
    uint8_t colmult[kMaxCols];   // c_j (1 for padding columns)
    uint16_t colmode[kMaxCols];  // gather mode: column of U feeding compact column j
    uint16_t rowmode[kMaxDigits + 1];  // gather mode: row of U of the pinned row / of digit d-1
};

// Column tables of a WIDE problem (more than kMaxCols active columns, S = 32):
// kept out of LapProblem so that the sampler's descriptors stay small.
struct LapWide {
    uint8_t colmult[kLapMaxCols];
    uint16_t colmode[kLapMaxCols];
};

struct LapParams {
    const LapProblem *prob;
    const LapWide *wide; // [nprob] when the launch is a wide one, else nullptr
    const double2 *U;    // gather mode (sampler): ldu x ldu matrix every problem is a minor of
    int ldu;
    const double2 *A2;   // packed mode: per-problem matrices
    int perm_only;       // 1: only the full product is wanted (batched permanents)
    int max_D;           // the CTAs' shared matrix area holds (max_D + 1) x NCP entries
    double *partials;    // [total CTAs][NCP + 1][4]: re hi, re lo, im hi, im lo
    double2 *out;        // [nprob][NCP + 1]: per compact column, then the full product
    int nprob;
};

// ---- double-double accumulation -------------------------------------------

struct dd {
    double hi, lo;
};

// a += b with the rounding error of the high part collected in lo (TwoSum).
__device__ __forceinline__ void dd_add(dd &a, double b)
{
    const double s = a.hi + b;
    const double bb = s - a.hi;
    const double e = (a.hi - (s - bb)) + (b - bb);
    a.hi = s;
    a.lo += e;
}

__device__ __forceinline__ void dd_add(dd &a, const dd &b)
{
    dd_add(a, b.hi);
    a.lo += b.lo;
}

__device__ __forceinline__ dd dd_shfl_down(const dd &a, int delta)
{
    dd r;
    r.hi = __shfl_down_sync(0xffffffffu, a.hi, delta);
    r.lo = __shfl_down_sync(0xffffffffu, a.lo, delta);
    return r;
}

// Sum (re, im) double-double pairs over the block; thread 0 writes 4 doubles.
// Fixed tree => bit-reproducible for a fixed launch geometry.
template <int NT>
__device__ __forceinline__ void block_reduce_store(dd re, dd im, double *out4)
{
    __shared__ double red[(NT / 32) * 4];
#pragma unroll
    for (int delta = 16; delta > 0; delta >>= 1) {
        dd_add(re, dd_shfl_down(re, delta));
        dd_add(im, dd_shfl_down(im, delta));
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        red[warp * 4 + 0] = re.hi;
        red[warp * 4 + 1] = re.lo;
        red[warp * 4 + 2] = im.hi;
        red[warp * 4 + 3] = im.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dd r{red[0], red[1]}, i{red[2], red[3]};
        for (int w = 1; w < NT / 32; w++) {
            dd_add(r, dd{red[w * 4 + 0], red[w * 4 + 1]});
            dd_add(i, dd{red[w * 4 + 2], red[w * 4 + 3]});
        }
        out4[0] = r.hi;
        out4[1] = r.lo;
        out4[2] = i.hi;
        out4[3] = i.lo;
    }
}

// End of a walk kernel: the CTA's partial goes to P.partials; the CTA that
// finishes LAST sums all partials in index order (fixed tree, so the result is
// reproducible for a fixed grid), writes the four doubles to P.out4 and re-arms the
// two counters for the next launch -- no separate reduction kernel, no memset.
template <int NT>
__device__ __forceinline__ void finish_grid(const WalkParams &P, dd re, dd im)
{
    __shared__ bool s_last;
    if (gridDim.x == 1) { // nobody to wait for (the static path leaves the counters at zero)
        block_reduce_store<NT>(re, im, P.out4);
        return;
    }
    block_reduce_store<NT>(re, im, P.partials + 4 * (size_t)blockIdx.x);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(P.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    dd r{0.0, 0.0}, i{0.0, 0.0};
    for (unsigned b = threadIdx.x; b < gridDim.x; b += NT) {
        const double *q = P.partials + 4 * (size_t)b;
        dd_add(r, dd{__ldcg(q + 0), __ldcg(q + 1)});
        dd_add(i, dd{__ldcg(q + 2), __ldcg(q + 3)});
    }
    __syncthreads(); // block_reduce_store's scratch is reused
    block_reduce_store<NT>(r, i, P.out4);
    if (threadIdx.x == 0) {
        *P.counter = 0ull;
        *P.done = 0u;
    }
}

// complex multiply-in-place: (pr, pi) *= (sr, si); 2 DMUL + 2 DFMA
__device__ __forceinline__ void cmul(double &pr, double &pi, double sr, double si)
{
    const double nr = __fma_rn(pr, sr, -(pi * si));
    const double ni = __fma_rn(pr, si, pi * sr);
    pr = nr;
    pi = ni;
}

// complex multiply-accumulate: (ar, ai) += (pr, pi) * (sr, si); 4 DFMA
__device__ __forceinline__ void cfma(double &ar, double &ai, double pr, double pi, double sr,
                                     double si)
{
    ar = __fma_rn(pr, sr, __fma_rn(-pi, si, ar));
    ai = __fma_rn(pr, si, __fma_rn(pi, sr, ai));
}

} // namespace pqperm
