// pqperm_laplace.cuh -- batched Laplace-expansion walk (sm_100a).
//
// Replaces the hot loop of permanent_laplace_cpp, src/permanent_laplace.cpp:
// 191-223 of the reference: the same Gray-code walk as the permanent, but every
// term contributes to C sums, sum l getting the product with ONE factor of
// column l removed (Lemma 1 of arXiv:2005.04214):
//   out_l * 2^(N-1) = sum_offset (-1)^{sum g} prod_d C(r_d,g_d)
//                       * s_l^{c_l-1} prod_{k != l} s_k^{c_k}
// The reference recomputes each of the C products from scratch (O(C*M) per
// term).  Here all leave-one-out products of a term come from ONE balanced
// product tree over the lane's columns (unit column multiplicities, what the
// sampler issues): an up-sweep stores the product of every subtree, a
// down-sweep hands every node the product of everything outside it, and a
// leaf's outside product is accumulated by a complex FMA
//   acc_l += outside(parent) * inside(sibling)
// -- per column two complex multiplies and one complex multiply-add, 14 FP64
// instructions with the row-sum update, exactly what a suffix/prefix pass
// costs, but with a dependency depth of 2 log2(C) multiplies instead of 2 C:
// the FP64 pipe sees C/2 independent chains instead of two.  (Column
// multiplicities > 1 keep the chunked suffix/prefix passes.)
//
// Batch layout: one launch walks many independent problems (the sampler's
// (shot, photon) problems).  A CTA works on one problem; S adjacent lanes
// ("group") share one Gray segment and split the columns (lane h owns columns
// h, h+S, ...), so that row sums, subtree products and accumulators of up to 64
// columns stay in registers.
//
// Accumulation: a thread sums the terms of one segment (<= 256) in plain FP64
// registers and then folds them into its double-double totals, which live in
// shared memory ([slot][thread], conflict-free); warp, CTA and cross-CTA
// reductions stay double-double (TwoSum) up to the final scaling.
#pragma once

#include "pqperm_device.cuh"

namespace pqperm {

__device__ __forceinline__ double small_binom(int n, int k)
{
    if (k > n - k)
        k = n - k;
    double r = 1.0;
    for (int i = 1; i <= k; i++)
        r = r * (double)(n - k + i) / (double)i;
    return r;
}

// ---- balanced product tree over the N columns of a lane ------------------------
// Node [LO, HI) with HI - LO >= 2 splits at MID = (LO + HI) / 2; every boundary
// between two adjacent leaves is the split point of exactly one node, so the
// product of node [LO, HI) is stored at index MID of (nr, ni).

// product of the subtree [LO, HI): a leaf's row sum or a stored node product
template <int LO, int HI, int N>
__device__ __forceinline__ void tree_value(const double (&sr)[N], const double (&si)[N],
                                           const double (&nr)[N], const double (&ni)[N],
                                           double &vr, double &vi)
{
    if constexpr (HI - LO == 1) {
        vr = sr[LO];
        vi = si[LO];
    } else {
        vr = nr[(LO + HI) / 2];
        vi = ni[(LO + HI) / 2];
    }
}

// up-sweep: products of all internal nodes below [LO, HI)
template <int LO, int HI, int N>
__device__ __forceinline__ void tree_up(const double (&sr)[N], const double (&si)[N],
                                        double (&nr)[N], double (&ni)[N])
{
    if constexpr (HI - LO >= 2) {
        constexpr int MID = (LO + HI) / 2;
        tree_up<LO, MID, N>(sr, si, nr, ni);
        tree_up<MID, HI, N>(sr, si, nr, ni);
        double lr, li, rr, ri;
        tree_value<LO, MID, N>(sr, si, nr, ni, lr, li);
        tree_value<MID, HI, N>(sr, si, nr, ni, rr, ri);
        nr[MID] = __fma_rn(lr, rr, -(li * ri));
        ni[MID] = __fma_rn(lr, ri, li * rr);
    }
}

// down-sweep from node [LO, HI) whose outside product (term weight included) is
// (our, oui) -- real when REALW.  A leaf child receives outside * inside(sibling)
// straight into its accumulator; its row sum is then moved to the NEXT term
// (s += sg * row), which is the last thing this term needs it for.
template <int LO, int HI, int N, int S, bool REALW>
__device__ __forceinline__ void tree_down(double (&sr)[N], double (&si)[N],
                                          const double (&nr)[N], const double (&ni)[N],
                                          double our, double oui, double (&accr)[N],
                                          double (&acci)[N], const double2 *row, double sg)
{
    static_assert(HI - LO >= 2, "tree_down needs an internal node");
    constexpr int MID = (LO + HI) / 2;
    constexpr bool LLEAF = (MID - LO == 1), RLEAF = (HI - MID == 1);
    double lr, li, rr, ri;
    tree_value<LO, MID, N>(sr, si, nr, ni, lr, li);
    tree_value<MID, HI, N>(sr, si, nr, ni, rr, ri);
    double xlr = 0.0, xli = 0.0, xrr = 0.0, xri = 0.0;
    if constexpr (LLEAF) {
        if (REALW) {
            accr[LO] = __fma_rn(our, rr, accr[LO]);
            acci[LO] = __fma_rn(our, ri, acci[LO]);
        } else {
            cfma(accr[LO], acci[LO], our, oui, rr, ri);
        }
    } else {
        if (REALW) {
            xlr = our * rr;
            xli = our * ri;
        } else {
            xlr = our;
            xli = oui;
            cmul(xlr, xli, rr, ri);
        }
    }
    if constexpr (RLEAF) {
        if (REALW) {
            accr[MID] = __fma_rn(our, lr, accr[MID]);
            acci[MID] = __fma_rn(our, li, acci[MID]);
        } else {
            cfma(accr[MID], acci[MID], our, oui, lr, li);
        }
    } else {
        if (REALW) {
            xrr = our * lr;
            xri = our * li;
        } else {
            xrr = our;
            xri = oui;
            cmul(xrr, xri, lr, li);
        }
    }
    if constexpr (LLEAF) {
        const double2 a = row[LO * S];
        sr[LO] = __fma_rn(sg, a.x, sr[LO]);
        si[LO] = __fma_rn(sg, a.y, si[LO]);
    }
    if constexpr (RLEAF) {
        const double2 a = row[MID * S];
        sr[MID] = __fma_rn(sg, a.x, sr[MID]);
        si[MID] = __fma_rn(sg, a.y, si[MID]);
    }
    if constexpr (!LLEAF)
        tree_down<LO, MID, N, S, false>(sr, si, nr, ni, xlr, xli, accr, acci, row, sg);
    if constexpr (!RLEAF)
        tree_down<MID, HI, N, S, false>(sr, si, nr, ni, xrr, xri, accr, acci, row, sg);
}

// Product of the OTHER lanes' totals in a group of S lanes (butterfly: T = product
// of my sub-group, O = the same without me; log2 S shuffle rounds).
template <int S>
__device__ __forceinline__ void others_product(double lr, double li, double &olr, double &oli)
{
    double tr = lr, ti = li;
    olr = 1.0;
    oli = 0.0;
#pragma unroll
    for (int x = 1; x < S; x <<= 1) {
        const double pr = __shfl_xor_sync(0xffffffffu, tr, x);
        const double pi = __shfl_xor_sync(0xffffffffu, ti, x);
        if (x == 1) {
            olr = pr;
            oli = pi;
        } else {
            cmul(olr, oli, pr, pi);
        }
        if (2 * x < S)
            cmul(tr, ti, pr, pi);
    }
}

// Slots of a thread's double-double totals in shared memory: column j uses
// slots 4j .. 4j+3 (re hi, re lo, im hi, im lo), the full product 4 NCL .. 4 NCL+3
// (batched permanents, mode kLapPerm: only the product, in slots 0 .. 3).
__device__ __forceinline__ void dd_fold(double *tot, int slot, int nt, double v)
{
    // tot[slot] (hi), tot[slot + 1] (lo) += v   (TwoSum)
    double *hi = tot + (size_t)slot * nt, *lo = hi + nt;
    const double a = *hi;
    const double s = a + v;
    const double bb = s - a;
    const double e = (a - (s - bb)) + (v - bb);
    *hi = s;
    *lo += e;
}

// MODE: what a term contributes to.  kLapLoo: the C leave-one-out sums only (the
// sampler: every column of the compact problem has multiplicity >= 1);
// kLapLooFull: also the full product, which the columns of the CALLER's matrix
// with multiplicity 0 receive (reference quirk, src/permanent_laplace.cpp:180,210);
// kLapPerm: the full product only (batched permanents).
constexpr int kLapLoo = 0, kLapLooFull = 1, kLapPerm = 2;

// terms of the walk loop unrolled together (tuning knob; 1 = no unrolling)
#ifndef PQ_LAP_UNROLL
#define PQ_LAP_UNROLL 1
#endif
constexpr int kLapUnroll = PQ_LAP_UNROLL;

// One step of the low counter as the walk consumes it: byte offset of the row that
// moves INTO local term m (relative to row 0 of the CTA's matrix) and the bit of
// that digit in the threads' direction masks.
struct LapStep {
    unsigned rowoff;
    unsigned bit;
};

template <int NCL, int S, bool UNITCOLS, int MODE>
__global__ void __launch_bounds__(kLapThreads) laplace_walk_kernel(const LapParams P)
{
    constexpr int NCP = NCL * S;
    constexpr int NT = kLapThreads;
    // slots of a thread's double-double totals: batched permanents keep only the
    // full product (slots 0..3), the other modes the columns and then the product
    constexpr int kSlots = MODE == kLapPerm ? 4 : 4 * NCL + 4;
    constexpr int kFull = MODE == kLapPerm ? 0 : 4 * NCL;
    // chunked suffix / prefix chains of the general (column multiplicity) flavour
    constexpr int CH = NCL >= 4 ? 2 : 1;
    constexpr int CLEN = (NCL + CH - 1) / CH;              // longest chunk
    extern __shared__ double2 smA[];              // (D+1) x NCP, then the totals
    __shared__ double s_wtab[kLapMaxSegLen];
    __shared__ LapStep s_step[kLapMaxSegLen + 1];
    __shared__ int s_prob;

    // ---- which problem does this CTA belong to (binary search over first_block)
    if (threadIdx.x == 0) {
        int lo = 0, hi = P.nprob - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.prob[mid].first_block <= (int)blockIdx.x)
                lo = mid;
            else
                hi = mid - 1;
        }
        s_prob = lo;
    }
    __syncthreads();
    const LapProblem &Q = P.prob[s_prob];
    const uint8_t *colmult = P.wide ? P.wide[s_prob].colmult : Q.colmult;
    const uint16_t *colmode = P.wide ? P.wide[s_prob].colmode : Q.colmode;
    const int D = Q.D, q = Q.q, W = Q.W;
    // this thread's double-double totals: tot[slot * NT]
    double *tot = reinterpret_cast<double *>(smA + (size_t)(P.max_D + 1) * NCP) + threadIdx.x;
    {
        const int nelem = (D + 1) * NCP;
        if (P.U) {
            // gather this problem's minor straight from the shared matrix: row 0 =
            // pinned row, rows 1..D doubled (src/permanent_laplace.cpp:100-102);
            // padding columns are (1, 0, 0, ...) so that their s_j == 1
            for (int i = threadIdx.x; i < nelem; i += NT) {
                const int r = i / NCP, j = i - r * NCP;
                double2 v = make_double2(r == 0 ? 1.0 : 0.0, 0.0);
                if (j < Q.nc) {
                    v = P.U[(size_t)Q.rowmode[r] * P.ldu + colmode[j]];
                    if (r > 0) {
                        v.x *= 2.0;
                        v.y *= 2.0;
                    }
                }
                smA[i] = v;
            }
        } else {
            const double2 *src = P.A2 + Q.a_off;
            for (int i = threadIdx.x; i < nelem; i += NT)
                smA[i] = src[i];
        }
        // step table of the low counter: digit moved on the step into m, and
        // (-1)^m prod_{d<q} C(r_d, c_d(m))   (cf. pqperm_plan.cpp)
        for (int m = threadIdx.x; m < W; m += NT) {
            int rest = m, p = -1;
            double w = (m & 1) ? -1.0 : 1.0;
            for (int d = 0; d < q; d++) {
                const int L = Q.mult[d] + 1;
                const int c = rest % L;
                rest /= L;
                if (p < 0 && c != 0)
                    p = d;
                if (c != 0 && c != Q.mult[d])
                    w *= small_binom(Q.mult[d], c);
            }
            // m = 0 has no move; entry W (the move out of the last term, applied to
            // row sums nobody reads any more) is the pinned row with an empty bit
            s_step[m] = LapStep{(unsigned)((p < 0 ? 0 : p + 1) * NCP * (int)sizeof(double2)),
                                p < 0 ? 0u : 1u << p};
            s_wtab[m] = w;
        }
        if (threadIdx.x == 0)
            s_step[W] = LapStep{0u, 0u};
#pragma unroll 4
        for (int k = 0; k < kSlots; k++)
            tot[k * NT] = 0.0;
    }
    __syncthreads();

    const int h = threadIdx.x % S;                 // lane within the group
    const int groups_per_block = NT / S;
    const long long gstride = (long long)Q.nblocks * groups_per_block;
    // sums of the current segment, plain FP64; folded into `tot` after every segment
    double accr[NCL], acci[NCL], fullr = 0.0, fulli = 0.0;
#pragma unroll
    for (int j = 0; j < NCL; j++)
        accr[j] = acci[j] = 0.0;

    // All lanes of a warp run the same number of iterations (the group shuffles
    // below need the full warp); lanes past the last segment redo the last one
    // with weight 0.
    const int group_in_warp = (threadIdx.x & 31) / S;
    for (long long seg0 = (long long)((int)blockIdx.x - Q.first_block) * groups_per_block +
                          threadIdx.x / S;
         seg0 - group_in_warp < Q.nseg; seg0 += gstride) {
        const bool valid = seg0 < Q.nseg;
        const long long seg = Q.seg_begin + (valid ? seg0 : Q.nseg - 1);
        // ---- seed (same bookkeeping as seed_segment in pqperm_walk.cuh)
        double sr[NCL], si[NCL];
#pragma unroll
        for (int j = 0; j < NCL; j++) {
            const double2 a = smA[j * S + h];
            sr[j] = a.x;
            si[j] = a.y;
        }
        int odd = 0;
        double bin = 1.0;
        {
            uint8_t chain[kMaxDigits];
            unsigned long long rest = (unsigned long long)seg;
            for (int d = q; d < D; ++d) {
                const unsigned L = Q.mult[d] + 1u;
                if (rest >> 32) {
                    chain[d] = (uint8_t)(rest % L);
                    rest /= L;
                } else {
                    const unsigned r32 = (unsigned)rest;
                    chain[d] = (uint8_t)(r32 % L);
                    rest = r32 / L;
                }
            }
            for (int d = D - 1; d >= q; --d) {
                const int r = Q.mult[d];
                const int g = odd ? r - chain[d] : chain[d];
                odd ^= (g & 1);
                if (g != 0 && g != r)
                    bin *= small_binom(r, g);
                const double w = 0.5 * (double)(r - 2 * g);
                const double2 *row = smA + (d + 1) * NCP + h;
#pragma unroll
                for (int j = 0; j < NCL; j++) {
                    const double2 a = row[j * S];
                    sr[j] = __fma_rn(w, a.x, sr[j]);
                    si[j] = __fma_rn(w, a.y, si[j]);
                }
            }
        }
        unsigned dirmask = 0;
        for (int d = q - 1; d >= 0; --d) {
            const int r = Q.mult[d];
            dirmask |= (unsigned)odd << d;
            const double w = odd ? -0.5 * (double)r : 0.5 * (double)r;
            const double2 *row = smA + (d + 1) * NCP + h;
#pragma unroll
            for (int j = 0; j < NCL; j++) {
                const double2 a = row[j * S];
                sr[j] = __fma_rn(w, a.x, sr[j]);
                si[j] = __fma_rn(w, a.y, si[j]);
            }
            if (r & 1)
                odd = 0;
        }
        // weight of the segment's high digits; applied when the segment's sums are folded
        const double factor = valid ? (odd ? -bin : bin) : 0.0;

        // ---- walk
        // The move INTO term m+1 is applied column by column as soon as term m has
        // no further use for s_j: its 2*NCL independent FMAs and the row loads fill
        // latency gaps of the product chains instead of sitting in front of them.
        // Per term the integer side is one table entry (row offset, digit bit), one
        // bit test and one XOR: every instruction that is not FP64 still takes a
        // dispatch slot from the FP64 pipe (ncu: 62 of them per term cost 12 % before).
        const char *rows0 = reinterpret_cast<const char *>(smA + h);
#pragma unroll kLapUnroll
        for (int m = 0; m < W; ++m) {
            const double w = s_wtab[m];
            const LapStep st = s_step[m + 1];
            const double sg = (dirmask & st.bit) ? 1.0 : -1.0;
            dirmask ^= st.bit - 1u;
            const double2 *row = reinterpret_cast<const double2 *>(rows0 + st.rowoff);

            if constexpr (UNITCOLS) {
                // ---- product tree ---------------------------------------------------
                double nr[NCL], ni[NCL];
                tree_up<0, NCL, NCL>(sr, si, nr, ni);
                double lr, li; // product of this lane's columns
                tree_value<0, NCL, NCL>(sr, si, nr, ni, lr, li);
                double olr = 1.0, oli = 0.0; // other lanes (S > 1 only)
                if (S > 1)
                    others_product<S>(lr, li, olr, oli);
                if constexpr (MODE == kLapPerm) {
                    // batched permanents: the product of ALL columns is the term
                    if (S > 1)
                        cmul(lr, li, olr, oli);
                    fullr = __fma_rn(w, lr, fullr);
                    fulli = __fma_rn(w, li, fulli);
#pragma unroll
                    for (int j = 0; j < NCL; j++) {
                        const double2 a = row[j * S];
                        sr[j] = __fma_rn(sg, a.x, sr[j]);
                        si[j] = __fma_rn(sg, a.y, si[j]);
                    }
                } else if (S > 1) {
                    // outside of the lane's root: w * other lanes; w * the product of
                    // ALL columns is what a c_l = 0 column gets
                    const double wor = w * olr, woi = w * oli;
                    if constexpr (MODE == kLapLooFull)
                        cfma(fullr, fulli, wor, woi, lr, li);
                    if constexpr (NCL >= 2) {
                        tree_down<0, NCL, NCL, S, false>(sr, si, nr, ni, wor, woi, accr, acci,
                                                         row, sg);
                    } else {
                        accr[0] += wor;
                        acci[0] += woi;
                    }
                } else {
                    if constexpr (MODE == kLapLooFull) {
                        fullr = __fma_rn(w, lr, fullr);
                        fulli = __fma_rn(w, li, fulli);
                    }
                    if constexpr (NCL >= 2)
                        tree_down<0, NCL, NCL, S, true>(sr, si, nr, ni, w, 0.0, accr, acci, row,
                                                        sg);
                    else
                        accr[0] += w;
                }
                if constexpr (NCL == 1 && MODE != kLapPerm) {
                    const double2 a = row[0];
                    sr[0] = __fma_rn(sg, a.x, sr[0]);
                    si[0] = __fma_rn(sg, a.y, si[0]);
                }
            } else {
            // ---- column multiplicities: chunked suffix / prefix passes ------------
            // The lane's columns are cut into CH chunks whose suffix / prefix
            // chains are independent (interleaved below so that the FP64 pipe
            // sees CH chains at once); the term weight and the product of
            // everything OUTSIDE a chunk (other chunks, other lanes) are folded
            // into the start value of the chunk's prefix chain:
            //   acc_j += start_c * prod_{k<j in chunk} s_k^{c_k} * s_j^{c_j-1} * suf[j+1]
            double sufr[NCL], sufi[NCL]; // suf[j] = prod_{k >= j, k in chunk(j)} s_k^{c_k}
#pragma unroll
            for (int i = CLEN - 1; i >= 0; i--) {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int j0 = (NCL * c) / CH, j1 = (NCL * (c + 1)) / CH;
                    const int j = j0 + i;
                    if (j < j1) {
                        double tr = sr[j], ti = si[j];
                        const int cm = colmult[j * S + h];
                        for (int k = 1; k < cm; k++)
                            cmul(tr, ti, sr[j], si[j]);
                        if (j + 1 < j1)
                            cmul(tr, ti, sufr[j + 1], sufi[j + 1]);
                        sufr[j] = tr;
                        sufi[j] = ti;
                    }
                }
            }
            // lane-local leave-one-chunk-out products and the lane total (CH <= 2)
            double outr[CH], outi[CH], lr, li;
            lr = sufr[0];
            li = sufi[0];
            if (CH == 1) {
                outr[0] = 1.0;
                outi[0] = 0.0;
            } else {
                constexpr int F1 = NCL / 2; // first column of chunk 1
                outr[0] = sufr[F1];
                outi[0] = sufi[F1];
                outr[CH - 1] = sufr[0];
                outi[CH - 1] = sufi[0];
                cmul(lr, li, outr[0], outi[0]);
            }
            double olr = 1.0, oli = 0.0; // other lanes (S > 1 only)
            if (S > 1)
                others_product<S>(lr, li, olr, oli);
            if constexpr (MODE == kLapPerm) {
                if (S > 1)
                    cmul(lr, li, olr, oli);
                fullr = __fma_rn(w, lr, fullr);
                fulli = __fma_rn(w, li, fulli);
#pragma unroll
                for (int j = 0; j < NCL; j++) {
                    const double2 a = row[j * S];
                    sr[j] = __fma_rn(sg, a.x, sr[j]);
                    si[j] = __fma_rn(sg, a.y, si[j]);
                }
            } else {
            // start value of every chunk's prefix chain: w * other lanes * out_c.
            // Carrying w in the chain turns the accumulation into a complex FMA.
            double prer[CH], prei[CH];
            if (S > 1) {
                const double wor = w * olr, woi = w * oli;
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    prer[c] = wor;
                    prei[c] = woi;
                    if (CH > 1)
                        cmul(prer[c], prei[c], outr[c], outi[c]);
                }
                if constexpr (MODE == kLapLooFull)
                    cfma(fullr, fulli, wor, woi, lr, li);
            } else {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    prer[c] = w * outr[c];
                    prei[c] = w * outi[c];
                }
                if constexpr (MODE == kLapLooFull) {
                    fullr = __fma_rn(w, lr, fullr);
                    fulli = __fma_rn(w, li, fulli);
                }
            }
            // prefix chains, interleaved over the chunks
#pragma unroll
            for (int i = 0; i < CLEN; i++) {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int j0 = (NCL * c) / CH, j1 = (NCL * (c + 1)) / CH;
                    const int j = j0 + i;
                    if (j < j1) {
                        double pr = prer[c], pi = prei[c];
                        const int cm = colmult[j * S + h];
                        for (int k = 1; k < cm; k++)
                            cmul(pr, pi, sr[j], si[j]); // pre * s_j^{c_j - 1}
                        if (j + 1 < j1) {
                            cfma(accr[j], acci[j], pr, pi, sufr[j + 1], sufi[j + 1]);
                            cmul(pr, pi, sr[j], si[j]); // next start: pre * s_j^{c_j}
                            prer[c] = pr;
                            prei[c] = pi;
                        } else {
                            accr[j] += pr;
                            acci[j] += pi;
                        }
                        // s_j is not needed by this term any more: move to the next term
                        const double2 a = row[j * S];
                        sr[j] = __fma_rn(sg, a.x, sr[j]);
                        si[j] = __fma_rn(sg, a.y, si[j]);
                    }
                }
            }
            } // leave-one-out modes
            } // general flavour
        }
        // ---- fold the segment's sums, times the weight of its high digits, into the
        // double-double totals (the product acc * factor is split exactly)
        auto fold = [&](int slot, double v) {
            const double p = v * factor;
            dd_fold(tot, slot, NT, p);
            tot[(slot + 1) * NT] += __fma_rn(v, factor, -p);
        };
        if constexpr (MODE != kLapPerm) {
#pragma unroll
            for (int j = 0; j < NCL; j++) {
                fold(4 * j, accr[j]);
                fold(4 * j + 2, acci[j]);
                accr[j] = acci[j] = 0.0;
            }
        }
        if constexpr (MODE != kLapLoo) {
            fold(kFull, fullr);
            fold(kFull + 2, fulli);
            fullr = fulli = 0.0;
        }
    }

    // ---- CTA reduction (double-double): lanes with equal h hold the same columns.
    // Each warp reduces its totals with shuffles and parks the result in the slots
    // of its lanes 0..S-1; after the barrier NCP + 1 threads add the four warps.
    auto warp_reduce = [&](int slot, double restr, double resti) {
        dd re{tot[slot * NT], tot[(slot + 1) * NT]};
        dd im{tot[(slot + 2) * NT], tot[(slot + 3) * NT]};
        dd_add(re, restr); // what has not been folded yet
        dd_add(im, resti);
#pragma unroll
        for (int delta = 16; delta >= S; delta >>= 1) {
            dd_add(re, dd_shfl_down(re, delta));
            dd_add(im, dd_shfl_down(im, delta));
        }
        tot[slot * NT] = re.hi;
        tot[(slot + 1) * NT] = re.lo;
        tot[(slot + 2) * NT] = im.hi;
        tot[(slot + 3) * NT] = im.lo;
    };
    if constexpr (MODE != kLapPerm) {
#pragma unroll
        for (int j = 0; j < NCL; j++)
            warp_reduce(4 * j, accr[j], acci[j]);
    }
    warp_reduce(kFull, fullr, fulli);
    __syncthreads();
    const double *tot0 = reinterpret_cast<const double *>(smA + (size_t)(P.max_D + 1) * NCP);
    for (int k = threadIdx.x; k < NCP + 1; k += NT) {
        // compact column k = j * S + lane; the full product sits in lane 0's slots
        // (batched permanents: only the product is reduced and stored)
        if (MODE == kLapPerm && k < NCP)
            continue;
        const int slot = k < NCP ? 4 * (k / S) : kFull, lane = k < NCP ? k % S : 0;
        dd re{0.0, 0.0}, im{0.0, 0.0};
        for (int w = 0; w < NT / 32; w++) {
            const double *src = tot0 + w * 32 + lane;
            dd_add(re, dd{src[slot * NT], src[(slot + 1) * NT]});
            dd_add(im, dd{src[(slot + 2) * NT], src[(slot + 3) * NT]});
        }
        double *dst = P.partials + ((size_t)blockIdx.x * (NCP + 1) + k) * 4;
        dst[0] = re.hi;
        dst[1] = re.lo;
        dst[2] = im.hi;
        dst[3] = im.lo;
    }
}

// One warp per problem: sum the problem's CTA partials (double-double) in CTA
// order, scale by 2^-(sum_rows - 1) (src/permanent_laplace.cpp:226-235).
static __global__ void __launch_bounds__(128) laplace_reduce_kernel(const LapParams P, int ncp1)
{
    const int prob = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (prob >= P.nprob)
        return;
    const LapProblem &Q = P.prob[prob];
    const double scale = scalbn(1.0, -Q.exp2);
    // batched permanents: only the full product (index ncp1 - 1) was stored
    for (int k = (P.perm_only ? ncp1 - 1 : 0) + (threadIdx.x & 31); k < ncp1; k += 32) {
        dd re{0.0, 0.0}, im{0.0, 0.0};
        for (int b = 0; b < Q.nblocks; b++) {
            const double *v = P.partials + ((size_t)(Q.first_block + b) * ncp1 + k) * 4;
            dd_add(re, dd{v[0], v[1]});
            dd_add(im, dd{v[2], v[3]});
        }
        P.out[(size_t)prob * ncp1 + k] =
            make_double2((re.hi + re.lo) * scale, (im.hi + im.lo) * scale);
    }
}

// out[j] = result of compact column map[j] of problem 0 (map[j] = NCP selects the full
// product: a caller's zero-multiplicity column): the scatter the host does after a
// download, for results that are to stay on the device.
static __global__ void __launch_bounds__(128) laplace_scatter_kernel(const double2 *res,
                                                                     const int *map, int ncols,
                                                                     double2 *out)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j < ncols)
        out[j] = res[map[j]];
}

// Sampler epilogue (piquasso/_simulators/passive/sampling.py:736-747 of the
// reference): pmf[m] = | sum_j in_j * partial_j * U[m, nz_j] |^2 for the d output
// modes, from the Laplace results still in device memory.  One CTA per problem.
static __global__ void __launch_bounds__(128) sampler_pmf_kernel(const LapParams P, int ncp1,
                                                                 const double2 *U, int d,
                                                                 double *pmf)
{
    const LapProblem &Q = P.prob[blockIdx.x];
    const uint8_t *colmult = P.wide ? P.wide[blockIdx.x].colmult : Q.colmult;
    const uint16_t *colmode = P.wide ? P.wide[blockIdx.x].colmode : Q.colmode;
    __shared__ double2 w[kLapMaxCols];
    for (int k = threadIdx.x; k < Q.nc; k += 128) {
        const double2 v = P.out[(size_t)blockIdx.x * ncp1 + k];
        const double c = (double)colmult[k];
        w[k] = make_double2(c * v.x, c * v.y);
    }
    __syncthreads();
    for (int m = threadIdx.x; m < d; m += 128) {
        double ar = 0.0, ai = 0.0;
        const double2 *row = U + (size_t)m * d;
        for (int k = 0; k < Q.nc; k++) {
            const double2 u = row[colmode[k]];
            ar += u.x * w[k].x - u.y * w[k].y;
            ai += u.x * w[k].y + u.y * w[k].x;
        }
        pmf[(size_t)blockIdx.x * d + m] = ar * ar + ai * ai;
    }
}

// Draw from the pmf rows on the device: what the host does after
// _calculate_pmf (piquasso/_simulators/passive/sampling.py:736-753 of the
// reference), i.e. p = pmf / sum(pmf) (sequential sum) and numpy's
// Generator.choice(a, p=p) = searchsorted(cumsum(p) / cumsum(p)[-1], u, "right"),
// with the uniform variate u drawn by the caller from the shot's own generator.
// One thread per shot repeats numpy's operations in numpy's order (sequential
// adds, IEEE divisions; nothing here can be contracted into an FMA), so the
// index is the one the host would have computed from the same row.
static __global__ void __launch_bounds__(128) sampler_draw_kernel(const double *pmf, int n, int d,
                                                                  const double *u, int *index)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n)
        return;
    const double *row = pmf + (size_t)i * d;
    double total = 0.0;
    for (int m = 0; m < d; m++)
        total += row[m];
    double last = 0.0;
    for (int m = 0; m < d; m++)
        last += row[m] / total;
    const double ui = u[i];
    double c = 0.0;
    int idx = 0;
    for (int m = 0; m < d; m++) {
        c += row[m] / total;
        idx += (c / last <= ui) ? 1 : 0;
    }
    index[i] = (last == last) ? idx : -1; // NaN row (all-zero pmf): numpy raises
}

} // namespace pqperm
