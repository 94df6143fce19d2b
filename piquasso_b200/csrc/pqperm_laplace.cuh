// pqperm_laplace.cuh -- batched Laplace-expansion walk (sm_100a).
//
// Replaces the hot loop of permanent_laplace_cpp, src/permanent_laplace.cpp:
// 191-223 of the reference: the same Gray-code walk as the permanent, but every
// term contributes to C sums, sum l getting the product with ONE factor of
// column l removed (Lemma 1 of arXiv:2005.04214):
//   out_l * 2^(N-1) = sum_offset (-1)^{sum g} prod_d C(r_d,g_d)
//                       * s_l^{c_l-1} prod_{k != l} s_k^{c_k}
// The reference recomputes each of the C products from scratch (O(C*M) per
// term); here one suffix pass + one prefix pass give all of them: per column two
// complex multiplies and one complex multiply-add into the accumulator (the
// term weight rides in the prefix chain), 14 FP64 instructions with the row-sum
// update.
//
// Batch layout: one launch walks many independent problems (the sampler's
// (shot, photon) problems).  A CTA works on one problem; S adjacent lanes
// ("group") share one Gray segment and split the columns (lane h owns columns
// h, h+S, ...), so that row sums, suffix products and accumulators of up to 64
// columns stay in registers.
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

template <int NCL, int S, bool UNITCOLS>
__global__ void __launch_bounds__(kLapThreads) laplace_walk_kernel(const LapParams P)
{
    constexpr int NCP = NCL * S;
    constexpr int NT = kLapThreads;
    // independent suffix / prefix chains per lane.  Two are enough once the
    // next term's row-sum update is interleaved with the prefix pass (measured
    // on B200: k=25 walk 0.467 ms per 2^23 terms with 3 chunks, 0.453 with 2);
    // every extra chunk costs 2-3 complex multiplies per term.
    constexpr int CH = NCL >= 4 ? 2 : 1;
    constexpr int CLEN = (NCL + CH - 1) / CH;              // longest chunk
    extern __shared__ double2 smA[];              // (D+1) x NCP
    __shared__ double s_wtab[kLapMaxSegLen];
    __shared__ uint8_t s_sched[kLapMaxSegLen];
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
            s_sched[m] = (uint8_t)(p < 0 ? 0 : p);
            s_wtab[m] = w;
        }
    }
    __syncthreads();

    const int h = threadIdx.x % S;                 // lane within the group
    const int groups_per_block = NT / S;
    const long long gstride = (long long)Q.nblocks * groups_per_block;
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
        const double factor = valid ? (odd ? -bin : bin) : 0.0;

        // ---- walk
        // The move INTO term m+1 is applied column by column at the end of term
        // m's prefix pass (right after the last use of s_j): its 2*NCL independent
        // FMAs and the row loads then fill the latency gaps of the dependent
        // prefix chains instead of sitting in front of the suffix pass.
        int p_next = W > 1 ? s_sched[1] : 0;
        for (int m = 0; m < W; ++m) {
            const double w = factor * s_wtab[m];
            const bool more = m + 1 < W;
            const int p = p_next;
            const double sg = more ? (((dirmask >> p) & 1u) ? 1.0 : -1.0) : 0.0;
            if (more)
                dirmask ^= (1u << p) - 1u;
            p_next = m + 2 < W ? s_sched[m + 2] : 0;
            // row moved on the next step; after the last term the pinned row (always
            // present, finite) times sg = 0 leaves s untouched
            const double2 *row = smA + (more ? p + 1 : 0) * NCP + h;

            // ---- all C leave-one-out products of this term -------------------
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
                        if (!UNITCOLS) {
                            const int cm = colmult[j * S + h];
                            for (int k = 1; k < cm; k++)
                                cmul(tr, ti, sr[j], si[j]);
                        }
                        if (j + 1 < j1)
                            cmul(tr, ti, sufr[j + 1], sufi[j + 1]);
                        sufr[j] = tr;
                        sufi[j] = ti;
                    }
                }
            }
            // lane-local leave-one-chunk-out products out_c = prod_{c' != c} T_c'
            // (T_c = suf[first column of chunk c]) and the lane total
            double outr[CH], outi[CH], lr, li;
            {
                double pTr[CH], pTi[CH]; // prod_{c' < c} T_c'   (c >= 1)
                double sTr[CH], sTi[CH]; // prod_{c' > c} T_c'   (c <= CH-2)
#pragma unroll
                for (int c = 1; c < CH; c++) {
                    const int f = (NCL * (c - 1)) / CH;
                    pTr[c] = sufr[f];
                    pTi[c] = sufi[f];
                    if (c > 1)
                        cmul(pTr[c], pTi[c], pTr[c - 1], pTi[c - 1]);
                }
#pragma unroll
                for (int c = CH - 2; c >= 0; c--) {
                    const int f = (NCL * (c + 1)) / CH;
                    sTr[c] = sufr[f];
                    sTi[c] = sufi[f];
                    if (c < CH - 2)
                        cmul(sTr[c], sTi[c], sTr[c + 1], sTi[c + 1]);
                }
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    if (CH == 1) {
                        outr[c] = 1.0;
                        outi[c] = 0.0;
                    } else if (c == 0) {
                        outr[c] = sTr[0];
                        outi[c] = sTi[0];
                    } else if (c == CH - 1) {
                        outr[c] = pTr[c];
                        outi[c] = pTi[c];
                    } else {
                        outr[c] = pTr[c];
                        outi[c] = pTi[c];
                        cmul(outr[c], outi[c], sTr[c], sTi[c]);
                    }
                }
                lr = sufr[0];
                li = sufi[0];
                if (CH > 1)
                    cmul(lr, li, outr[0], outi[0]);
            }
            double olr = 1.0, oli = 0.0; // other lanes (S > 1 only)
            if (S > 1) {
#pragma unroll
                for (int x = 1; x < S; x++) {
                    const double orr = __shfl_xor_sync(0xffffffffu, lr, x);
                    const double oi = __shfl_xor_sync(0xffffffffu, li, x);
                    if (x == 1) {
                        olr = orr;
                        oli = oi;
                    } else {
                        cmul(olr, oli, orr, oi);
                    }
                }
            }
            if (P.perm_only) {
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
                continue;
            }
            // start value of every chunk's prefix chain: w * other lanes * out_c.
            // Carrying w in the chain turns the accumulation into a complex FMA
            // (acc += pre * suf: 4 DFMA instead of a complex multiply + 2 DFMA).
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
                // w * the product of ALL columns: what a c_l = 0 column gets
                cfma(fullr, fulli, wor, woi, lr, li);
            } else {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    prer[c] = w * outr[c];
                    prei[c] = w * outi[c];
                }
                fullr = __fma_rn(w, lr, fullr);
                fulli = __fma_rn(w, li, fulli);
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
                        if (!UNITCOLS) {
                            const int cm = colmult[j * S + h];
                            for (int k = 1; k < cm; k++)
                                cmul(pr, pi, sr[j], si[j]); // pre * s_j^{c_j - 1}
                        }
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
        }
    }

    // ---- CTA reduction: lanes with equal h hold the same columns
    __shared__ double red[(NT / 32) * (NCP + 1) * 2];
#pragma unroll
    for (int j = 0; j < NCL; j++) {
#pragma unroll
        for (int delta = 16; delta >= S; delta >>= 1) {
            accr[j] += __shfl_down_sync(0xffffffffu, accr[j], delta);
            acci[j] += __shfl_down_sync(0xffffffffu, acci[j], delta);
        }
    }
#pragma unroll
    for (int delta = 16; delta >= S; delta >>= 1) {
        fullr += __shfl_down_sync(0xffffffffu, fullr, delta);
        fulli += __shfl_down_sync(0xffffffffu, fulli, delta);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < S) {
        double *dst = red + warp * (NCP + 1) * 2;
#pragma unroll
        for (int j = 0; j < NCL; j++) {
            dst[(j * S + lane) * 2] = accr[j];
            dst[(j * S + lane) * 2 + 1] = acci[j];
        }
        if (lane == 0) {
            dst[NCP * 2] = fullr;
            dst[NCP * 2 + 1] = fulli;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NCP + 1; k += NT) {
        double re = 0.0, im = 0.0;
        for (int w = 0; w < NT / 32; w++) {
            re += red[(w * (NCP + 1) + k) * 2];
            im += red[(w * (NCP + 1) + k) * 2 + 1];
        }
        P.partials[(size_t)blockIdx.x * (NCP + 1) + k] = make_double2(re, im);
    }
}

// One warp per problem: sum the problem's CTA partials in CTA order, scale by
// 2^-(sum_rows - 1) (src/permanent_laplace.cpp:226-235).
static __global__ void __launch_bounds__(128) laplace_reduce_kernel(const LapParams P, int ncp1)
{
    const int prob = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (prob >= P.nprob)
        return;
    const LapProblem &Q = P.prob[prob];
    const double scale = scalbn(1.0, -Q.exp2);
    for (int k = threadIdx.x & 31; k < ncp1; k += 32) {
        double re = 0.0, im = 0.0;
        for (int b = 0; b < Q.nblocks; b++) {
            const double2 v = P.partials[(size_t)(Q.first_block + b) * ncp1 + k];
            re += v.x;
            im += v.y;
        }
        P.out[(size_t)prob * ncp1 + k] = make_double2(re * scale, im * scale);
    }
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
