// pqperm_permhyper.cuh -- batched permanents, hypercube flavour (sm_100a).
//
// The batched-permanent walk of pqperm_laplace.cuh (mode kLapPerm, one lane per
// Gray segment) reads one warp-uniform LDS.128 row operand per column and TERM; a
// uniform LDS.128 costs two shared-memory wavefronts and the LSU is shared by the
// SM's four schedulers, so that walk keeps the shared-memory pipe 57 % and the FP64
// pipe only 71 % busy (ncu, profiles/r02_ncu_summary.md).  Here the three lowest
// digits are BINARY rows (the host puts three rows of multiplicity 1 there; the sum
// over all Gray tuples does not depend on the digit order, src/permanent.cpp:218-250
// of the reference visits the same set) and a block of 2^3 terms is evaluated as in
// perm_walk_binary (pqperm_walk.cuh): column by column, the seven other vertices of
// the block's hypercube are one add each from its corner, multiplied straight into
// eight independent running products, and the same pass moves the corner to the next
// block.  Four operand fetches per column for eight terms instead of eight, eight-way
// ILP for the FP64 pipe, and the integer side (step table entry, direction bit) once
// per block.  The digits above the three walk as in the Laplace kernel: a step table
// per CTA for digits 3..q-1, one segment per thread for the rest.
//
// Unit column multiplicities only (single-photon inputs); everything else -- and
// problems with fewer than three rows of multiplicity 1 -- stays on the kLapPerm walk.
#pragma once

#include "pqperm_laplace.cuh"

namespace pqperm {

constexpr int kHyperB = kHyperDigits;  // binary digits evaluated as one hypercube
constexpr int kHyperTerms = 1 << kHyperB;

template <int NC>
__global__ void __launch_bounds__(kLapThreads) perm_hyper_kernel(const LapParams P)
{
    constexpr int NT = kLapThreads;
    constexpr int NP = kHyperTerms;
    extern __shared__ double2 smA[];              // (D+1) x NC, then the totals (4 slots)
    __shared__ double s_wtab[kLapMaxSegLen];      // per BLOCK of the low counter
    __shared__ LapStep s_step[kLapMaxSegLen + 1];
    __shared__ int s_prob;

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
    const int D = Q.D, q = Q.q;
    const int NB = Q.W >> kHyperB;                // blocks per segment
    double *tot = reinterpret_cast<double *>(smA + (size_t)(P.max_D + 1) * NC) + threadIdx.x;
    {
        // gather the minor from the shared matrix: row 0 = pinned row, rows 1..D
        // doubled (src/permanent.cpp:124-128); padding columns are (1, 0, 0, ...)
        const int nelem = (D + 1) * NC;
        for (int i = threadIdx.x; i < nelem; i += NT) {
            const int r = i / NC, j = i - r * NC;
            double2 v = make_double2(r == 0 ? 1.0 : 0.0, 0.0);
            if (j < Q.nc) {
                v = P.U[(size_t)Q.rowmode[r] * P.ldu + Q.colmode[j]];
                if (r > 0) {
                    v.x *= 2.0;
                    v.y *= 2.0;
                }
            }
            smA[i] = v;
        }
        // step table of the counter over digits 3..q-1, one entry per block: the digit
        // moved on the step into block mb and (-1)^mb prod_{3<=d<q} C(r_d, c_d(mb))
        for (int mb = threadIdx.x; mb < NB; mb += NT) {
            int rest = mb, p = -1;
            double w = (mb & 1) ? -1.0 : 1.0;
            for (int d = kHyperB; d < q; d++) {
                const int L = Q.mult[d] + 1;
                const int c = rest % L;
                rest /= L;
                if (p < 0 && c != 0)
                    p = d;
                if (c != 0 && c != Q.mult[d])
                    w *= small_binom(Q.mult[d], c);
            }
            s_step[mb] = LapStep{(unsigned)((p < 0 ? 0 : p + 1) * NC * (int)sizeof(double2)),
                                 p < 0 ? 0u : 1u << p};
            s_wtab[mb] = w;
        }
        if (threadIdx.x == 0)
            s_step[NB] = LapStep{0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; k++)
            tot[k * NT] = 0.0;
    }
    __syncthreads();

    const long long gstride = (long long)Q.nblocks * NT;
    double fullr = 0.0, fulli = 0.0;
    for (long long seg0 = (long long)((int)blockIdx.x - Q.first_block) * NT + threadIdx.x;
         seg0 < Q.nseg; seg0 += gstride) {
        const long long seg = Q.seg_begin + seg0;
        // ---- seed: corner of the segment's first block (the three lowest digits at
        // Gray value 0, digits 3..q-1 at counter 0, the rest from the segment index)
        double sr[NC], si[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double2 a = smA[j];
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
                const double2 *row = smA + (d + 1) * NC;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double2 a = row[j];
                    sr[j] = __fma_rn(w, a.x, sr[j]);
                    si[j] = __fma_rn(w, a.y, si[j]);
                }
            }
        }
        unsigned dirmask = 0;
        for (int d = q - 1; d >= 0; --d) {
            double w = 0.5; // the hypercube's corner: Gray value 0 of a binary digit
            if (d >= kHyperB) {
                const int r = Q.mult[d];
                dirmask |= (unsigned)odd << d;
                w = odd ? -0.5 * (double)r : 0.5 * (double)r;
                if (r & 1)
                    odd = 0;
            }
            const double2 *row = smA + (d + 1) * NC;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = row[j];
                sr[j] = __fma_rn(w, a.x, sr[j]);
                si[j] = __fma_rn(w, a.y, si[j]);
            }
        }
        // sign and weight of the digits >= 3 at the segment's first block
        const double factor = odd ? -bin : bin;

        // ---- walk, one hypercube per step
        const char *rows_base = reinterpret_cast<const char *>(smA);
        for (int mb = 0; mb < NB; ++mb) {
            const double w = s_wtab[mb];
            const LapStep st = s_step[mb + 1];
            const double sg = (dirmask & st.bit) ? 1.0 : -1.0;
            dirmask ^= st.bit - 1u;
            const double2 *rowh = reinterpret_cast<const double2 *>(rows_base + st.rowoff);
            // The three low rows are re-read in every block: kept in registers across
            // the loop (what the compiler does when it can prove them invariant) they
            // cost 12 registers per column and spill from 20 columns on.
            // (the opaque zero offset keeps the loads in the loop and in the shared
            // address space)
            unsigned opaque0 = 0;
            asm volatile("" : "+r"(opaque0));
            const double2 *row0 = smA + 1 * NC + opaque0;
            const double2 *row1 = row0 + NC, *row2 = row0 + 2 * NC;

            double pr[NP], pi[NP];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double c0r = sr[j], c0i = si[j];
                const double2 a0 = row0[j], a1 = row1[j], a2 = row2[j];
                // vertex e = vertex (e with its lowest set bit cleared) minus the doubled
                // row of that bit: every operand an exact matrix entry
                double vr[NP], vi[NP];
                vr[0] = c0r;
                vi[0] = c0i;
#pragma unroll
                for (int e = 1; e < NP; e++) {
                    const double2 a = (e & 1) ? a0 : ((e & 2) ? a1 : a2);
                    vr[e] = vr[e & (e - 1)] - a.x;
                    vi[e] = vi[e & (e - 1)] - a.y;
                }
#pragma unroll
                for (int e = 0; e < NP; e++) {
                    if (j == 0) {
                        pr[e] = vr[e];
                        pi[e] = vi[e];
                    } else {
                        cmul(pr[e], pi[e], vr[e], vi[e]);
                    }
                }
                // the move into the next block (the pinned row with weight -1 after the
                // last one: applied to row sums nobody reads any more)
                const double2 ah = rowh[j];
                sr[j] = __fma_rn(sg, ah.x, c0r);
                si[j] = __fma_rn(sg, ah.y, c0i);
            }
            // signed sum over the vertices, (-1)^{|e|}
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
            fullr = __fma_rn(w, br, fullr);
            fulli = __fma_rn(w, bi, fulli);
        }
        // ---- fold the segment's sum, times the weight of its high digits, into the
        // thread's double-double total (the product is split exactly)
        {
            const double p = fullr * factor;
            dd_fold(tot, 0, NT, p);
            tot[1 * NT] += __fma_rn(fullr, factor, -p);
            const double pq = fulli * factor;
            dd_fold(tot, 2, NT, pq);
            tot[3 * NT] += __fma_rn(fulli, factor, -pq);
            fullr = fulli = 0.0;
        }
    }

    // ---- CTA reduction (double-double) of the one sum
    {
        dd re{tot[0], tot[1 * NT]}, im{tot[2 * NT], tot[3 * NT]};
#pragma unroll
        for (int delta = 16; delta >= 1; delta >>= 1) {
            dd_add(re, dd_shfl_down(re, delta));
            dd_add(im, dd_shfl_down(im, delta));
        }
        tot[0] = re.hi;
        tot[1 * NT] = re.lo;
        tot[2 * NT] = im.hi;
        tot[3 * NT] = im.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *tot0 = reinterpret_cast<const double *>(smA + (size_t)(P.max_D + 1) * NC);
        dd re{0.0, 0.0}, im{0.0, 0.0};
        for (int w = 0; w < NT / 32; w++) {
            const double *src = tot0 + w * 32;
            dd_add(re, dd{src[0], src[1 * NT]});
            dd_add(im, dd{src[2 * NT], src[3 * NT]});
        }
        // same layout as the Laplace walk's partials: [CTA][NC + 1][4], product last
        double *dst = P.partials + ((size_t)blockIdx.x * (NC + 1) + NC) * 4;
        dst[0] = re.hi;
        dst[1] = re.lo;
        dst[2] = im.hi;
        dst[3] = im.lo;
    }
}

} // namespace pqperm
