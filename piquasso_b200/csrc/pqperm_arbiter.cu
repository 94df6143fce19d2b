// pqperm_arbiter.cu -- the permanent in double-double arithmetic (sm_100a).
//
// An ACCURACY ARBITER, not a fast path: every operation of the Glynn sum (row
// sums, complex products, term accumulation) is carried in double-double
// (~106-bit mantissa), so the result is good to ~1e-25 relative where the
// production walks -- and the reference's own double loop -- are at 1e-14..1e-9.
// It exists because the reference cannot arbitrate beyond n = 32 (its Gray
// counter truncates the offset to int, src/n_aryGrayCodeCounter.hpp:179) and a
// CPU arbiter in long double / binary128 takes hours there.
//
// It shares NOTHING with the production kernels beyond the host preprocessing
// (row split, compaction: pqperm_plan.cpp): the Gray-code counter below restates
// the reference's class literally (initialize(): n_aryGrayCodeCounter.hpp:170-194,
// next(): :209-254) with 64-bit offsets -- no direction masks, no hypercubes, no
// segment tables -- and the partition is a plain split of [0, idx_max) into
// contiguous chunks handed out by an atomic counter.
//
// Mapping: a group of 8 lanes owns one chunk; lane h holds the row sums of
// columns h, h+8, ... (<= 8 per lane) as double-double complex numbers, the
// lanes' products are combined by a shuffle butterfly, lane 0 accumulates.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pqperm.h"
#include "pqperm_ctx.h"
#include "pqperm_plan.h"

namespace pqperm {
namespace arb {

constexpr int kLanes = 8;
constexpr int kThreads = 128;
constexpr int kMaxColsPerLane = kMaxCols / kLanes;

struct Params {
    const double2 *A2;  // (D+1) x NC: row 0 pinned, rows 1..D doubled (pqperm_plan.cpp)
    unsigned long long *counter;
    double *partials;   // [grid][4]
    long long idx_max;
    long long chunk;    // offsets per work item
    int D, NC;
    uint8_t mult[kMaxDigits];
    uint8_t colmult[kMaxCols];
};

// ---- double-double arithmetic (Dekker / Knuth with FMA) -----------------------
struct dd2 {
    double hi, lo;
};

__host__ __device__ inline dd2 two_sum(double a, double b)
{
    const double s = a + b;
    const double bb = s - a;
    return dd2{s, (a - (s - bb)) + (b - bb)};
}

__host__ __device__ inline dd2 quick_two_sum(double a, double b) // |a| >= |b|
{
    const double s = a + b;
    return dd2{s, b - (s - a)};
}

__host__ __device__ inline dd2 add(dd2 a, dd2 b) // accurate (IEEE-style) addition
{
    dd2 s = two_sum(a.hi, b.hi);
    const dd2 t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}

__device__ __forceinline__ dd2 mul(dd2 a, dd2 b)
{
    const double p = a.hi * b.hi;
    double e = __fma_rn(a.hi, b.hi, -p);
    e = __fma_rn(a.hi, b.lo, e);
    e = __fma_rn(a.lo, b.hi, e);
    return quick_two_sum(p, e);
}

__device__ __forceinline__ dd2 mul_d(dd2 a, double b)
{
    const double p = a.hi * b;
    double e = __fma_rn(a.hi, b, -p);
    e = __fma_rn(a.lo, b, e);
    return quick_two_sum(p, e);
}

__device__ __forceinline__ dd2 neg(dd2 a) { return dd2{-a.hi, -a.lo}; }

struct cdd {
    dd2 re, im;
};

__device__ __forceinline__ cdd cmul(const cdd &a, const cdd &b)
{
    cdd r;
    r.re = add(mul(a.re, b.re), neg(mul(a.im, b.im)));
    r.im = add(mul(a.re, b.im), mul(a.im, b.re));
    return r;
}

__device__ __forceinline__ dd2 shfl_xor(dd2 a, int x, unsigned mask)
{
    return dd2{__shfl_xor_sync(mask, a.hi, x), __shfl_xor_sync(mask, a.lo, x)};
}

// exact product of a double and a small integer weight as a double-double
__device__ __forceinline__ dd2 exact_scale(double a, double w)
{
    const double p = a * w;
    return dd2{p, __fma_rn(a, w, -p)};
}

template <int NCL>
__global__ void __launch_bounds__(kThreads) arbiter_kernel(const Params P)
{
    const int h = threadIdx.x % kLanes;
    const int D = P.D, NC = P.NC;
    cdd acc{{0.0, 0.0}, {0.0, 0.0}};

    // shuffles stay inside the group: groups of one warp run different trip counts
    const int leader = (threadIdx.x & 31) & ~(kLanes - 1);
    const unsigned gmask = ((1u << kLanes) - 1u) << leader;
    for (;;) {
        // ---- next chunk of offsets for this group
        unsigned long long item = 0;
        if (h == 0)
            item = atomicAdd(P.counter, 1ull);
        item = __shfl_sync(gmask, item, leader);
        const long long begin = (long long)item * P.chunk;
        if (begin >= P.idx_max)
            break;
        long long end = begin + P.chunk; // exclusive
        if (end > P.idx_max)
            end = P.idx_max;

        // ---- initialize(offset): n_aryGrayCodeCounter.hpp:170-194, 64-bit offset
        uint8_t chain[kMaxDigits], gray[kMaxDigits];
        {
            unsigned long long rest = (unsigned long long)begin;
            for (int d = 0; d < D; d++) {
                const unsigned L = P.mult[d] + 1u;
                chain[d] = (uint8_t)(rest % L);
                rest /= L;
            }
            int parity = 0;
            for (int d = D - 1; d >= 0; d--) {
                const int g = parity ? P.mult[d] - chain[d] : chain[d];
                gray[d] = (uint8_t)g;
                parity ^= (g & 1);
            }
        }
        // ---- seed: s_j = a_0j + sum_d a_dj (r_d - 2 g_d)   (src/permanent.cpp:174-202)
        cdd s[NCL];
#pragma unroll
        for (int j = 0; j < NCL; j++) {
            const int col = j * kLanes + h;
            const double2 a = col < NC ? P.A2[col] : make_double2(1.0, 0.0);
            s[j].re = dd2{a.x, 0.0};
            s[j].im = dd2{a.y, 0.0};
        }
        int minus = 0;
        double binom = 1.0;
        for (int d = 0; d < D; d++) {
            const int r = P.mult[d], g = gray[d];
            minus += g;
            for (int i = 1; i <= g; i++) // C(r, g), exact in double for the sizes at hand
                binom = binom * (double)(r - g + i) / (double)i;
            const double w = 0.5 * (double)(r - 2 * g); // rows are stored doubled
#pragma unroll
            for (int j = 0; j < NCL; j++) {
                const int col = j * kLanes + h;
                if (col < NC) {
                    const double2 a = P.A2[(size_t)(d + 1) * NC + col];
                    s[j].re = add(s[j].re, exact_scale(a.x, w));
                    s[j].im = add(s[j].im, exact_scale(a.y, w));
                }
            }
        }
        double sign = (minus & 1) ? -1.0 : 1.0;

        const long long nterms = end - begin;
        for (long long t = 0; t < nterms; t++) {
            // ---- prod_j s_j^{c_j}: own columns, then the butterfly over the lanes
            cdd prod{{1.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int j = 0; j < NCL; j++) {
                const int col = j * kLanes + h;
                const int c = col < NC ? P.colmult[col] : 0;
                for (int k = 0; k < c; k++)
                    prod = cmul(prod, s[j]);
            }
#pragma unroll
            for (int x = 1; x < kLanes; x <<= 1) {
                cdd other;
                other.re = shfl_xor(prod.re, x, gmask);
                other.im = shfl_xor(prod.im, x, gmask);
                prod = cmul(prod, other);
            }
            if (h == 0) {
                const double w = sign * binom;
                acc.re = add(acc.re, mul_d(prod.re, w));
                acc.im = add(acc.im, mul_d(prod.im, w));
            }
            if (t + 1 >= nterms)
                break;
            // ---- next(): n_aryGrayCodeCounter.hpp:209-254
            int i = 0;
            while (i < D && chain[i] == P.mult[i]) {
                chain[i] = 0;
                i++;
            }
            // i < D: t + 1 < nterms keeps the counter inside [0, idx_max)
            chain[i]++;
            int parity = 0, changed = 0, prev = 0, value = 0;
            for (int d = D - 1; d >= 0; d--) {
                const int g = parity ? P.mult[d] - chain[d] : chain[d];
                parity ^= (g & 1);
                if (g != gray[d]) {
                    changed = d;
                    prev = gray[d];
                    value = g;
                    gray[d] = (uint8_t)g;
                    break;
                }
            }
            // ---- src/permanent.cpp:226-247: row-sum update, sign, binomial
            const double w = (double)(prev - value); // +-1, rows are stored doubled
#pragma unroll
            for (int j = 0; j < NCL; j++) {
                const int col = j * kLanes + h;
                if (col < NC) {
                    const double2 a = P.A2[(size_t)(changed + 1) * NC + col];
                    s[j].re = add(s[j].re, dd2{w * a.x, 0.0});
                    s[j].im = add(s[j].im, dd2{w * a.y, 0.0});
                }
            }
            sign = -sign;
            const int r = P.mult[changed];
            binom = value < prev ? binom * (double)prev / (double)(r - value)
                                 : binom * (double)(r - prev) / (double)value;
        }
    }

    // ---- block reduction: lane-0 threads hold the sums
    __shared__ double red[kThreads / kLanes][4];
    if (h == 0) {
        double *dst = red[threadIdx.x / kLanes];
        dst[0] = acc.re.hi;
        dst[1] = acc.re.lo;
        dst[2] = acc.im.hi;
        dst[3] = acc.im.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dd2 re{0.0, 0.0}, im{0.0, 0.0};
        for (int g = 0; g < kThreads / kLanes; g++) {
            re = add(re, dd2{red[g][0], red[g][1]});
            im = add(im, dd2{red[g][2], red[g][3]});
        }
        double *dst = P.partials + 4 * (size_t)blockIdx.x;
        dst[0] = re.hi;
        dst[1] = re.lo;
        dst[2] = im.hi;
        dst[3] = im.lo;
    }
}

template <int NCL>
static cudaError_t launch_ncl(const Params &P, int grid, cudaStream_t stream)
{
    arbiter_kernel<NCL><<<grid, kThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

static cudaError_t launch(int ncl, const Params &P, int grid, cudaStream_t stream)
{
    switch (ncl) {
    case 1: return launch_ncl<1>(P, grid, stream);
    case 2: return launch_ncl<2>(P, grid, stream);
    case 3: return launch_ncl<3>(P, grid, stream);
    case 4: return launch_ncl<4>(P, grid, stream);
    case 5: return launch_ncl<5>(P, grid, stream);
    case 6: return launch_ncl<6>(P, grid, stream);
    case 7: return launch_ncl<7>(P, grid, stream);
    case 8: return launch_ncl<8>(P, grid, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace arb
} // namespace pqperm

using namespace pqperm;

extern "C" int pq_perm_arbiter_c128(const double *A, int R, int C, const int32_t *rows,
                                    const int32_t *cols, double out[4])
{
    if (!out || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    Plan plan;
    std::string err;
    PlanOptions o;
    o.kernel_choice = 1; // no column expansion games: the generic layout
    int rc = make_plan(A, R, C, rows, cols, o, plan, err);
    if (rc)
        return fail(rc, err);
    if (plan.trivial) {
        out[0] = plan.triv[0];
        out[1] = 0.0;
        out[2] = plan.triv[1];
        out[3] = 0.0;
        return PQ_OK;
    }
    DeviceCtx *c = nullptr;
    if ((rc = ctx_get(g_devices[0], &c)))
        return rc;
    std::lock_guard<std::mutex> dev_lock(c->mu);
    // the plan's matrix is (D+1) x NCP with NCP >= NC padded columns of (1, 0, ...)
    arb::Params P;
    std::memset(&P, 0, sizeof(P));
    P.D = plan.D;
    P.NC = plan.NCP;
    P.idx_max = plan.idx_max;
    for (int d = 0; d < plan.D; d++)
        P.mult[d] = (uint8_t)plan.mult[d];
    for (int j = 0; j < plan.NCP; j++)
        P.colmult[j] = (uint8_t)plan.colmult[j];
    const int ncl = (plan.NCP + arb::kLanes - 1) / arb::kLanes;
    if (ncl > arb::kMaxColsPerLane)
        return fail(PQ_ERR_TOO_LARGE, "arbiter: more than 64 columns");
    const int groups_per_block = arb::kThreads / arb::kLanes;
    const int grid_max = c->num_sms * 4;
    // chunks: long enough to amortise the seed, short enough to balance the groups
    const long long groups = (long long)grid_max * groups_per_block;
    long long chunk = plan.idx_max / (groups * 8);
    chunk = std::max<long long>(64, std::min<long long>(chunk, 1LL << 16));
    P.chunk = chunk;
    const long long items = (plan.idx_max + chunk - 1) / chunk;
    const int grid = (int)std::max<long long>(
        1, std::min<long long>(grid_max, (items + groups_per_block - 1) / groups_per_block));
    const size_t a2_bytes = plan.A2.size() * sizeof(double);
    if ((rc = grow_dev(c, 1, a2_bytes)) || (rc = grow_dev(c, 2, (size_t)grid * 4 * sizeof(double))) ||
        (rc = grow_host(c, 2, (size_t)grid * 4 * sizeof(double))))
        return rc;
    cudaStream_t st = c->stream;
    if (c->busy_valid) // a walk on a caller's stream may still be drawing from the dispenser
        PQ_CUDA(cudaStreamWaitEvent(st, c->ev_busy, 0));
    PQ_CUDA(cudaMemcpyAsync(c->d_lap[1], plan.A2.data(), a2_bytes, cudaMemcpyHostToDevice, st));
    PQ_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
    P.A2 = reinterpret_cast<const double2 *>(c->d_lap[1]);
    P.counter = c->d_counter;
    P.partials = reinterpret_cast<double *>(c->d_lap[2]);
    PQ_CUDA(cudaEventRecord(c->lap_ev0, st));
    cudaError_t e = arb::launch(ncl, P, grid, st);
    if (e != cudaSuccess)
        return fail_cuda(e, "launch arbiter_kernel");
    g_launches += 1;
    PQ_CUDA(cudaEventRecord(c->lap_ev1, st));
    // the walk kernels expect their dispenser at zero between launches
    PQ_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), st));
    PQ_CUDA(cudaMemcpyAsync(c->h_lap[2], c->d_lap[2], (size_t)grid * 4 * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
    PQ_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->lap_ev0, c->lap_ev1) == cudaSuccess)
        c->last_kernel_ms = ms;
    const double *hp = reinterpret_cast<const double *>(c->h_lap[2]);
    arb::dd2 re{0.0, 0.0}, im{0.0, 0.0};
    for (int b = 0; b < grid; b++) {
        re = arb::add(re, arb::dd2{hp[4 * b], hp[4 * b + 1]});
        im = arb::add(im, arb::dd2{hp[4 * b + 2], hp[4 * b + 3]});
    }
    // src/permanent.cpp:259: exact power-of-two scaling of both parts
    const int e2 = -(plan.sum_rows - 1);
    out[0] = std::ldexp(re.hi, e2);
    out[1] = std::ldexp(re.lo, e2);
    out[2] = std::ldexp(im.hi, e2);
    out[3] = std::ldexp(im.lo, e2);
    return PQ_OK;
}
