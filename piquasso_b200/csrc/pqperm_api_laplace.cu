// pqperm_api_laplace.cu -- host side of permanent_laplace: single call, batch
// and the sampler's photon step (kernels: pqperm_laplace.cuh).
//
// A batch is described problem by problem with a lean planner (no allocation
// per problem), bucketed by kernel variant (lanes per segment S, columns per
// lane NCL, unit-column flavour), and each bucket is one walk launch + one
// reduce launch (+ the pmf epilogue for the sampler).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "pqperm_ctx.h"
#include "pqperm_launch.h"

using namespace pqperm;

namespace {

// Terms per segment: every segment is seeded by a direct evaluation (O(D*C) FMAs
// behind shared-memory latency plus the decode of its Gray digits), which costs
// ~10 % at 64 terms; long segments are used where the problem still yields
// thousands of them (measured: k=25 walk 0.465 -> 0.388 ms).
constexpr int kLapSegLen = 64;
constexpr int kLapSegLenBig = kLapMaxSegLen;          // 512
constexpr long long kLapBigProblem = 1LL << 18;       // terms
// the sampler step plans its shots on up to this many host threads ...
constexpr int kPlanThreadsMax = 8;
// ... as long as every thread gets at least this many shots
constexpr int kPlanShotsPerThread = 256;

// Finer split of the device phase of the calling thread's last sampler step (ms):
// [0] scratch growth, [1] staging into pinned memory, [2] enqueueing (copies,
// launches), [3] waiting for the stream, [4] scattering results, [5] uploading U.
thread_local double g_sampler_detail[8] = {0, 0, 0, 0, 0, 0, 0, 0};
struct Lap {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void to(double &slot)
    {
        const auto now = std::chrono::steady_clock::now();
        slot += std::chrono::duration<double, std::milli>(now - t).count();
        t = now;
    }
};

// What the lean planner extracts from one problem's multiplicity vectors
// (zeros allowed).  Restates src/permanent_laplace.cpp:49-118 of the reference
// plus the compaction of pqperm_plan.cpp.
struct LapShape {
    int trivial;             // reference early-out: result is [1]
    int D, NC, M;            // digits, active columns, sum of column multiplicities
    int pinned;              // row pinned to delta = +1
    int sum_rows;
    bool unit;
    bool hyper;              // batched permanents: three binary digits moved to the front
    int src_row[kMaxDigits]; // row feeding digit d
    int mult[kMaxDigits];
    int src_col[kLapMaxCols];
    int colmult[kLapMaxCols];
};

int lap_shape(int R, int C, const int32_t *rows, const int32_t *cols, LapShape &sh,
              std::string &err)
{
    long long sr = 0, sc = 0;
    for (int i = 0; i < R; i++) {
        if (rows[i] < 0) {
            err = "negative row multiplicity";
            return PQ_ERR_BAD_ARG;
        }
        sr += rows[i];
    }
    for (int j = 0; j < C; j++) {
        if (cols[j] < 0) {
            err = "negative column multiplicity";
            return PQ_ERR_BAD_ARG;
        }
        sc += cols[j];
    }
    sh.sum_rows = (int)sr;
    sh.hyper = false;
    sh.trivial = (R == 0 || C == 0 || sr == 0 || sc == 0); // src/permanent_laplace.cpp:52-57
    if (sh.trivial)
        return PQ_OK;
    int min_idx = 0, minelem = 0; // :59-69, first smallest non-zero multiplicity
    for (int i = 0; i < R; i++)
        if (minelem == 0 || (rows[i] < minelem && rows[i] != 0)) {
            minelem = rows[i];
            min_idx = i;
        }
    sh.pinned = min_idx;
    sh.D = 0;
    double idx_max = 1.0;
    for (int i = 0; i < R; i++) {
        const int r = rows[i] - (i == min_idx ? 1 : 0);
        if (r == 0)
            continue;
        if (r > kMaxMultiplicity || sh.D >= kMaxDigits) {
            err = "more than 64 active rows or a multiplicity above 254";
            return PQ_ERR_TOO_LARGE;
        }
        sh.src_row[sh.D] = i;
        sh.mult[sh.D] = r;
        sh.D++;
        idx_max *= (r + 1);
    }
    if (idx_max > 4.6e18) {
        err = "term space exceeds 2^62";
        return PQ_ERR_TOO_LARGE;
    }
    sh.NC = 0;
    sh.M = 0;
    sh.unit = true;
    for (int j = 0; j < C; j++) {
        if (cols[j] == 0)
            continue;
        if (cols[j] > kMaxMultiplicity || sh.NC >= kLapMaxCols) {
            err = "more than 256 active columns or a multiplicity above 254";
            return PQ_ERR_TOO_LARGE;
        }
        sh.src_col[sh.NC] = j;
        sh.colmult[sh.NC] = cols[j];
        sh.M += cols[j];
        if (cols[j] != 1)
            sh.unit = false;
        sh.NC++;
    }
    return PQ_OK;
}

// Batched permanents: a column of multiplicity c is written out as c unit columns while
// the expanded width fits one lane (what pqperm_plan.cpp does for the single permanent):
// s_j^c becomes c factors of the product tree instead of a run-time loop.
void lap_expand_columns(LapShape &sh)
{
    if (sh.unit || sh.M > kPermS1MaxCols)
        return;
    int src[kPermS1MaxCols], n = 0;
    for (int j = 0; j < sh.NC; j++)
        for (int k = 0; k < sh.colmult[j]; k++)
            src[n++] = sh.src_col[j];
    for (int j = 0; j < n; j++) {
        sh.src_col[j] = src[j];
        sh.colmult[j] = 1;
    }
    sh.NC = n;
    sh.unit = true;
}

// Batched permanents, hypercube flavour (pqperm_permhyper.cuh): with unit columns and
// at least three rows of multiplicity 1, the first three such rows become digits 0..2
// (the other digits keep their order).  The sum over all Gray tuples does not depend
// on the order of the digits; only the order of the additions changes.
bool lap_make_hyper(LapShape &sh)
{
    if (!sh.unit || sh.NC < kHyperMinCols || sh.NC > kPermS1MaxCols)
        return false;
    int pos[kHyperDigits], found = 0;
    for (int d = 0; d < sh.D && found < kHyperDigits; d++)
        if (sh.mult[d] == 1)
            pos[found++] = d;
    if (found < kHyperDigits)
        return false;
    int mult[kMaxDigits], src[kMaxDigits], n = 0;
    for (int k = 0; k < kHyperDigits; k++) {
        mult[n] = 1;
        src[n++] = sh.src_row[pos[k]];
    }
    for (int d = 0; d < sh.D; d++) {
        bool taken = false;
        for (int k = 0; k < kHyperDigits; k++)
            taken = taken || pos[k] == d;
        if (!taken) {
            mult[n] = sh.mult[d];
            src[n++] = sh.src_row[d];
        }
    }
    for (int d = 0; d < sh.D; d++) {
        sh.mult[d] = mult[d];
        sh.src_row[d] = src[d];
    }
    sh.hyper = true;
    return true;
}

// Fills the walk parameters of a described problem: digits 0..q-1 are walked
// inside a segment of W <= kLapSegLen terms, the rest index the segments.
// Problems wider than kMaxCols keep their column tables in `wide`.
// min_segs > 1 (batched permanents): small problems are cut into shorter segments
// so that they still yield that many of them -- a CTA's threads each take one.
void lap_fill(const LapShape &sh, int ncp, LapProblem &q, LapWide *wide = nullptr,
              int min_segs = 1)
{
    std::memset(&q, 0, sizeof(q));
    q.D = sh.D;
    long long total = 1;
    for (int d = 0; d < sh.D; d++)
        total *= (sh.mult[d] + 1);
    int seglen = total >= kLapBigProblem ? kLapSegLenBig : kLapSegLen;
    // batched permanents (min_segs > 1): the segment count is bounded from below anyway, so
    // mid-size problems take the long segments too (a seed costs ~20 terms of a 16-column walk)
    static const bool long_small = [] {
        const char *e = std::getenv("PQ_BATCH_LONG_SEGS");
        return e ? std::atoi(e) != 0 : true;
    }();
    if (min_segs > 1 && long_small)
        seglen = kLapSegLenBig;
    if (sh.hyper && (total >= kLapBigProblem || (min_segs > 1 && long_small))) {
        // the hypercube flavour tabulates BLOCKS of 8 terms: up to 8 x 256 terms per segment
        // (B200 sweep, n = 24 batches: 256 terms 10.6 ms, 512 10.2, 1024 10.1, 2048 10.2;
        // n = 20: 8.14 / 8.09 / 8.19 / 8.87)
        static const int hyper_seglen = [] {
            const char *e = std::getenv("PQ_HYPER_SEGLEN");
            const int v = e ? std::atoi(e) : 512;
            return std::max(64, std::min(v, 8 * kLapMaxSegLen));
        }();
        seglen = hyper_seglen;
    }
    int qd = 0;
    long long W = 1;
    for (int d = 0; d < sh.D; d++) {
        // (hypercube flavour: the three binary digits of a block are always low)
        if (qd == d && qd < kMaxLowDigits &&
            ((sh.hyper && d < kHyperDigits) ||
             (W * (sh.mult[d] + 1) <= seglen &&
              (min_segs <= 1 || total / (W * (sh.mult[d] + 1)) >= min_segs)))) {
            W *= (sh.mult[d] + 1);
            qd++;
        }
        q.mult[d] = (uint8_t)sh.mult[d];
    }
    q.q = qd;
    q.W = (int)W;
    q.nseg = total / W;
    q.exp2 = sh.sum_rows - 1;
    q.nc = sh.NC;
    if (wide) {
        std::memset(wide, 0, sizeof(*wide));
        for (int j = 0; j < ncp; j++)
            wide->colmult[j] = (uint8_t)(j < sh.NC ? sh.colmult[j] : 1);
        for (int j = 0; j < sh.NC; j++)
            wide->colmode[j] = (uint16_t)sh.src_col[j];
    } else {
        for (int j = 0; j < ncp; j++)
            q.colmult[j] = (uint8_t)(j < sh.NC ? sh.colmult[j] : 1);
        for (int j = 0; j < sh.NC; j++)
            q.colmode[j] = (uint16_t)sh.src_col[j];
    }
}

// The (D+1) x NCP matrix is staged per CTA beside the threads' double-double
// totals (lap_smem_bytes, pqperm_limits.h); wide problems with many rows do not fit.
const char *const kLapTooWide =
    "problem too large for the lane-split walk: (active rows + 1) x padded columns x 16 bytes "
    "plus the accumulators exceeds 224 KB of shared memory";

struct Bucket {
    int S = 0, NCL = 0;
    bool unit = false;
    bool hyper = false;       // batched permanents: hypercube flavour
    std::vector<LapProblem> probs;
    std::vector<LapWide> wide; // S = 32 only: column tables, aligned with probs
    std::vector<double> a2;   // packed mode
    int max_D = 0;
    bool need_full = false;   // some problem has a zero-multiplicity column (it gets the full product)
};

// buckets[((S index) * 33 + NCL) * 3 + flavour], flavour 0 general columns, 1 unit
// columns, 2 unit columns + hypercube
struct Buckets {
    std::vector<Bucket> b;
    Buckets() : b(4 * 33 * 3) {}
    Bucket &get(const LapVariant &v, bool unit, bool hyper = false)
    {
        const int si = v.S == 1 ? 0 : (v.S == 2 ? 1 : (v.S == 4 ? 2 : 3));
        Bucket &k = b[(si * 33 + v.NCL) * 3 + (hyper ? 2 : (unit ? 1 : 0))];
        k.S = v.S;
        k.NCL = v.NCL;
        k.unit = unit;
        k.hyper = hyper;
        return k;
    }
    void reset()
    {
        for (Bucket &k : b) {
            k.probs.clear();
            k.wide.clear();
            k.a2.clear();
            k.max_D = 0;
            k.need_full = false;
        }
    }
};
// per calling thread (the sampler step plans outside the library lock); keeps its
// capacity between calls
thread_local Buckets g_buckets;

// One planner thread's share of a batch.
struct PlanPart {
    Buckets buckets;
    int rc = PQ_OK;
    std::string err;
    bool any = false;
};

// Plans problems [0, n) into g_buckets: plan_range(begin, end, buckets, part) fills
// `buckets` for its range.  With thousands of problems the range is cut over a few
// host threads (PQ_PLAN_THREADS caps them, 1 = the calling thread), each filling its
// own buckets, merged in index order.
template <typename F>
int plan_parallel(int n, F &plan_range, bool &any)
{
    g_buckets.reset();
    any = false;
    static const int thread_cap = [] {
        const char *env = std::getenv("PQ_PLAN_THREADS");
        const int v = env ? std::atoi(env) : kPlanThreadsMax;
        return std::max(1, std::min(v, 64));
    }();
    const int nthreads = (int)std::min<long long>(
        {(long long)thread_cap, (long long)n / kPlanShotsPerThread,
         (long long)std::max(1u, std::thread::hardware_concurrency())});
    if (nthreads <= 1) {
        PlanPart part;
        plan_range(0, n, g_buckets, part);
        if (part.rc)
            return fail(part.rc, part.err);
        any = part.any;
        return PQ_OK;
    }
    static thread_local std::vector<PlanPart> parts; // keeps the buckets' capacity
    if ((int)parts.size() < nthreads)
        parts.resize(nthreads);
    std::vector<std::thread> workers;
    for (int t = 0; t < nthreads; t++) {
        parts[t].buckets.reset();
        parts[t].rc = PQ_OK;
        parts[t].any = false;
        const int begin = (int)((long long)n * t / nthreads);
        const int end = (int)((long long)n * (t + 1) / nthreads);
        workers.emplace_back(std::ref(plan_range), begin, end, std::ref(parts[t].buckets),
                             std::ref(parts[t]));
    }
    for (std::thread &w : workers)
        w.join();
    for (int t = 0; t < nthreads; t++) {
        if (parts[t].rc)
            return fail(parts[t].rc, parts[t].err);
        any = any || parts[t].any;
        for (size_t i = 0; i < parts[t].buckets.b.size(); i++) {
            const Bucket &src = parts[t].buckets.b[i];
            if (src.probs.empty())
                continue;
            Bucket &dst = g_buckets.b[i];
            dst.S = src.S;
            dst.NCL = src.NCL;
            dst.unit = src.unit;
            dst.hyper = src.hyper;
            dst.max_D = std::max(dst.max_D, src.max_D);
            dst.probs.insert(dst.probs.end(), src.probs.begin(), src.probs.end());
            dst.wide.insert(dst.wide.end(), src.wide.begin(), src.wide.end());
        }
    }
    return PQ_OK;
}

struct Epilogue {
    const double2 *d_U = nullptr; // gather mode / sampler: matrix on the device
    int ldu = 0;
    double *pmf = nullptr;        // sampler: host [nshots][ldu], row = LapProblem.tag
    const double *u = nullptr;    // sampler draw: host [nshots] uniform variates ...
    int32_t *index = nullptr;     // ... and the drawn output modes (the pmf stays on the device)
    bool perm_only = false;       // batched permanents: only the full product
    // single problem whose results stay on the device: d_dst[j] = result of compact
    // column map[j] (ncols entries, map on the host), nothing is downloaded
    double *d_dst = nullptr;
    const int32_t *map = nullptr;
    int ncols = 0;
};

// One bucket = one walk launch + one reduce launch (+ pmf epilogue).  Without
// an epilogue the compact results land in c->h_lap[2] ([n][NCP+1] complex).
int run_bucket(DeviceCtx *c, Bucket &bk, const Epilogue *epi)
{
    const int n = (int)bk.probs.size();
    const int NCP = bk.S * bk.NCL, ncp1 = NCP + 1;
    const int groups_per_block = kLapThreads / bk.S;
    const long long wave = (long long)c->num_sms * 8;
    // CTAs per problem: enough of them for ~32 waves over the whole launch.  Problems of
    // one photon step differ in size by factors of 2-4 (output collisions), and CTAs are
    // dispatched in problem order: fine-grained slices even the tail out (B200, 10^4 shots:
    // kernels 1.36 s at 4 waves, 1.28 s at 32; 1250 shots: 0.176 -> 0.162 s).  Tuning
    // knob: PQ_LAP_WAVES.
    static const long long kWaves = [] {
        const char *e = std::getenv("PQ_LAP_WAVES");
        const long long v = e ? std::atoll(e) : 32;
        return std::max<long long>(1, std::min<long long>(v, 64));
    }();
    const long long budget = std::max<long long>(1, (kWaves * wave) / n);
    int total_blocks = 0;
    for (LapProblem &q : bk.probs) {
        long long nb = (q.nseg + groups_per_block - 1) / groups_per_block;
        nb = std::max<long long>(1, std::min(nb, budget));
        q.first_block = total_blocks;
        q.nblocks = (int)nb;
        total_blocks += (int)nb;
    }
    const size_t out_bytes = (size_t)n * ncp1 * sizeof(double2);
    int rc;
    Lap lap;
    if ((rc = grow_dev(c, 0, sizeof(LapProblem) * (size_t)n)) ||
        (rc = grow_dev(c, 2, (size_t)total_blocks * ncp1 * 4 * sizeof(double))) ||
        (rc = grow_dev(c, 3, out_bytes)))
        return rc;
    cudaStream_t st = c->stream;
    // descriptors go through pinned staging: a pageable source makes the copy a
    // driver-staged, host-synchronous one whose cost depends on the state of the
    // host's pages (megabytes per photon step in the sampler)
    if ((rc = grow_host(c, 4, sizeof(LapProblem) * (size_t)n)))
        return rc;
    lap.to(g_sampler_detail[0]);
    std::memcpy(c->h_lap[4], bk.probs.data(), sizeof(LapProblem) * (size_t)n);
    lap.to(g_sampler_detail[1]);
    PQ_CUDA(cudaMemcpyAsync(c->d_lap[0], c->h_lap[4], sizeof(LapProblem) * (size_t)n,
                            cudaMemcpyHostToDevice, st));
    LapParams P;
    std::memset(&P, 0, sizeof(P));
    P.prob = reinterpret_cast<const LapProblem *>(c->d_lap[0]);
    if (epi && epi->d_U) {
        P.U = epi->d_U;
        P.ldu = epi->ldu;
    } else {
        if ((rc = grow_dev(c, 1, bk.a2.size() * sizeof(double))))
            return rc;
        PQ_CUDA(cudaMemcpyAsync(c->d_lap[1], bk.a2.data(), bk.a2.size() * sizeof(double),
                                cudaMemcpyHostToDevice, st));
        P.A2 = reinterpret_cast<const double2 *>(c->d_lap[1]);
    }
    if (!bk.wide.empty()) {
        const size_t wbytes = bk.wide.size() * sizeof(LapWide);
        if ((rc = grow_dev(c, 8, wbytes)))
            return rc;
        PQ_CUDA(cudaMemcpyAsync(c->d_lap[8], bk.wide.data(), wbytes, cudaMemcpyHostToDevice, st));
        P.wide = reinterpret_cast<const LapWide *>(c->d_lap[8]);
    }
    P.partials = reinterpret_cast<double *>(c->d_lap[2]);
    P.out = reinterpret_cast<double2 *>(c->d_lap[3]);
    P.nprob = n;
    P.perm_only = (epi && epi->perm_only) ? 1 : 0;
    P.max_D = bk.max_D;
    const size_t smem = lap_smem_bytes(bk.max_D, bk.S, bk.NCL, P.perm_only != 0);
    PQ_CUDA(cudaEventRecord(c->lap_ev0, st));
    // accumulation mode of the walk: full product only (batched permanents), the
    // leave-one-out sums, or both (a caller's zero-multiplicity column gets the full product)
    const int mode = P.perm_only ? 2 : (bk.need_full ? 1 : 0);
    cudaError_t e = bk.hyper
                        ? launch_perm_hyper(bk.NCL, P, total_blocks, smem, st)
                        : launch_laplace(bk.S, bk.NCL, bk.unit, mode, P, total_blocks, smem, st);
    if (e != cudaSuccess) {
        const std::string what = "launch laplace_walk_kernel<NCL=" + std::to_string(bk.NCL) +
                                 ", S=" + std::to_string(bk.S) + ", unit=" +
                                 std::to_string((int)bk.unit) + ", mode=" + std::to_string(mode) +
                                 "> grid=" + std::to_string(total_blocks) +
                                 " smem=" + std::to_string(smem);
        return fail_cuda(e, what.c_str());
    }
    e = launch_laplace_reduce(P, ncp1, st);
    if (e != cudaSuccess)
        return fail_cuda(e, "launch laplace_reduce_kernel");
    g_launches += 2;
    void *h_dst = nullptr;
    size_t bytes = 0;
    const void *d_src = nullptr;
    if (epi && epi->index) {
        bytes = (size_t)n * sizeof(int32_t);
        if ((rc = grow_dev(c, 5, (size_t)n * epi->ldu * sizeof(double))) ||
            (rc = grow_dev(c, 6, (size_t)n * sizeof(double))) || (rc = grow_dev(c, 7, bytes)) ||
            (rc = grow_host(c, 0, (size_t)n * sizeof(double))) || (rc = grow_host(c, 1, bytes)))
            return rc;
        double *hu = reinterpret_cast<double *>(c->h_lap[0]);
        for (int i = 0; i < n; i++)
            hu[i] = epi->u[bk.probs[i].tag];
        PQ_CUDA(cudaMemcpyAsync(c->d_lap[6], hu, (size_t)n * sizeof(double),
                                cudaMemcpyHostToDevice, st));
        e = launch_sampler_pmf(P, ncp1, epi->d_U, epi->ldu,
                               reinterpret_cast<double *>(c->d_lap[5]), st);
        if (e != cudaSuccess)
            return fail_cuda(e, "launch sampler_pmf_kernel");
        e = launch_sampler_draw(reinterpret_cast<const double *>(c->d_lap[5]), n, epi->ldu,
                                reinterpret_cast<const double *>(c->d_lap[6]),
                                reinterpret_cast<int *>(c->d_lap[7]), st);
        if (e != cudaSuccess)
            return fail_cuda(e, "launch sampler_draw_kernel");
        g_launches += 2;
        h_dst = c->h_lap[1];
        d_src = c->d_lap[7];
    } else if (epi && epi->pmf) {
        bytes = (size_t)n * epi->ldu * sizeof(double);
        if ((rc = grow_dev(c, 5, bytes)) || (rc = grow_host(c, 3, bytes)))
            return rc;
        e = launch_sampler_pmf(P, ncp1, epi->d_U, epi->ldu,
                               reinterpret_cast<double *>(c->d_lap[5]), st);
        if (e != cudaSuccess)
            return fail_cuda(e, "launch sampler_pmf_kernel");
        g_launches += 1;
        h_dst = c->h_lap[3];
        d_src = c->d_lap[5];
    } else if (epi && epi->d_dst) {
        // results stay on the device: scatter compact -> caller's columns there
        const size_t mbytes = (size_t)epi->ncols * sizeof(int32_t);
        if ((rc = grow_dev(c, 7, mbytes)) || (rc = grow_host(c, 1, mbytes)))
            return rc;
        std::memcpy(c->h_lap[1], epi->map, mbytes);
        PQ_CUDA(cudaMemcpyAsync(c->d_lap[7], c->h_lap[1], mbytes, cudaMemcpyHostToDevice, st));
        e = launch_laplace_scatter(reinterpret_cast<const double2 *>(c->d_lap[3]),
                                   reinterpret_cast<const int *>(c->d_lap[7]), epi->ncols,
                                   reinterpret_cast<double2 *>(epi->d_dst), st);
        if (e != cudaSuccess)
            return fail_cuda(e, "launch laplace_scatter_kernel");
        g_launches += 1;
    } else {
        bytes = out_bytes;
        if ((rc = grow_host(c, 2, bytes)))
            return rc;
        h_dst = c->h_lap[2];
        d_src = c->d_lap[3];
    }
    PQ_CUDA(cudaEventRecord(c->lap_ev1, st));
    if (bytes)
        PQ_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    lap.to(g_sampler_detail[2]);
    PQ_CUDA(cudaStreamSynchronize(st));
    lap.to(g_sampler_detail[3]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->lap_ev0, c->lap_ev1) == cudaSuccess)
        c->last_kernel_ms = (c->last_kernel_ms < 0 ? 0.0 : c->last_kernel_ms) + ms;
    if (epi && epi->index) {
        const int32_t *src = reinterpret_cast<const int32_t *>(c->h_lap[1]);
        for (int i = 0; i < n; i++)
            epi->index[bk.probs[i].tag] = src[i];
    } else if (epi && epi->pmf) {
        const double *src = reinterpret_cast<const double *>(c->h_lap[3]);
        for (int i = 0; i < n; i++)
            std::memcpy(epi->pmf + (size_t)bk.probs[i].tag * epi->ldu, src + (size_t)i * epi->ldu,
                        (size_t)epi->ldu * sizeof(double));
    }
    lap.to(g_sampler_detail[4]);
    return PQ_OK;
}

// part / nparts: walk only this rank's contiguous share of every problem's
// segments (results are then partial sums, already scaled by 2^-(sum_rows-1)).
int laplace_batch_locked(int nprob, const double *A, const int64_t *a_off, const int32_t *R,
                         const int32_t *C, const int32_t *rows, const int64_t *r_off,
                         const int32_t *cols, const int64_t *c_off, double *out,
                         const int64_t *o_off, int32_t *out_len, int part = 0, int nparts = 1,
                         int device = -1, double *d_out = nullptr)
{
    // d_out (device, 2 * C[0] doubles; nprob must be 1): leave the results there instead
    // of `out`; trivial problems are still answered on the host (out_len = 1)
    std::string err;
    LapShape sh;
    g_buckets.reset();
    bool any = false;
    for (int b = 0; b < nprob; b++) {
        if (R[b] < 0 || C[b] < 0)
            return fail(PQ_ERR_BAD_ARG, "negative shape in batch");
        if ((R[b] > 0 && !rows) || (C[b] > 0 && !cols) || (R[b] > 0 && C[b] > 0 && !A))
            return fail(PQ_ERR_BAD_ARG, "null matrix or multiplicity vector in batch");
        const int32_t *rw = rows + r_off[b], *cl = cols + c_off[b];
        const int rc = lap_shape(R[b], C[b], rw, cl, sh, err);
        if (rc)
            return fail(rc, err);
        if (sh.trivial) {
            out[2 * o_off[b]] = part == 0 ? 1.0 : 0.0; // the parts sum to the early-out [1]
            out[2 * o_off[b] + 1] = 0.0;
            out_len[b] = 1;
            continue;
        }
        out_len[b] = C[b];
        const LapVariant v = laplace_variant(sh.NC);
        Bucket &bk = g_buckets.get(v, sh.unit);
        const int NCP = v.S * v.NCL;
        if (lap_smem_bytes(sh.D, v.S, v.NCL) > kLapSmemLimit)
            return fail(PQ_ERR_TOO_LARGE, kLapTooWide);
        LapProblem q;
        LapWide w;
        lap_fill(sh, NCP, q, v.S == 32 ? &w : nullptr);
        if (v.S == 32)
            bk.wide.push_back(w);
        if (nparts > 1) {
            const long long lo = (long long)(((__int128)q.nseg * part) / nparts);
            const long long hi = (long long)(((__int128)q.nseg * (part + 1)) / nparts);
            q.seg_begin = lo;
            q.nseg = hi - lo;
        }
        q.tag = b;
        q.a_off = (long long)(bk.a2.size() / 2);
        // compacted, pre-doubled matrix: row 0 = pinned row, rows 1..D = 2 a_d
        const double *Ab = A + 2 * a_off[b];
        const size_t base = bk.a2.size();
        bk.a2.resize(base + (size_t)(sh.D + 1) * NCP * 2, 0.0);
        double *dst = bk.a2.data() + base;
        for (int j = 0; j < NCP; j++) {
            if (j < sh.NC) {
                const size_t src = ((size_t)sh.pinned * C[b] + sh.src_col[j]) * 2;
                dst[2 * j] = Ab[src];
                dst[2 * j + 1] = Ab[src + 1];
            } else {
                dst[2 * j] = 1.0; // padding column: s_j == 1 for every term
            }
        }
        for (int d = 0; d < sh.D; d++)
            for (int j = 0; j < sh.NC; j++) {
                const size_t src = ((size_t)sh.src_row[d] * C[b] + sh.src_col[j]) * 2;
                dst[((size_t)(d + 1) * NCP + j) * 2] = 2.0 * Ab[src];
                dst[((size_t)(d + 1) * NCP + j) * 2 + 1] = 2.0 * Ab[src + 1];
            }
        bk.max_D = std::max(bk.max_D, sh.D);
        if (sh.NC < C[b])
            bk.need_full = true;
        bk.probs.push_back(q);
        any = true;
    }
    if (!any)
        return PQ_OK;
    DeviceCtx *c = nullptr;
    int rc = ctx_get(device >= 0 ? device : g_devices[0], &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> dev_lock(c->mu);
    c->last_kernel_ms = -1.0;
    for (Bucket &bk : g_buckets.b) {
        if (bk.probs.empty())
            continue;
        if (d_out) {
            // one problem, results scattered on the device
            const int NCPd = bk.S * bk.NCL;
            std::vector<int32_t> map((size_t)C[0]);
            const int32_t *cm = cols + c_off[0];
            int k = 0;
            for (int j = 0; j < C[0]; j++)
                map[j] = cm[j] > 0 ? k++ : NCPd;
            Epilogue epi;
            epi.d_dst = d_out;
            epi.map = map.data();
            epi.ncols = C[0];
            rc = run_bucket(c, bk, &epi);
            if (rc)
                return rc;
            continue;
        }
        rc = run_bucket(c, bk, nullptr);
        if (rc)
            return rc;
        // scatter: compact column k -> original column; columns with multiplicity
        // 0 receive the full product (reference quirk)
        const int NCP = bk.S * bk.NCL, ncp1 = NCP + 1;
        const double *ho = reinterpret_cast<const double *>(c->h_lap[2]);
        for (size_t i = 0; i < bk.probs.size(); i++) {
            const int b = bk.probs[i].tag;
            const double *res = ho + i * ncp1 * 2;
            double *dst = out + 2 * o_off[b];
            const int32_t *cm = cols + c_off[b];
            int k = 0;
            for (int j = 0; j < C[b]; j++) {
                const int from = cm[j] > 0 ? k++ : NCP;
                dst[2 * j] = res[2 * from];
                dst[2 * j + 1] = res[2 * from + 1];
            }
        }
    }
    return PQ_OK;
}

// One photon step of the sampler for nshots shots sharing one interferometer.
// Host-side phases of the calling thread's last sampler step, in milliseconds:
// planning, waiting for the device lock, device phase (uploads, kernels,
// download, scatter), and the kernels alone (CUDA events).
thread_local double g_sampler_profile[4] = {0.0, 0.0, 0.0, 0.0};
// Gray-code terms and algorithmic flops (22 k per term of a k-column problem,
// SURVEY.md 8d) of all sampler steps since the last reset, over all threads
std::atomic<double> g_sampler_terms{0.0}, g_sampler_flops{0.0};
void add_work(double terms, double flops)
{
    double cur = g_sampler_terms.load();
    while (!g_sampler_terms.compare_exchange_weak(cur, cur + terms)) {
    }
    cur = g_sampler_flops.load();
    while (!g_sampler_flops.compare_exchange_weak(cur, cur + flops)) {
    }
}

// numpy's Generator.choice(d, p = row / sum(row)) for the uniform variate u, on
// the host (shots whose Laplace problem is the reference's early-out)
int draw_from_row(const double *row, int d, double u)
{
    double total = 0.0;
    for (int m = 0; m < d; m++)
        total += row[m];
    double last = 0.0;
    for (int m = 0; m < d; m++)
        last += row[m] / total;
    double c = 0.0;
    int idx = 0;
    for (int m = 0; m < d; m++) {
        c += row[m] / total;
        idx += (c / last <= u) ? 1 : 0;
    }
    return last == last ? idx : -1;
}

// pmf != nullptr: return the unnormalised pmf rows.  Otherwise draw on the
// device: index[s] = the output mode numpy's choice would pick for u[s].
// device < 0: the library's first device (pq_set_devices).
int sampler_step(const double *U, int d, int nshots, const int32_t *out_occ,
                 const int32_t *in_occ, double *pmf, const double *u = nullptr,
                 int32_t *index = nullptr, int device = -1)
{
    // Planning (zero filtering, shapes, problem descriptors) touches only the
    // caller's arguments and this thread's buckets: it runs OUTSIDE the library
    // lock, so that one thread can plan its photon step while another thread's
    // step occupies the GPU.
    {
        // ... but never without a device: there is no CPU path, not even for the
        // shots that need no kernel
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            std::lock_guard<std::mutex> lock(g_mu);
            DeviceCtx *c0 = nullptr;
            const int rc0 = ctx_get(g_devices[0], &c0);
            return rc0 ? rc0 : fail(PQ_ERR_NO_DEVICE, "no usable CUDA device");
        }
    }
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a)
            .count();
    };
    g_sampler_profile[0] = g_sampler_profile[1] = g_sampler_profile[2] = g_sampler_profile[3] = 0.0;
    // Shots are planned independently: with thousands of them the range is cut
    // over a few host threads, each filling its own buckets, merged in shot order.
    using Part = PlanPart;
    auto plan_range = [&](int begin, int end, Buckets &buckets, Part &part) {
        std::vector<double> trivial_row(pmf ? 0 : d);
        LapShape sh;
        for (int s = begin; s < end; s++) {
            const int32_t *oo = out_occ + (size_t)s * d, *io = in_occ + (size_t)s * d;
            // The shape is taken from the UNFILTERED occupations: dropping the zeros
            // first (_filter_zeros, sampling.py:711-720) keeps the order of the
            // remaining modes, so the same row is pinned and the same digits result.
            part.rc = lap_shape(d, d, oo, io, sh, part.err);
            if (part.rc)
                return;
            double *prow = pmf ? pmf + (size_t)s * d : trivial_row.data();
            if (sh.trivial) {
                // permanent_laplace returns [1] (src/permanent_laplace.cpp:52-57) and
                // _calculate_pmf then uses the first non-zero input mode only
                int j = -1;
                for (int m = 0; m < d && j < 0; m++)
                    if (io[m] > 0)
                        j = m;
                for (int m = 0; m < d; m++) {
                    double v = 0.0;
                    if (j >= 0) {
                        const double w = (double)io[j];
                        const double re = w * U[((size_t)m * d + j) * 2];
                        const double im = w * U[((size_t)m * d + j) * 2 + 1];
                        v = re * re + im * im;
                    }
                    prow[m] = v;
                }
                if (!pmf)
                    index[s] = draw_from_row(prow, d, u[s]);
                continue;
            }
            const LapVariant v = laplace_variant(sh.NC);
            Bucket &bk = buckets.get(v, sh.unit);
            if (lap_smem_bytes(sh.D, v.S, v.NCL) > kLapSmemLimit) {
                part.rc = PQ_ERR_TOO_LARGE;
                part.err = kLapTooWide;
                return;
            }
            LapProblem q;
            LapWide w;
            lap_fill(sh, v.S * v.NCL, q, v.S == 32 ? &w : nullptr);
            q.tag = s;
            q.rowmode[0] = (uint16_t)sh.pinned;
            for (int k = 0; k < sh.D; k++)
                q.rowmode[k + 1] = (uint16_t)sh.src_row[k];
            bk.max_D = std::max(bk.max_D, sh.D);
            bk.probs.push_back(q);
            if (v.S == 32)
                bk.wide.push_back(w);
            part.any = true;
        }
    };
    int rc = PQ_OK;
    bool any = false;
    if ((rc = plan_parallel(nshots, plan_range, any)))
        return rc;
    g_sampler_profile[0] = ms_since(t_begin);
    {
        double terms = 0.0, flops = 0.0;
        for (const Bucket &bk : g_buckets.b)
            for (const LapProblem &q : bk.probs) {
                const double t = (double)q.nseg * (double)q.W;
                terms += t;
                flops += t * 22.0 * (double)q.nc;
            }
        add_work(terms, flops);
    }
    if (!any) {
        // nothing to launch (e.g. the first photon of every shot): no kernel time
        std::lock_guard<std::mutex> lock(g_mu);
        DeviceCtx *c0 = nullptr;
        if ((rc = ctx_get(device >= 0 ? device : g_devices[0], &c0)))
            return rc;
        c0->last_kernel_ms = 0.0;
        return PQ_OK;
    }
    // The device phase holds only this device's lock: steps of other host threads
    // on other devices run concurrently (single-process multi-GPU sampling).
    DeviceCtx *c = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if ((rc = ctx_get(device >= 0 ? device : g_devices[0], &c)))
            return rc;
    }
    const auto t_planned = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> dev_lock(c->mu);
    g_sampler_profile[1] = ms_since(t_planned);
    const auto t_locked = std::chrono::steady_clock::now();
    for (double &x : g_sampler_detail)
        x = 0.0;
    Lap lap;
    const size_t ubytes = (size_t)d * d * sizeof(double2);
    if ((rc = grow_dev(c, 4, ubytes)))
        return rc;
    {
        // Size the scratch ONCE for this many shots: the photon steps of a run come
        // with growing column counts, and regrowing pinned / device buffers step
        // after step (free + allocate, megabytes each) is what the device phase was
        // losing milliseconds to.  Bounds: NCP + 1 <= min(d, 64) + 1 columns per
        // result, at most 64 waves (the cap of PQ_LAP_WAVES) + one CTA per shot.
        const size_t ns = (size_t)nshots;
        const size_t ncp1_max = (size_t)std::min(d, (int)kMaxCols) + 1;
        const size_t blocks_max = (size_t)64 * c->num_sms * 8 + ns;
        if ((rc = grow_dev(c, 0, sizeof(LapProblem) * ns)) ||
            (rc = grow_host(c, 4, sizeof(LapProblem) * ns)) ||
            (rc = grow_dev(c, 2, blocks_max * ncp1_max * 4 * sizeof(double))) ||
            (rc = grow_dev(c, 3, ns * ncp1_max * sizeof(double2))) ||
            (rc = grow_dev(c, 5, ns * (size_t)d * sizeof(double))))
            return rc;
        if (pmf) {
            if ((rc = grow_host(c, 3, ns * (size_t)d * sizeof(double))))
                return rc;
        } else if ((rc = grow_dev(c, 6, ns * sizeof(double))) ||
                   (rc = grow_dev(c, 7, ns * sizeof(int32_t))) ||
                   (rc = grow_host(c, 0, ns * sizeof(double))) ||
                   (rc = grow_host(c, 1, ns * sizeof(int32_t)))) {
            return rc;
        }
    }
    lap.to(g_sampler_detail[0]);
    // the interferometer changes only between runs: upload it when it differs from
    // the resident copy (compared on the host, 16 d^2 bytes)
    if (c->u_host.size() != ubytes / sizeof(double) ||
        std::memcmp(c->u_host.data(), U, ubytes) != 0) {
        c->u_host.assign(U, U + ubytes / sizeof(double));
        PQ_CUDA(cudaMemcpyAsync(c->d_lap[4], c->u_host.data(), ubytes, cudaMemcpyHostToDevice,
                                c->stream));
        PQ_CUDA(cudaStreamSynchronize(c->stream)); // u_host is pageable: settle it now
    }
    lap.to(g_sampler_detail[5]);
    Epilogue epi;
    epi.d_U = reinterpret_cast<const double2 *>(c->d_lap[4]);
    epi.ldu = d;
    epi.pmf = pmf;
    epi.u = u;
    epi.index = pmf ? nullptr : index;
    c->last_kernel_ms = -1.0;
    for (Bucket &bk : g_buckets.b) {
        if (bk.probs.empty())
            continue;
        rc = run_bucket(c, bk, &epi);
        if (rc)
            return rc;
    }
    g_sampler_profile[2] = ms_since(t_locked);
    g_sampler_profile[3] = c->last_kernel_ms > 0.0 ? c->last_kernel_ms : 0.0;
    return PQ_OK;
}

// Many permanents of minors (with multiplicities) of one matrix: the batched
// form of connector.permanent(interferometer, cols=input, rows=output)
// (piquasso/_simulators/passive/utils.py:131-138) for probability tables.
int perm_batch_locked(const double *U, int R, int C, int nprob, const int32_t *row_mult,
                      const int32_t *col_mult, double *out)
{
    // PQ_PERM_HYPER=0 keeps every problem on the term-by-term walk (A/B measurements)
    static const bool use_hyper = [] {
        const char *e = std::getenv("PQ_PERM_HYPER");
        return e ? std::atoi(e) != 0 : true;
    }();
    // segments a problem should yield so that the batch fills ~3 CTAs on each of 148 SMs
    const int min_segs_fill = (int)std::min<long long>(1 << 20, (3LL * 148 * kLapThreads) / nprob);
    // Large batches go through in chunks: 440 bytes of descriptor per problem, staged in
    // pinned host memory and uploaded, stay bounded (58 MB) however many problems come.
    constexpr int kChunk = 1 << 17;
    DeviceCtx *c = nullptr;
    std::unique_lock<std::mutex> dev_lock;
    Epilogue epi;
    int rc = PQ_OK;
    for (int base = 0; base < nprob; base += kChunk) {
        const int nchunk = std::min(kChunk, nprob - base);
        auto plan_range = [&](int begin, int end, Buckets &buckets, PlanPart &part) {
            LapShape sh;
            for (int b = base + begin; b < base + end; b++) {
                const int32_t *rw = row_mult + (size_t)b * R, *cl = col_mult + (size_t)b * C;
                part.rc = lap_shape(R, C, rw, cl, sh, part.err);
                if (part.rc)
                    return;
                long long sr = 0, sc = 0;
                for (int i = 0; i < R; i++)
                    sr += rw[i];
                for (int j = 0; j < C; j++)
                    sc += cl[j];
                if (sr != sc) { // src/permanent.cpp:97-104
                    part.rc = PQ_ERR_SUM_MISMATCH;
                    part.err = "Number of input and output states should be equal (problem " +
                               std::to_string(b) + ")";
                    return;
                }
                if (sh.trivial) { // src/permanent.cpp:106-108
                    out[2 * (size_t)b] = 1.0;
                    out[2 * (size_t)b + 1] = 0.0;
                    continue;
                }
                // full product only: one lane per segment up to kPermS1MaxCols columns,
                // and blocks of 2^3 terms where three binary rows exist
                lap_expand_columns(sh);
                const LapVariant v = perm_variant(sh.NC);
                const bool hyper = use_hyper && lap_make_hyper(sh);
                Bucket &bk = buckets.get(v, sh.unit, hyper);
                if (lap_smem_bytes(sh.D, v.S, v.NCL, true) > kLapSmemLimit) {
                    part.rc = PQ_ERR_TOO_LARGE;
                    part.err = kLapTooWide;
                    return;
                }
                LapProblem q;
                LapWide w;
                // small problems: short segments, so that a CTA's threads all get one; a
                // batch of a few problems: enough segments to fill the device
                lap_fill(sh, v.S * v.NCL, q, v.S == 32 ? &w : nullptr,
                         std::max(kLapThreads / v.S, min_segs_fill));
                q.tag = b;
                q.rowmode[0] = (uint16_t)sh.pinned;
                for (int k = 0; k < sh.D; k++)
                    q.rowmode[k + 1] = (uint16_t)sh.src_row[k];
                bk.max_D = std::max(bk.max_D, sh.D);
                bk.probs.push_back(q);
                if (v.S == 32)
                    bk.wide.push_back(w);
                part.any = true;
            }
        };
        bool any = false;
        {
            const int prc = plan_parallel(nchunk, plan_range, any);
            if (prc)
                return prc;
        }
        if (!any)
            continue;
        if (!c) {
            if ((rc = ctx_get(g_devices[0], &c)))
                return rc;
            dev_lock = std::unique_lock<std::mutex>(c->mu);
            // gather mode needs a square leading dimension only for addressing: ldu = C
            const size_t ubytes = (size_t)R * C * sizeof(double2);
            if ((rc = grow_dev(c, 4, ubytes)))
                return rc;
            c->u_host.clear(); // the sampler's resident interferometer is overwritten
            PQ_CUDA(cudaMemcpyAsync(c->d_lap[4], U, ubytes, cudaMemcpyHostToDevice, c->stream));
            epi.d_U = reinterpret_cast<const double2 *>(c->d_lap[4]);
            epi.ldu = C;
            epi.perm_only = true;
            c->last_kernel_ms = -1.0;
        }
        for (Bucket &bk : g_buckets.b) {
            if (bk.probs.empty())
                continue;
            rc = run_bucket(c, bk, &epi);
            if (rc)
                return rc;
            const int ncp1 = bk.S * bk.NCL + 1;
            const double *ho = reinterpret_cast<const double *>(c->h_lap[2]);
            for (size_t i = 0; i < bk.probs.size(); i++) {
                const int b = bk.probs[i].tag;
                out[2 * (size_t)b] = ho[(i * ncp1 + (ncp1 - 1)) * 2];
                out[2 * (size_t)b + 1] = ho[(i * ncp1 + (ncp1 - 1)) * 2 + 1];
            }
        }
    } // chunks
    return PQ_OK;
}

} // namespace

namespace pqperm {
int perm_wide_locked(const double *A, int R, int C, const int32_t *rows, const int32_t *cols,
                     double out[2])
{
    return perm_batch_locked(A, R, C, 1, rows, cols, out);
}
} // namespace pqperm

extern "C" int pq_perm_batch_plan(int R, int C, const int32_t *rows, const int32_t *cols,
                                  int nprob, pq_plan_info *info)
{
    if (!info || R < 0 || C < 0 || nprob < 1 || (R > 0 && !rows) || (C > 0 && !cols))
        return fail(PQ_ERR_BAD_ARG, "bad batch plan arguments");
    std::memset(info, 0, sizeof(*info));
    LapShape sh;
    std::string err;
    int rc = lap_shape(R, C, rows, cols, sh, err);
    if (rc)
        return fail(rc, err);
    long long sc = 0;
    for (int j = 0; j < C; j++)
        sc += cols[j];
    if (sh.sum_rows != sc)
        return fail(PQ_ERR_SUM_MISMATCH, "Number of input and output states should be equal");
    info->sum_rows = sh.sum_rows;
    info->idx_max = 1;
    info->seg_len = 1;
    info->nseg = 1;
    if (sh.trivial) {
        info->trivial = 1;
        return PQ_OK;
    }
    // the same steps as perm_batch_locked's planner
    static const bool use_hyper = [] {
        const char *e = std::getenv("PQ_PERM_HYPER");
        return e ? std::atoi(e) != 0 : true;
    }();
    lap_expand_columns(sh);
    const LapVariant v = perm_variant(sh.NC);
    const bool hyper = use_hyper && lap_make_hyper(sh);
    if (lap_smem_bytes(sh.D, v.S, v.NCL, true) > kLapSmemLimit)
        return fail(PQ_ERR_TOO_LARGE, kLapTooWide);
    const int min_segs_fill =
        (int)std::min<long long>(1 << 20, (3LL * 148 * kLapThreads) / nprob);
    LapProblem q;
    LapWide w;
    lap_fill(sh, v.S * v.NCL, q, v.S == 32 ? &w : nullptr,
             std::max(kLapThreads / v.S, min_segs_fill));
    info->seg_len = q.W;
    info->nseg = q.nseg;
    info->idx_max = (int64_t)q.W * q.nseg;
    info->active_rows = sh.D;
    info->active_cols = sh.NC;
    info->low_digits = q.q;
    info->kernel = hyper ? 4 : 3;
    info->cols_padded = v.S * v.NCL;
    info->flops_per_term = 2.0 * sh.NC + 6.0 * sh.M + 2.0;
    return PQ_OK;
}

extern "C" int pq_perm_batch_c128(const double *A, int R, int C, int nprob,
                                  const int32_t *row_mult, const int32_t *col_mult,
                                  double *out)
{
    if (R < 0 || C < 0 || R > 65535 || C > 65535 || nprob < 0 ||
        (nprob > 0 && (!out || (R > 0 && !row_mult) || (C > 0 && !col_mult))) ||
        (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "bad batch arguments");
    if (nprob == 0)
        return PQ_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    return perm_batch_locked(A, R, C, nprob, row_mult, col_mult, out);
}

extern "C" int pq_sampler_pmf_c128(const double *U, int d, int nshots, const int32_t *out_occ,
                                   const int32_t *in_occ, double *pmf)
{
    if (d < 1 || d > 65535 || nshots < 0 || !U || (nshots > 0 && (!out_occ || !in_occ || !pmf)))
        return fail(PQ_ERR_BAD_ARG, "bad sampler arguments");
    if (nshots == 0)
        return PQ_OK;
    return sampler_step(U, d, nshots, out_occ, in_occ, pmf);
}

extern "C" int pq_sampler_pmf_dev_c128(int device, const double *U, int d, int nshots,
                                       const int32_t *out_occ, const int32_t *in_occ,
                                       double *pmf)
{
    if (device < 0 || d < 1 || d > 65535 || nshots < 0 || !U ||
        (nshots > 0 && (!out_occ || !in_occ || !pmf)))
        return fail(PQ_ERR_BAD_ARG, "bad sampler arguments");
    if (nshots == 0)
        return PQ_OK;
    return sampler_step(U, d, nshots, out_occ, in_occ, pmf, nullptr, nullptr, device);
}

extern "C" int pq_sampler_draw_c128(const double *U, int d, int nshots, const int32_t *out_occ,
                                    const int32_t *in_occ, const double *u, int32_t *index)
{
    if (d < 1 || d > 65535 || nshots < 0 || !U ||
        (nshots > 0 && (!out_occ || !in_occ || !u || !index)))
        return fail(PQ_ERR_BAD_ARG, "bad sampler arguments");
    if (nshots == 0)
        return PQ_OK;
    return sampler_step(U, d, nshots, out_occ, in_occ, nullptr, u, index);
}

extern "C" void pq_last_sampler_detail(double out_ms[8])
{
    for (int i = 0; i < 8; i++)
        out_ms[i] = g_sampler_detail[i];
}

extern "C" void pq_sampler_work(double out[2])
{
    out[0] = g_sampler_terms.load();
    out[1] = g_sampler_flops.load();
}

extern "C" void pq_sampler_work_reset(void)
{
    g_sampler_terms.store(0.0);
    g_sampler_flops.store(0.0);
}

extern "C" void pq_last_sampler_profile(double out_ms[4])
{
    for (int i = 0; i < 4; i++)
        out_ms[i] = g_sampler_profile[i];
}

extern "C" int pq_sampler_draw_dev_c128(int device, const double *U, int d, int nshots,
                                        const int32_t *out_occ, const int32_t *in_occ,
                                        const double *u, int32_t *index)
{
    if (device < 0 || d < 1 || d > 65535 || nshots < 0 || !U ||
        (nshots > 0 && (!out_occ || !in_occ || !u || !index)))
        return fail(PQ_ERR_BAD_ARG, "bad sampler arguments");
    if (nshots == 0)
        return PQ_OK;
    return sampler_step(U, d, nshots, out_occ, in_occ, nullptr, u, index, device);
}

extern "C" int pq_perm_laplace_batch_c128(int nprob, const double *A, const int64_t *a_off,
                                          const int32_t *R, const int32_t *C,
                                          const int32_t *rows, const int64_t *r_off,
                                          const int32_t *cols, const int64_t *c_off,
                                          double *out, const int64_t *o_off,
                                          int32_t *out_len)
{
    if (nprob < 0 || (nprob > 0 && (!a_off || !R || !C || !r_off || !c_off || !out ||
                                    !o_off || !out_len)))
        return fail(PQ_ERR_BAD_ARG, "null pointer in batch call");
    if (nprob == 0)
        return PQ_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    return laplace_batch_locked(nprob, A, a_off, R, C, rows, r_off, cols, c_off, out, o_off,
                                out_len);
}

extern "C" int pq_perm_laplace_c128(const double *A, int R, int C, const int32_t *rows,
                                    const int32_t *cols, double *out, int *out_len)
{
    if (!out || !out_len || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    const int64_t zero = 0;
    const int32_t r32 = R, c32 = C;
    int32_t len = 0;
    const int rc = pq_perm_laplace_batch_c128(1, A, &zero, &r32, &c32, rows, &zero, cols, &zero,
                                              out, &zero, &len);
    if (rc)
        return rc;
    *out_len = len;
    return PQ_OK;
}

extern "C" int pq_perm_laplace_partial_c128(const double *A, int R, int C, const int32_t *rows,
                                            const int32_t *cols, int part, int nparts,
                                            double *out, int *out_len)
{
    if (!out || !out_len || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    if (R < 0 || C < 0 || nparts < 1 || part < 0 || part >= nparts)
        return fail(PQ_ERR_BAD_ARG, "bad shape or part index");
    const int64_t zero = 0;
    const int32_t r32 = R, c32 = C;
    int32_t len = 0;
    std::lock_guard<std::mutex> lock(g_mu);
    const int rc = laplace_batch_locked(1, A, &zero, &r32, &c32, rows, &zero, cols, &zero, out,
                                        &zero, &len, part, nparts);
    if (rc)
        return rc;
    *out_len = len;
    return PQ_OK;
}

extern "C" int pq_perm_laplace_partial_dev_c128(const double *A, int R, int C,
                                                const int32_t *rows, const int32_t *cols,
                                                int part, int nparts, int device,
                                                double *d_out, double *trivial, int *out_len)
{
    if (!d_out || !out_len || !trivial || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    if (R < 0 || C < 0 || nparts < 1 || part < 0 || part >= nparts || device < 0)
        return fail(PQ_ERR_BAD_ARG, "bad shape, part index or device");
    const int64_t zero = 0;
    const int32_t r32 = R, c32 = C;
    int32_t len = 0;
    // the early-out is answered on the host: recognised by the sentinel being overwritten
    std::vector<double> host(2 * (size_t)std::max(C, 1), std::nan(""));
    std::lock_guard<std::mutex> lock(g_mu);
    const int rc = laplace_batch_locked(1, A, &zero, &r32, &c32, rows, &zero, cols, &zero,
                                        host.data(), &zero, &len, part, nparts, device, d_out);
    if (rc)
        return rc;
    *out_len = len;
    trivial[0] = host[0]; // NaN unless the problem was the reference's early-out
    trivial[1] = host[1];
    return PQ_OK;
}

extern "C" int pq_perm_laplace_c64(const float *A, int R, int C, const int32_t *rows,
                                   const int32_t *cols, float *out, int *out_len)
{
    if (!out || !out_len || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    std::vector<double> Ad((size_t)(R > 0 ? R : 0) * (C > 0 ? C : 0) * 2);
    for (size_t i = 0; i < Ad.size(); i++)
        Ad[i] = (double)A[i];
    std::vector<double> o(2 * (size_t)(C > 0 ? C : 1));
    const int rc = pq_perm_laplace_c128(Ad.data(), R, C, rows, cols, o.data(), out_len);
    if (rc)
        return rc;
    for (int i = 0; i < 2 * *out_len; i++)
        out[i] = (float)o[i];
    return PQ_OK;
}
