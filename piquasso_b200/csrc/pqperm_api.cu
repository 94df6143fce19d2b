// pqperm_api.cu -- the C ABI of libpqperm.so (include/pqperm.h): host-side
// preprocessing, device contexts, launches and result assembly.
//
// There is no CPU implementation of the term sum anywhere in this library: if
// no CUDA device is usable every compute entry returns PQ_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pqperm.h"
#include "pqperm_ctx.h"
#include "pqperm_launch.h"
#include "pqperm_plan.h"

namespace pqperm {

// parts of the binary kernel family (pqperm_kernels_binary.cu)
#define PQ_DECL_PART(k)                                                                 \
    cudaError_t launch_binary_part_##k(int, int, const WalkParams &, const double *,        \
                                       const double2 *, int, int, cudaStream_t, LaunchInfo *);
PQ_DECL_PART(0)
PQ_DECL_PART(1)
PQ_DECL_PART(2)
PQ_DECL_PART(3)
#undef PQ_DECL_PART

cudaError_t launch_binary(int nc, int B, const WalkParams &P, const double *h_A2,
                          const double2 *d_A2, int num_sms, int max_grid, cudaStream_t stream,
                          LaunchInfo *info)
{
    if (nc <= 20)
        return launch_binary_part_0(nc, B, P, h_A2, d_A2, num_sms, max_grid, stream, info);
    if (nc <= 30)
        return launch_binary_part_1(nc, B, P, h_A2, d_A2, num_sms, max_grid, stream, info);
    if (nc <= 40)
        return launch_binary_part_2(nc, B, P, h_A2, d_A2, num_sms, max_grid, stream, info);
    return launch_binary_part_3(nc, B, P, h_A2, d_A2, num_sms, max_grid, stream, info);
}

} // namespace pqperm

using namespace pqperm;

// ---------------------------------------------------------------------------
// error plumbing (declared in pqperm_ctx.h)
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

namespace pqperm {
int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

int fail_cuda(cudaError_t e, const char *what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    // a missing driver / device must read as "no device", not as a CUDA bug
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ||
        e == cudaErrorInitializationError)
        return PQ_ERR_NO_DEVICE;
    return PQ_ERR_CUDA;
}
} // namespace pqperm

extern "C" const char *pq_last_error(void) { return g_last_error.c_str(); }

// ---------------------------------------------------------------------------
// per-device context
// ---------------------------------------------------------------------------
namespace pqperm {

std::mutex g_mu;
std::vector<std::unique_ptr<DeviceCtx>> g_ctx;
std::vector<int> g_devices = {0};
std::atomic<int64_t> g_launches{0};
static int g_kernel_choice = 0;
static bool g_timing = true; // CUDA-event pairs around the walk kernels (pq_set_timing)
static int64_t g_seg_len_hint = 0;

int ctx_get(int device, DeviceCtx **out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PQ_ERR_NO_DEVICE,
                    std::string("no usable CUDA device (") +
                        (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                        "); libpqperm has no CPU fallback");
    }
    if (device < 0 || device >= ndev)
        return fail(PQ_ERR_BAD_ARG, "device index out of range");
    if ((int)g_ctx.size() < ndev)
        g_ctx.resize(ndev);
    if (!g_ctx[device])
        g_ctx[device].reset(new DeviceCtx());
    DeviceCtx *c = g_ctx[device].get();
    PQ_CUDA(cudaSetDevice(device));
    if (!c->ready) {
        c->device = device;
        PQ_CUDA(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
        PQ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        PQ_CUDA(cudaEventCreate(&c->ev0));
        PQ_CUDA(cudaEventCreate(&c->ev1));
        PQ_CUDA(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
        PQ_CUDA(cudaEventCreate(&c->lap_ev0));
        PQ_CUDA(cudaEventCreate(&c->lap_ev1));
        for (int i = 0; i < kTimingRing; i++) {
            PQ_CUDA(cudaEventCreate(&c->ring0[i]));
            PQ_CUDA(cudaEventCreate(&c->ring1[i]));
        }
        PQ_CUDA(cudaEventCreateWithFlags(&c->ev_busy, cudaEventDisableTiming));
        PQ_CUDA(cudaMalloc(&c->d_blob, kBlobBytes));
        PQ_CUDA(cudaMallocHost(&c->h_blob, kBlobBytes));
        PQ_CUDA(cudaMalloc(&c->d_partials, (size_t)kMaxGrid * 4 * sizeof(double)));
        PQ_CUDA(cudaHostAlloc(&c->h_out, 4 * sizeof(double), cudaHostAllocMapped));
        PQ_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->d_hout), c->h_out, 0));
        {
            // dispenser and finished-CTA count: zero once, the kernels re-arm them
            unsigned char *cnt = nullptr;
            PQ_CUDA(cudaMalloc(&cnt, 16));
            PQ_CUDA(cudaMemset(cnt, 0, 16));
            c->d_counter = reinterpret_cast<unsigned long long *>(cnt);
            c->d_done = reinterpret_cast<unsigned int *>(cnt + 8);
        }
        c->ready = true;
    }
    *out = c;
    return PQ_OK;
}

// Grow-only scratch of the batched Laplace / sampler path.  The FIRST allocation of a
// slot is at least its floor, sized for a 10^4-shot, 100-mode sampler run: regrowing
// (free + allocate; pinning megabytes of host memory) was measured to take up to
// 0.47 s now and then on the GPU box (tools/diag_sampler2.py: 2 of 8 fresh processes),
// which a run pays once, at its first small call, instead of in the middle of its first
// large one.
static const size_t kDevFloor[9] = {8u << 20, 1u << 20, 32u << 20, 16u << 20, 1u << 20,
                                    16u << 20, 1u << 20, 1u << 20, 1u << 20};
static const size_t kHostFloor[5] = {1u << 20, 1u << 20, 16u << 20, 16u << 20, 8u << 20};

int grow_dev(DeviceCtx *c, int slot, size_t bytes)
{
    if (c->d_lap_cap[slot] >= bytes)
        return PQ_OK;
    if (c->d_lap[slot])
        cudaFree(c->d_lap[slot]);
    if (slot == 4)
        c->u_host.clear(); // the resident interferometer goes with its buffer
    c->d_lap[slot] = nullptr;
    c->d_lap_cap[slot] = 0;
    const size_t cap = std::max(bytes + bytes / 2 + 4096, kDevFloor[slot]);
    PQ_CUDA(cudaMalloc(&c->d_lap[slot], cap));
    c->d_lap_cap[slot] = cap;
    return PQ_OK;
}

int grow_host(DeviceCtx *c, int slot, size_t bytes)
{
    if (c->h_lap_cap[slot] >= bytes)
        return PQ_OK;
    if (c->h_lap[slot])
        cudaFreeHost(c->h_lap[slot]);
    c->h_lap[slot] = nullptr;
    c->h_lap_cap[slot] = 0;
    const size_t cap = std::max(bytes + bytes / 2 + 4096, kHostFloor[slot]);
    PQ_CUDA(cudaMallocHost(&c->h_lap[slot], cap));
    c->h_lap_cap[slot] = cap;
    return PQ_OK;
}

// Layout of a plan's inputs inside the blob (device and pinned staging alike):
// the matrix, then the binomial tables of the n-ary flavours.
struct BlobLayout {
    size_t a2 = 0, binom = 0, total = 0;
    bool tables = false;
};

static BlobLayout blob_layout(const Plan &plan)
{
    BlobLayout L;
    L.tables = !(plan.binary && plan.unitcols); // the binary fast paths need no tables
    L.a2 = 0;
    L.binom = plan.A2.size() * sizeof(double);
    L.total = L.binom + (L.tables ? plan.binom.size() * sizeof(double) : 0);
    return L;
}

// the binary walk whose matrix rides in the kernel parameters needs no upload at all
static bool rides_in_params(const Plan &plan)
{
    return plan.kernel == 2 && plan.NC <= kBinMaxParamCols && plan.D + 1 == plan.NC;
}

static void fill_params(const Plan &plan, DeviceCtx *c, WalkParams &P)
{
    const BlobLayout L = blob_layout(plan);
    std::memset(&P, 0, sizeof(P));
    P.A2 = reinterpret_cast<const double2 *>(c->d_blob + L.a2);
    P.binom = reinterpret_cast<const double *>(c->d_blob + L.binom);
    P.partials = c->d_partials;
    P.segsums = nullptr;
    P.counter = c->d_counter;
    P.done = c->d_done;
    P.W = plan.W;
    P.D = plan.D;
    P.q = plan.q;
    for (int d = 0; d < plan.D; d++) {
        P.radix[d] = (uint8_t)(plan.mult[d] + 1);
        P.mult[d] = (uint8_t)plan.mult[d];
        P.binom_off[d] = (uint16_t)plan.binom_off[d];
    }
    for (int j = 0; j < plan.NCP; j++)
        P.colmult[j] = (uint8_t)plan.colmult[j];
}

// Before `stream` touches the device's shared scratch: wait (on the device) for
// the last launch if that ran on another stream.
static int scratch_acquire(DeviceCtx *c, cudaStream_t stream)
{
    if (c->busy_valid && c->busy_stream != stream)
        PQ_CUDA(cudaStreamWaitEvent(stream, c->ev_busy, 0));
    return PQ_OK;
}

static int scratch_release(DeviceCtx *c, cudaStream_t stream)
{
    if (stream != c->stream) { // the library's own stream is always synchronised by its caller
        PQ_CUDA(cudaEventRecord(c->ev_busy, stream));
        c->busy_stream = stream;
        c->busy_valid = true;
    } else {
        c->busy_valid = false;
    }
    return PQ_OK;
}

// Stage the plan's matrix and tables into c's blob on `stream`: ONE copy.
static int upload_plan(const Plan &plan, DeviceCtx *c, cudaStream_t stream)
{
    if (rides_in_params(plan))
        return PQ_OK; // the blob (and whichever job is resident in it) stays untouched
    c->resident_job = 0;
    int rc = scratch_acquire(c, stream);
    if (rc)
        return rc;
    // the pinned staging buffer may still feed the previous (caller-stream) upload
    if (c->up_pending) {
        PQ_CUDA(cudaEventSynchronize(c->ev_up));
        c->up_pending = false;
    }
    const BlobLayout L = blob_layout(plan);
    if (L.total > kBlobBytes)
        return fail(PQ_ERR_TOO_LARGE, "plan does not fit the staging blob");
    std::memcpy(c->h_blob + L.a2, plan.A2.data(), plan.A2.size() * sizeof(double));
    if (L.tables && !plan.binom.empty())
        std::memcpy(c->h_blob + L.binom, plan.binom.data(), plan.binom.size() * sizeof(double));
    PQ_CUDA(cudaMemcpyAsync(c->d_blob, c->h_blob, L.total, cudaMemcpyHostToDevice, stream));
    if (stream != c->stream) {
        PQ_CUDA(cudaEventRecord(c->ev_up, stream));
        c->up_pending = true;
    }
    return PQ_OK;
}

// Enqueue the walk of segments [seg_begin, seg_end) on `stream`, inputs already
// resident (upload_plan) or riding in the kernel parameters; the four-double sum
// lands in d_dst (device or mapped host memory), written by the kernel's last CTA.
// ONE launch, bracketed by an event pair of the timing ring.
static int launch_plan(const Plan &plan, DeviceCtx *c, int64_t seg_begin, int64_t seg_end,
                       double *d_dst, double *d_segsums, cudaStream_t stream)
{
    int rc = scratch_acquire(c, stream);
    if (rc)
        return rc;
    WalkParams P;
    fill_params(plan, c, P);
    P.seg_begin = seg_begin;
    P.seg_end = seg_end;
    P.segsums = d_segsums;
    P.out4 = d_dst;
    LaunchInfo info;
    cudaError_t e;
    const bool fast = plan.binary && plan.unitcols;
    const int slot = (int)(c->ring_next % kTimingRing);
    if (g_timing)
        PQ_CUDA(cudaEventRecord(c->ring0[slot], stream));
    if (plan.kernel == 2) {
        e = launch_binary(plan.NC, plan.B, P, rides_in_params(plan) ? plan.A2.data() : nullptr,
                          reinterpret_cast<const double2 *>(c->d_blob), c->num_sms, kMaxGrid,
                          stream, &info);
    } else {
        e = launch_generic(plan.NCP, fast, plan.unitcols, P, c->num_sms, kMaxGrid, stream, &info);
    }
    if (e != cudaSuccess)
        return fail_cuda(e, plan.kernel == 2 ? "launch perm_walk_binary" : "launch perm_walk_generic");
    if (g_timing) {
        PQ_CUDA(cudaEventRecord(c->ring1[slot], stream));
        c->ring_next++;
    }
    g_launches += 1;
    return scratch_release(c, stream);
}

static int enqueue_walk(const Plan &plan, DeviceCtx *c, int64_t seg_begin, int64_t seg_end,
                        double *d_dst, double *d_segsums, cudaStream_t stream)
{
    const int rc = upload_plan(plan, c, stream);
    if (rc)
        return rc;
    return launch_plan(plan, c, seg_begin, seg_end, d_dst, d_segsums, stream);
}

// duration of the most recent launch_plan on c (its stream must be synchronised)
static void note_last_kernel_ms(DeviceCtx *c)
{
    if (!g_timing) {
        c->last_kernel_ms = -1.0;
        return;
    }
    if (c->ring_next == 0)
        return;
    const int slot = (int)((c->ring_next - 1) % kTimingRing);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ring0[slot], c->ring1[slot]) == cudaSuccess)
        c->last_kernel_ms = ms;
    else
        cudaGetLastError();
}

static PlanOptions plan_options(int num_sms)
{
    PlanOptions o;
    o.kernel_choice = g_kernel_choice;
    o.seg_len_hint = g_seg_len_hint;
    o.num_sms = num_sms;
    return o;
}

static void split_range(int64_t nseg, int part, int nparts, int64_t *b, int64_t *e)
{
    *b = (int64_t)(((__int128)nseg * part) / nparts);
    *e = (int64_t)(((__int128)nseg * (part + 1)) / nparts);
}

} // namespace pqperm

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int pq_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pq_set_devices(const int *device_ids, int n)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (n < 1 || !device_ids)
        return fail(PQ_ERR_BAD_ARG, "need at least one device id");
    const int ndev = pq_device_count();
    std::vector<int> ids(device_ids, device_ids + n);
    for (int id : ids)
        if (id < 0 || id >= ndev)
            return fail(PQ_ERR_BAD_ARG, "device id out of range");
    g_devices = ids;
    return PQ_OK;
}

extern "C" int pq_set_kernel_choice(int choice)
{
    std::lock_guard<std::mutex> lock(g_mu);
    g_kernel_choice = choice;
    return PQ_OK;
}

extern "C" int pq_set_timing(int on)
{
    std::lock_guard<std::mutex> lock(g_mu);
    g_timing = on != 0;
    return PQ_OK;
}

extern "C" int pq_set_seg_len_hint(int64_t seg_len)
{
    std::lock_guard<std::mutex> lock(g_mu);
    g_seg_len_hint = seg_len;
    return PQ_OK;
}

extern "C" int64_t pq_launch_count(void) { return g_launches.load(); }

extern "C" double pq_last_kernel_ms(int device)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (device < 0 || device >= (int)g_ctx.size() || !g_ctx[device])
        return -1.0;
    std::lock_guard<std::mutex> dev_lock(g_ctx[device]->mu);
    return g_ctx[device]->last_kernel_ms;
}

// Error-free accumulation of (hi, lo) pairs on the host: the rank partials can
// be orders of magnitude larger than their sum, so a plain sum of the hi parts
// would throw away what the double-double kernels preserved.
static void dd_acc_host(double &hi, double &lo, double bhi, double blo)
{
    const double s = hi + bhi;
    const double bb = s - hi;
    const double e = (hi - (s - bb)) + (bhi - bb);
    hi = s;
    lo += e + blo;
}

extern "C" int pq_perm_combine(const double *quads, int n, double out4[4])
{
    if (!quads || !out4 || n < 0)
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    double rh = 0.0, rl = 0.0, ih = 0.0, il = 0.0;
    for (int i = 0; i < n; i++) { // fixed order => reproducible
        dd_acc_host(rh, rl, quads[4 * i + 0], quads[4 * i + 1]);
        dd_acc_host(ih, il, quads[4 * i + 2], quads[4 * i + 3]);
    }
    out4[0] = rh;
    out4[1] = rl;
    out4[2] = ih;
    out4[3] = il;
    return PQ_OK;
}

extern "C" int pq_perm_finish(const double partial[4], int sum_rows, double out[2])
{
    if (!partial || !out)
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    // src/permanent.cpp:259: permanent /= 2^(sum_rows - 1); exact power-of-two scaling
    out[0] = std::ldexp(partial[0] + partial[1], -(sum_rows - 1));
    out[1] = std::ldexp(partial[2] + partial[3], -(sum_rows - 1));
    return PQ_OK;
}

static void fill_info(const Plan &plan, pq_plan_info *info)
{
    std::memset(info, 0, sizeof(*info));
    info->idx_max = plan.idx_max;
    info->seg_len = plan.W;
    info->nseg = plan.nseg;
    info->active_rows = plan.D;
    info->active_cols = plan.NC_active;
    info->low_digits = plan.q;
    info->kernel = plan.kernel;
    info->cols_padded = plan.NCP;
    info->sum_rows = plan.sum_rows;
    info->trivial = plan.trivial;
    info->flops_per_term = 2.0 * plan.NC_active + 6.0 * plan.M + 2.0;
}

extern "C" int pq_perm_plan(int R, int C, const int32_t *rows, const int32_t *cols,
                            pq_plan_info *info)
{
    if (!info)
        return fail(PQ_ERR_BAD_ARG, "null info");
    Plan plan;
    std::string err;
    int num_sms = 148;
    PlanOptions o;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        o = plan_options(num_sms);
    }
    const int rc = make_plan(nullptr, R, C, rows, cols, o, plan, err);
    if (rc)
        return fail(rc, err);
    fill_info(plan, info);
    return PQ_OK;
}

extern "C" int pq_perm_gray_of_offset(int R, const int32_t *rows, int64_t offset,
                                      int32_t *gray)
{
    if (!gray || R <= 0 || !rows)
        return fail(PQ_ERR_BAD_ARG, "bad arguments");
    // multiplicity-only plan: one unit column per photon keeps the sum check happy
    int64_t n = 0;
    for (int i = 0; i < R; i++)
        n += rows[i];
    std::vector<int32_t> cols(1, (int32_t)n);
    Plan plan;
    std::string err;
    PlanOptions o;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        o = plan_options(148);
    }
    const int rc = make_plan(nullptr, R, 1, rows, cols.data(), o, plan, err);
    if (rc)
        return fail(rc, err);
    if (offset < 0 || offset >= plan.idx_max)
        return fail(PQ_ERR_BAD_ARG, "offset outside [0, idx_max)");
    plan_gray_of_offset(plan, offset, gray);
    return PQ_OK;
}

// Shared body of pq_perm_c128 / partial / segment sums.
// PQ_HOST_PROFILE=1: where the host side of pq_perm_c128 spends its time (ns per
// call, printed at exit): plan, context, enqueue (events + launch), wait, finish.
namespace {
struct HostProfile {
    bool on = std::getenv("PQ_HOST_PROFILE") != nullptr;
    double ns[5] = {0, 0, 0, 0, 0};
    long calls = 0;
    ~HostProfile()
    {
        if (on && calls)
            std::fprintf(stderr,
                         "pq host profile over %ld calls (us per call): plan %.2f, context %.2f, "
                         "enqueue %.2f, wait %.2f, finish %.2f\n",
                         calls, ns[0] / calls / 1e3, ns[1] / calls / 1e3, ns[2] / calls / 1e3,
                         ns[3] / calls / 1e3, ns[4] / calls / 1e3);
    }
};
HostProfile g_host_profile;
struct HostLap {
    std::chrono::steady_clock::time_point t;
    HostLap() { if (g_host_profile.on) t = std::chrono::steady_clock::now(); }
    void to(int slot)
    {
        if (!g_host_profile.on)
            return;
        const auto now = std::chrono::steady_clock::now();
        g_host_profile.ns[slot] += std::chrono::duration<double, std::nano>(now - t).count();
        t = now;
    }
};
} // namespace

static int perm_run(const double *A, int R, int C, const int32_t *rows, const int32_t *cols,
                    double out[2])
{
    if (!out || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    HostLap lap;
    {
        // more than PQ_MAX_COLS active columns (necessarily few rows): the
        // warp-per-segment batch walk, which holds up to 256 columns
        int nc = 0;
        for (int j = 0; j < C && cols; j++)
            nc += cols[j] > 0 ? 1 : 0;
        if (nc > PQ_MAX_COLS && R > 0 && rows)
            return perm_wide_locked(A, R, C, rows, cols, out);
    }
    // validate / early-outs first: they need no device, exactly like the reference
    Plan plan;
    std::string err;
    {
        PlanOptions o = plan_options(148);
        const int rc = make_plan(A, R, C, rows, cols, o, plan, err);
        if (rc)
            return fail(rc, err);
        if (plan.trivial) {
            out[0] = plan.triv[0];
            out[1] = plan.triv[1];
            return PQ_OK;
        }
    }
    lap.to(0);
    const std::vector<int> devices = g_devices;
    const int ndev = (int)devices.size();
    std::vector<DeviceCtx *> ctx(ndev, nullptr);
    for (int i = 0; i < ndev; i++) {
        const int rc = ctx_get(devices[i], &ctx[i]);
        if (rc)
            return rc;
    }
    if (ctx[0]->num_sms != 148) {
        // the partition is planned for ONE device's SM count whatever the number of
        // devices or ranks, so that every split of a problem cuts the same segments
        PlanOptions o = plan_options(ctx[0]->num_sms);
        const int rc = make_plan(A, R, C, rows, cols, o, plan, err);
        if (rc)
            return fail(rc, err);
    }
    // small problems are not worth a second device
    const int used = (ndev > 1 && plan.nseg >= (int64_t)ndev * 4096) ? std::min(ndev, 64) : 1;
    // device phase: under the per-device locks (lock order: g_mu, then devices)
    std::vector<std::unique_lock<std::mutex>> dev_locks;
    dev_locks.reserve(used);
    for (int i = 0; i < used; i++)
        dev_locks.emplace_back(ctx[i]->mu);
    lap.to(1);
    for (int i = 0; i < used; i++) {
        DeviceCtx *c = ctx[i];
        if (ndev > 1)
            PQ_CUDA(cudaSetDevice(c->device));
        int64_t b, e;
        split_range(plan.nseg, i, used, &b, &e);
        // the kernel's last CTA writes the sum straight into mapped host memory
        const int rc = enqueue_walk(plan, c, b, e, c->d_hout, nullptr, c->stream);
        if (rc)
            return rc;
    }
    lap.to(2);
    double quads[4 * 64];
    for (int i = 0; i < used; i++) {
        DeviceCtx *c = ctx[i];
        if (ndev > 1)
            PQ_CUDA(cudaSetDevice(c->device));
        PQ_CUDA(cudaStreamSynchronize(c->stream));
        lap.to(3);
        note_last_kernel_ms(c);
        for (int k = 0; k < 4; k++)
            quads[4 * i + k] = c->h_out[k];
    }
    double tot[4];
    pq_perm_combine(quads, used, tot); // fixed device order, error-free
    const int rc_fin = pq_perm_finish(tot, plan.sum_rows, out);
    lap.to(4);
    g_host_profile.calls++;
    return rc_fin;
}

extern "C" int pq_perm_c128(const double *A, int R, int C, const int32_t *rows,
                            const int32_t *cols, double out[2])
{
    return perm_run(A, R, C, rows, cols, out);
}

extern "C" int pq_perm_c64(const float *A, int R, int C, const int32_t *rows,
                           const int32_t *cols, float out[2])
{
    if (!out || (R > 0 && C > 0 && !A))
        return fail(PQ_ERR_BAD_ARG, "null pointer");
    std::vector<double> Ad((size_t)(R > 0 ? R : 0) * (C > 0 ? C : 0) * 2);
    for (size_t i = 0; i < Ad.size(); i++)
        Ad[i] = (double)A[i];
    double o[2];
    const int rc = perm_run(Ad.data(), R, C, rows, cols, o);
    if (rc)
        return rc;
    out[0] = (float)o[0];
    out[1] = (float)o[1];
    return PQ_OK;
}

extern "C" int pq_perm_partial_c128(const double *A, int R, int C, const int32_t *rows,
                                    const int32_t *cols, int part, int nparts, int device,
                                    void *stream, double *d_partial, int *status,
                                    double trivial[2])
{
    if (!d_partial || nparts < 1 || part < 0 || part >= nparts)
        return fail(PQ_ERR_BAD_ARG, "bad partition arguments");
    std::lock_guard<std::mutex> lock(g_mu);
    Plan plan;
    std::string err;
    {
        PlanOptions o = plan_options(148);
        const int rc = make_plan(A, R, C, rows, cols, o, plan, err);
        if (rc)
            return fail(rc, err);
        if (plan.trivial) {
            if (status)
                *status = 1;
            if (trivial) {
                trivial[0] = plan.triv[0];
                trivial[1] = plan.triv[1];
            }
            return PQ_OK;
        }
    }
    DeviceCtx *c = nullptr;
    int rc = ctx_get(device, &c);
    if (rc)
        return rc;
    if (c->num_sms != 148) {
        // planned for ONE device's SM count whatever nparts is: every rank, and every
        // rank count, cuts the term space into the same segments
        PlanOptions o = plan_options(c->num_sms);
        rc = make_plan(A, R, C, rows, cols, o, plan, err);
        if (rc)
            return fail(rc, err);
    }
    int64_t b, e;
    split_range(plan.nseg, part, nparts, &b, &e);
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::mutex> dev_lock(c->mu);
    if (b < e) {
        rc = enqueue_walk(plan, c, b, e, d_partial, nullptr, s);
        if (rc)
            return rc;
    } else {
        PQ_CUDA(cudaMemsetAsync(d_partial, 0, 4 * sizeof(double), s));
    }
    if (!stream) {
        PQ_CUDA(cudaStreamSynchronize(s));
        if (b < e)
            note_last_kernel_ms(c);
    }
    if (status)
        *status = 0;
    return PQ_OK;
}

extern "C" int pq_perm_segment_sums_c128(const double *A, int R, int C, const int32_t *rows,
                                         const int32_t *cols, int64_t seg_begin,
                                         int64_t nseg, double *out)
{
    if (!out || nseg < 1 || seg_begin < 0)
        return fail(PQ_ERR_BAD_ARG, "bad segment range");
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceCtx *c = nullptr;
    int rc = ctx_get(g_devices[0], &c);
    if (rc)
        return rc;
    Plan plan;
    std::string err;
    PlanOptions o = plan_options(c->num_sms);
    rc = make_plan(A, R, C, rows, cols, o, plan, err);
    if (rc)
        return fail(rc, err);
    if (plan.trivial)
        return fail(PQ_ERR_BAD_ARG, "trivial problem has no segments");
    if (seg_begin + nseg > plan.nseg)
        return fail(PQ_ERR_BAD_ARG, "segment range outside the plan");
    std::lock_guard<std::mutex> dev_lock(c->mu);
    double *d_seg = nullptr;
    PQ_CUDA(cudaMalloc(&d_seg, (size_t)nseg * 2 * sizeof(double)));
    PQ_CUDA(cudaMemsetAsync(d_seg, 0, (size_t)nseg * 2 * sizeof(double), c->stream));
    rc = enqueue_walk(plan, c, seg_begin, seg_begin + nseg, c->d_hout, d_seg, c->stream);
    if (rc == PQ_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_seg, (size_t)nseg * 2 * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess)
            rc = fail_cuda(e, "segment sums copy");
    }
    cudaFree(d_seg);
    return rc;
}

extern "C" double pq_fp64_peak_tflops(int device, int iters)
{
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceCtx *c = nullptr;
    if (ctx_get(device, &c))
        return -1.0;
    std::lock_guard<std::mutex> dev_lock(c->mu);
    if (iters < 1)
        iters = 1 << 16;
    double best = -1.0;
    for (int rep = 0; rep < 4; rep++) { // first pass is the warm-up
        double flops = 0.0;
        if (cudaEventRecord(c->ev0, c->stream) != cudaSuccess)
            return -1.0;
        if (launch_dfma_probe(c->num_sms, iters, c->d_partials, c->stream, &flops) != cudaSuccess)
            return -1.0;
        g_launches += 1;
        cudaEventRecord(c->ev1, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess)
            return -1.0;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (rep > 0 && ms > 0.f)
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    return best;
}

// ---------------------------------------------------------------------------
// resident jobs: inputs uploaded once, kernels launched many times
// ---------------------------------------------------------------------------
struct pq_perm_job {
    uint64_t id;
    int device;
    Plan plan;
    int64_t seg_begin, seg_end;
};
static std::atomic<uint64_t> g_job_ids{1};

extern "C" int pq_perm_job_create_c128(const double *A, int R, int C, const int32_t *rows,
                                       const int32_t *cols, int part, int nparts, int device,
                                       pq_perm_job **job, int *status, double trivial[2])
{
    if (!job || nparts < 1 || part < 0 || part >= nparts)
        return fail(PQ_ERR_BAD_ARG, "bad job arguments");
    *job = nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    std::unique_ptr<pq_perm_job> j(new pq_perm_job());
    std::string err;
    {
        PlanOptions o = plan_options(148);
        const int rc = make_plan(A, R, C, rows, cols, o, j->plan, err);
        if (rc)
            return fail(rc, err);
        if (j->plan.trivial) {
            if (status)
                *status = 1;
            if (trivial) {
                trivial[0] = j->plan.triv[0];
                trivial[1] = j->plan.triv[1];
            }
            return PQ_OK;
        }
    }
    DeviceCtx *c = nullptr;
    int rc = ctx_get(device, &c);
    if (rc)
        return rc;
    if (c->num_sms != 148) {
        PlanOptions o = plan_options(c->num_sms); // one device's SM count: same cut at every nparts
        rc = make_plan(A, R, C, rows, cols, o, j->plan, err);
        if (rc)
            return fail(rc, err);
    }
    std::lock_guard<std::mutex> dev_lock(c->mu);
    j->id = g_job_ids++;
    j->device = device;
    split_range(j->plan.nseg, part, nparts, &j->seg_begin, &j->seg_end);
    if (!rides_in_params(j->plan)) {
        rc = upload_plan(j->plan, c, c->stream);
        if (rc)
            return rc;
        PQ_CUDA(cudaStreamSynchronize(c->stream));
        c->resident_job = j->id;
    }
    if (status)
        *status = 0;
    *job = j.release();
    return PQ_OK;
}

extern "C" int pq_perm_job_launch(pq_perm_job *job, void *stream, double *d_partial)
{
    if (!job || !d_partial)
        return fail(PQ_ERR_BAD_ARG, "null job or output");
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceCtx *c = nullptr;
    int rc = ctx_get(job->device, &c);
    if (rc)
        return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    std::lock_guard<std::mutex> dev_lock(c->mu);
    if (!rides_in_params(job->plan) && c->resident_job != job->id) {
        // evicted by another call on this device
        rc = upload_plan(job->plan, c, s);
        if (rc)
            return rc;
        c->resident_job = job->id;
    }
    if (job->seg_begin < job->seg_end) {
        rc = launch_plan(job->plan, c, job->seg_begin, job->seg_end, d_partial, nullptr, s);
        if (rc)
            return rc;
    } else {
        PQ_CUDA(cudaMemsetAsync(d_partial, 0, 4 * sizeof(double), s));
    }
    if (!stream)
        PQ_CUDA(cudaStreamSynchronize(s));
    return PQ_OK;
}

extern "C" int pq_perm_job_info(const pq_perm_job *job, pq_plan_info *info)
{
    if (!job || !info)
        return fail(PQ_ERR_BAD_ARG, "null job or info");
    fill_info(job->plan, info);
    return PQ_OK;
}

extern "C" int64_t pq_perm_job_terms(const pq_perm_job *job)
{
    return job ? (job->seg_end - job->seg_begin) * job->plan.W : 0;
}

extern "C" int pq_perm_job_destroy(pq_perm_job *job)
{
    if (!job)
        return PQ_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    if (job->device >= 0 && job->device < (int)g_ctx.size() && g_ctx[job->device] &&
        g_ctx[job->device]->resident_job == job->id)
        g_ctx[job->device]->resident_job = 0;
    delete job;
    return PQ_OK;
}

// Durations (ms) of the most recent walk+reduce launches on `device`, newest
// first; the caller must have synchronised the stream they ran on.
extern "C" int pq_kernel_ms_history(int device, double *out, int max)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (device < 0 || device >= (int)g_ctx.size() || !g_ctx[device] || !out || max < 1)
        return 0;
    DeviceCtx *c = g_ctx[device].get();
    std::lock_guard<std::mutex> dev_lock(c->mu);
    int n = 0;
    for (uint64_t k = c->ring_next; k > 0 && n < max && n < kTimingRing; k--) {
        const int slot = (int)((k - 1) % kTimingRing);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ring0[slot], c->ring1[slot]) != cudaSuccess) {
            cudaGetLastError();
            break;
        }
        out[n++] = ms;
    }
    return n;
}

