// pqperm_ctx.h -- per-device context and error plumbing shared by the host
// translation units of libpqperm (pqperm_api.cu, pqperm_api_laplace.cu).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pqperm.h"
#include "pqperm_limits.h"

namespace pqperm {

constexpr int kMaxGrid = 148 * 32 * 2;
constexpr int kTimingRing = 64;
constexpr size_t kA2Doubles = (size_t)(kMaxDigits + 1) * kMaxCols * 2;

// Staging blob of one permanent: [A2 | binomial tables], one H2D copy.
constexpr size_t kBlobBytes =
    kA2Doubles * sizeof(double) + (size_t)kMaxDigits * 256 * sizeof(double);

struct DeviceCtx {
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_up = nullptr;   // the last upload has left h_blob
    bool up_pending = false;
    // The walk kernels of one device share the scratch below (blob, partials,
    // counters, the __constant__ matrix).  A launch on a caller's stream may still
    // be running when the next call arrives on another stream: ev_busy is recorded
    // behind every launch and waited for (on the device, cudaStreamWaitEvent) by
    // the next upload / launch that uses a different stream.
    cudaEvent_t ev_busy = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy_valid = false;
    unsigned char *d_blob = nullptr;   // kBlobBytes: matrix and tables of the resident plan
    unsigned char *h_blob = nullptr;   // pinned staging of the same
    double *d_partials = nullptr;  // kMaxGrid x 4
    double *h_out = nullptr;       // 4 doubles, pinned + mapped: the kernels write the result here
    double *d_hout = nullptr;      // device alias of h_out
    unsigned long long *d_counter = nullptr;  // segment dispenser of the walk kernels ...
    unsigned int *d_done = nullptr;           // ... and their finished-CTA count (self-resetting)
    double last_kernel_ms = -1.0;
    bool ready = false;
    uint64_t resident_job = 0;     // id of the pq_perm_job whose inputs sit in the blob
    cudaEvent_t ring0[kTimingRing], ring1[kTimingRing];  // event pairs around the kernels
    uint64_t ring_next = 0;
    // ONE lock story: `mu` guards everything of this context that a device phase
    // touches -- `stream`, the scratch above and below, the timers.  Every entry
    // point takes it for its device phase.  g_mu (library-wide) only guards the
    // global tables and settings; the permanent entries additionally keep it for the
    // whole call (one permanent at a time), the sampler steps release it before their
    // device phase so that host threads driving DIFFERENT devices overlap.  Lock
    // order: g_mu, then device locks in the order of the device list.
    std::mutex mu;
    cudaEvent_t lap_ev0 = nullptr, lap_ev1 = nullptr;
    // growable scratch of the batched Laplace path
    // prob, A2, partials, out, U, pmf, uniform draws, drawn indices, wide column tables
    void *d_lap[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr};
    size_t d_lap_cap[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> u_host;  // host copy of the matrix resident in d_lap[4] (sampler)
    // (pinned) uniform draws, drawn indices, out, pmf, problem descriptors
    void *h_lap[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t h_lap_cap[5] = {0, 0, 0, 0, 0};
};

extern std::mutex g_mu;                 // one caller at a time (GIL-held callers anyway)
extern std::vector<std::unique_ptr<DeviceCtx>> g_ctx;
extern std::vector<int> g_devices;
extern std::atomic<int64_t> g_launches;

// Sets the calling thread's pq_last_error() message and returns `code`.
int fail(int code, const std::string &msg);
int fail_cuda(cudaError_t e, const char *what);

#define PQ_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return ::pqperm::fail_cuda(e__, #call);                                     \
    } while (0)

// Lazily creates the context of `device` (stream, events, fixed buffers) and
// makes it current.
int ctx_get(int device, DeviceCtx **out);

// permanent() of a problem with more than kMaxCols active columns, through the
// lane-split batch walk (pqperm_api_laplace.cu); g_mu held by the caller.
int perm_wide_locked(const double *A, int R, int C, const int32_t *rows, const int32_t *cols,
                     double out[2]);

// Grow-only device / pinned-host scratch of the batched Laplace path.
int grow_dev(DeviceCtx *c, int slot, size_t bytes);
int grow_host(DeviceCtx *c, int slot, size_t bytes);

} // namespace pqperm
