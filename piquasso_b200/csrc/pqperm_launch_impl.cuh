// pqperm_launch_impl.cuh -- launch helper shared by the kernel TUs.
#pragma once

#include <map>
#include <mutex>
#include <utility>

#include "pqperm_launch.h"
#include "pqperm_walk.cuh"

namespace pqperm {

struct OccKey {
    const void *fn;
    size_t smem;
    int dev;
    bool operator<(const OccKey &o) const
    {
        if (fn != o.fn)
            return fn < o.fn;
        if (smem != o.smem)
            return smem < o.smem;
        return dev < o.dev;
    }
};

// Resident blocks per SM for (kernel, smem) on the current device, cached; also
// raises the kernel's dynamic shared memory limit when needed -- only ever RAISES
// it (a later, smaller request above 48 KB must not lower the limit a cached,
// larger configuration relies on).
template <typename KernelT>
inline cudaError_t resident_blocks(KernelT kernel, int NT, size_t smem, int *nb)
{
    static std::mutex mu;
    static std::map<OccKey, int> cache;
    static std::map<std::pair<const void *, int>, size_t> raised; // (kernel, device) -> limit
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    std::lock_guard<std::mutex> lock(mu);
    OccKey key{(const void *)kernel, smem, dev};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *nb = it->second;
        return cudaSuccess;
    }
    // static (step tables of the n-ary walk, ~9.3 KB) + dynamic shared memory above
    // 48 KB needs the opt-in
    if (smem + 10 * 1024 > 48 * 1024) {
        size_t &limit = raised[std::make_pair((const void *)kernel, dev)];
        if (limit < smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
            if (e != cudaSuccess)
                return e;
            limit = smem;
        }
    }
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, NT, smem);
    if (e != cudaSuccess)
        return e;
    if (n < 1)
        n = 1;
    cache[key] = n;
    *nb = n;
    return cudaSuccess;
}

// One wave of resident CTAs, each thread striding over its segments.
template <typename KernelT, typename ParamT>
inline cudaError_t launch_walk(KernelT kernel, const ParamT &Q, const WalkParams &P, int NT,
                               size_t smem, int num_sms, int max_grid, cudaStream_t stream,
                               LaunchInfo *info)
{
    int nb = 1;
    cudaError_t e = resident_blocks(kernel, NT, smem, &nb);
    if (e != cudaSuccess)
        return e;
    const long long nseg = P.seg_end - P.seg_begin;
    long long want = (nseg + NT - 1) / NT;
    long long cap = (long long)num_sms * nb;
    if (cap > max_grid)
        cap = max_grid;
    if (want > cap)
        want = cap;
    if (want < 1)
        want = 1;
    kernel<<<(int)want, NT, smem, stream>>>(Q);
    if (info) {
        info->grid = (int)want;
        info->block = NT;
        info->smem = smem;
        info->blocks_per_sm = nb;
    }
    return cudaGetLastError();
}

} // namespace pqperm
