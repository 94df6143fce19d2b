// pqperm_kernels_permhyper.cu -- instantiations of the hypercube flavour of the
// batched-permanent walk (pqperm_permhyper.cuh) for 3..kPermS1MaxCols columns.
#include <map>
#include <mutex>

#include "pqperm_launch.h"
#include "pqperm_permhyper.cuh"

namespace pqperm {

template <int NC>
static cudaError_t launch_hyper_nc(const LapParams &P, int total_blocks, size_t smem,
                                   cudaStream_t stream)
{
    auto kernel = perm_hyper_kernel<NC>;
    // static (step tables) + dynamic shared memory above 48 KB needs the opt-in
    if (smem + kLapStaticSmem + 1024 > 48 * 1024) {
        static std::mutex mu;
        static std::map<int, size_t> raised; // device -> largest limit set
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(mu);
        if (raised[dev] < smem) {
            cudaError_t e = cudaFuncSetAttribute(
                kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return e;
            raised[dev] = smem;
        }
    }
    kernel<<<total_blocks, kLapThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

template <int NC>
static cudaError_t dispatch_hyper(int nc, const LapParams &P, int total_blocks, size_t smem,
                                  cudaStream_t stream)
{
    if (nc == NC)
        return launch_hyper_nc<NC>(P, total_blocks, smem, stream);
    if constexpr (NC < kPermS1MaxCols)
        return dispatch_hyper<NC + 1>(nc, P, total_blocks, smem, stream);
    else
        return cudaErrorInvalidValue;
}

cudaError_t launch_perm_hyper(int nc, const LapParams &P, int total_blocks, size_t smem,
                              cudaStream_t stream)
{
    return dispatch_hyper<kHyperMinCols>(nc, P, total_blocks, smem, stream);
}

} // namespace pqperm
