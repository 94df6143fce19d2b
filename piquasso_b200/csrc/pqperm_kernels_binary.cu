// pqperm_kernels_binary.cu -- instantiations of the binary constant-bank walk
// for column counts PQ_BIN_LO..PQ_BIN_HI.  Compiled once per range
// (-DPQ_BIN_PART=k -DPQ_BIN_LO=a -DPQ_BIN_HI=b) so the ranges build in
// parallel; each part owns its own __constant__ copy of the matrix (separate
// cubins, no relocatable device code).
#include "pqperm_device.cuh"

#ifndef PQ_BIN_PART
#error "PQ_BIN_PART / PQ_BIN_LO / PQ_BIN_HI must be defined"
#endif

namespace pqperm {
// (D+1) x NC doubled matrix of the permanent being computed; rows 1..B are
// addressed with compile-time offsets inside the unrolled inner block, the
// rest with a warp-uniform run-time row index (LDCU, uniform datapath).
__constant__ double2 c_matrix[(kBinMaxCols + 1) * kBinMaxCols];
} // namespace pqperm
#define PQ_BINARY_CONST_MATRIX ::pqperm::c_matrix

#include "pqperm_launch_impl.cuh"

#define PQ_CONCAT2(a, b) a##b
#define PQ_CONCAT(a, b) PQ_CONCAT2(a, b)

namespace pqperm {

template <int NC>
static cudaError_t launch_binary_nc(int B, int chains, const WalkParams &P, int num_sms, int max_grid,
                                    cudaStream_t stream, LaunchInfo *info)
{
    // 64-thread CTAs: at ~4*NC+40 registers per thread the register file holds
    // only a few warps per SM, and small CTAs waste the fewest of them.
    constexpr int NT = 64;
#define PQ_VARIANT(BB, CC)                                                              \
    if (B == BB && chains == CC)                                                        \
        return launch_walk(perm_walk_binary<NC, BB, CC, NT>, P, P, NT, 0, num_sms,      \
                           max_grid, stream, info);
    PQ_VARIANT(1, 2)
    PQ_VARIANT(2, 1)
    PQ_VARIANT(2, 2)
    PQ_VARIANT(3, 1)
    PQ_VARIANT(3, 2)
#undef PQ_VARIANT
    return cudaErrorInvalidValue;
}

template <int NC, int HI>
static cudaError_t dispatch_binary(int nc, int B, int chains, const WalkParams &P, int num_sms,
                                   int max_grid, cudaStream_t stream, LaunchInfo *info)
{
    if (nc == NC)
        return launch_binary_nc<NC>(B, chains, P, num_sms, max_grid, stream, info);
    if constexpr (NC < HI)
        return dispatch_binary<NC + 1, HI>(nc, B, chains, P, num_sms, max_grid, stream, info);
    else
        return cudaErrorInvalidValue;
}

cudaError_t PQ_CONCAT(launch_binary_part_, PQ_BIN_PART)(int nc, int B, int chains,
                                                        const WalkParams &P,
                                                        const double2 *A2_src,
                                                        cudaMemcpyKind kind, int num_sms,
                                                        int max_grid, cudaStream_t stream,
                                                        LaunchInfo *info)
{
    if (nc < PQ_BIN_LO || nc > PQ_BIN_HI)
        return cudaErrorInvalidValue;
    cudaError_t e = cudaMemcpyToSymbolAsync(c_matrix, A2_src,
                                            (size_t)(P.D + 1) * nc * sizeof(double2), 0, kind,
                                            stream);
    if (e != cudaSuccess)
        return e;
    return dispatch_binary<PQ_BIN_LO, PQ_BIN_HI>(nc, B, chains, P, num_sms, max_grid, stream, info);
}

} // namespace pqperm
