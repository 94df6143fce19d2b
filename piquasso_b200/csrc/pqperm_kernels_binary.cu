// pqperm_kernels_binary.cu -- instantiations of the binary hypercube walk
// (kernel 2) for column counts PQ_BIN_LO..PQ_BIN_HI.  Compiled once per range
// (-DPQ_BIN_PART=k -DPQ_BIN_LO=a -DPQ_BIN_HI=b) so the ranges build in
// parallel; each part owns its own __constant__ copies (separate cubins, no
// relocatable device code).
#include "pqperm_device.cuh"

#ifndef PQ_BIN_PART
#error "PQ_BIN_PART / PQ_BIN_LO / PQ_BIN_HI must be defined"
#endif

namespace pqperm {
// (D+1) x NC doubled matrix of the permanent being computed: rows 1..B with
// compile-time offsets inside the unrolled column loop, the others with a
// warp-uniform run-time row index (block-level Gray moves, seeding); both
// through the uniform datapath (LDCU).
__constant__ double2 c_matrix[(kBinMaxCols + 1) * kBinMaxCols];
} // namespace pqperm
#define PQ_BINARY_CONST_MATRIX ::pqperm::c_matrix

#include <cstring>

#include "pqperm_launch_impl.cuh"

#define PQ_CONCAT2(a, b) a##b
#define PQ_CONCAT(a, b) PQ_CONCAT2(a, b)

namespace pqperm {

template <int NC>
static cudaError_t launch_binary_nc(int B, const WalkParams &P, const double *h_A2,
                                    int num_sms, int max_grid, cudaStream_t stream,
                                    LaunchInfo *info)
{
    // 64-thread CTAs: at ~4*NC + 4*2^B registers per thread the register file
    // holds only a few warps per SM, and small CTAs waste the fewest of them.
    constexpr int NT = 64;
    if constexpr (NC <= kBinMaxParamCols) {
        if (h_A2) {
            // the matrix rides in the kernel's parameter block (constant bank 0)
            static_assert(sizeof(double2) == 2 * sizeof(double), "layout");
            WalkParamsM<NC> Q;
            Q.P = P;
            std::memcpy(Q.m, h_A2, sizeof(Q.m));
#define PQ_VARIANT(BB)                                                                  \
    if (B == BB)                                                                        \
        return launch_walk(perm_walk_binary_pm<NC, BB, NT>, Q, P, NT, 0, num_sms, max_grid, \
                           stream, info);
            PQ_VARIANT(2)
            PQ_VARIANT(3)
            if constexpr (NC <= 32) {
                PQ_VARIANT(4)
            }
#undef PQ_VARIANT
            return cudaErrorInvalidValue;
        }
    }
#define PQ_VARIANT(BB)                                                                  \
    if (B == BB)                                                                        \
        return launch_walk(perm_walk_binary<NC, BB, NT>, P, P, NT, 0, num_sms, max_grid,    \
                           stream, info);
    PQ_VARIANT(2)
    PQ_VARIANT(3)
    if constexpr (NC <= 32) {
        PQ_VARIANT(4)
    }
#undef PQ_VARIANT
    return cudaErrorInvalidValue;
}

template <int NC, int HI>
static cudaError_t dispatch_binary(int nc, int B, const WalkParams &P, const double *h_A2,
                                   int num_sms, int max_grid, cudaStream_t stream,
                                   LaunchInfo *info)
{
    if (nc == NC)
        return launch_binary_nc<NC>(B, P, h_A2, num_sms, max_grid, stream, info);
    if constexpr (NC < HI)
        return dispatch_binary<NC + 1, HI>(nc, B, P, h_A2, num_sms, max_grid, stream, info);
    else
        return cudaErrorInvalidValue;
}

// `h_A2` (host, (P.D+1) x nc double2) feeds the parameter-block kernel when the
// matrix fits; otherwise `d_A2` (device, same layout) is copied into the
// __constant__ matrix on `stream` first.
cudaError_t PQ_CONCAT(launch_binary_part_, PQ_BIN_PART)(int nc, int B, const WalkParams &P,
                                                        const double *h_A2,
                                                        const double2 *d_A2, int num_sms,
                                                        int max_grid, cudaStream_t stream,
                                                        LaunchInfo *info)
{
    if (nc < PQ_BIN_LO || nc > PQ_BIN_HI)
        return cudaErrorInvalidValue;
    if (nc > kBinMaxParamCols || P.D + 1 != nc)
        h_A2 = nullptr;
    if (!h_A2) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_matrix, d_A2,
                                                (size_t)(P.D + 1) * nc * sizeof(double2), 0,
                                                cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess)
            return e;
    }
    return dispatch_binary<PQ_BIN_LO, PQ_BIN_HI>(nc, B, P, h_A2, num_sms, max_grid, stream,
                                                 info);
}

} // namespace pqperm
