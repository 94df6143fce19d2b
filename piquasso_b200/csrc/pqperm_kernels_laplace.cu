// pqperm_kernels_laplace.cu -- instantiations of the batched Laplace walk.
// Compiled once per flavour: -DPQ_LAP_UNIT=1 (all column multiplicities 1, what the
// sampler issues for single-photon inputs) or 0 (general), times -DPQ_LAP_MODE=0
// (leave-one-out sums only), 1 (plus the full product) or 2 (full product only).
#include <map>
#include <mutex>

#include "pqperm_launch.h"
#include "pqperm_laplace.cuh"

#if !defined(PQ_LAP_UNIT) || !defined(PQ_LAP_MODE)
#error "PQ_LAP_UNIT and PQ_LAP_MODE must be defined"
#endif
#define PQ_CONCAT3_(a, b, c, d) a##b##c##d
#define PQ_CONCAT3(a, b, c, d) PQ_CONCAT3_(a, b, c, d)
#define PQ_LAP_LAUNCHER PQ_CONCAT3(launch_laplace_u, PQ_LAP_UNIT, _m, PQ_LAP_MODE)

namespace pqperm {

template <int NCL, int S>
static cudaError_t launch_one(const LapParams &P, int total_blocks, size_t smem,
                              cudaStream_t stream)
{
    auto kernel = laplace_walk_kernel<NCL, S, PQ_LAP_UNIT != 0, PQ_LAP_MODE>;
    // static (step tables, ~4.1 KB) + dynamic shared memory above 48 KB needs the opt-in
    if (smem + 6 * 1024 > 48 * 1024) {
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

cudaError_t PQ_LAP_LAUNCHER(
    int S, int NCL, const LapParams &P, int total_blocks, size_t smem, cudaStream_t stream)
{
#define PQ_CASE(SS, NN)                                                                 \
    if (S == SS && NCL == NN)                                                           \
        return launch_one<NN, SS>(P, total_blocks, smem, stream);
    PQ_CASE(1, 1) PQ_CASE(1, 2) PQ_CASE(1, 3) PQ_CASE(1, 4) PQ_CASE(1, 5) PQ_CASE(1, 6)
    PQ_CASE(1, 7) PQ_CASE(1, 8)
    PQ_CASE(2, 5) PQ_CASE(2, 6) PQ_CASE(2, 7) PQ_CASE(2, 8) PQ_CASE(2, 9) PQ_CASE(2, 10) PQ_CASE(2, 11) PQ_CASE(2, 12)
    PQ_CASE(2, 13)
    PQ_CASE(4, 5) PQ_CASE(4, 6) PQ_CASE(4, 7) PQ_CASE(4, 8) PQ_CASE(4, 9) PQ_CASE(4, 10) PQ_CASE(4, 11) PQ_CASE(4, 12)
    PQ_CASE(4, 13) PQ_CASE(4, 14) PQ_CASE(4, 15) PQ_CASE(4, 16)
    PQ_CASE(32, 3) PQ_CASE(32, 4) PQ_CASE(32, 5) PQ_CASE(32, 6) PQ_CASE(32, 7) PQ_CASE(32, 8)
#undef PQ_CASE
    return cudaErrorInvalidValue;
}

#if PQ_LAP_UNIT == 1 && PQ_LAP_MODE == 0
#define PQ_DECL(u, m)                                                                   \
    cudaError_t launch_laplace_u##u##_m##m(int, int, const LapParams &, int, size_t, cudaStream_t);
PQ_DECL(0, 0) PQ_DECL(0, 1) PQ_DECL(0, 2) PQ_DECL(1, 1) PQ_DECL(1, 2)
#undef PQ_DECL

cudaError_t launch_laplace(int S, int NCL, bool unitcols, int mode, const LapParams &P,
                           int total_blocks, size_t smem, cudaStream_t stream)
{
#define PQ_GO(u, m)                                                                     \
    if ((unitcols ? 1 : 0) == u && mode == m)                                           \
        return launch_laplace_u##u##_m##m(S, NCL, P, total_blocks, smem, stream);
    PQ_GO(0, 0) PQ_GO(0, 1) PQ_GO(0, 2) PQ_GO(1, 0) PQ_GO(1, 1) PQ_GO(1, 2)
#undef PQ_GO
    return cudaErrorInvalidValue;
}

cudaError_t launch_laplace_reduce(const LapParams &P, int ncp1, cudaStream_t stream)
{
    laplace_reduce_kernel<<<(P.nprob + 3) / 4, 128, 0, stream>>>(P, ncp1);
    return cudaGetLastError();
}

cudaError_t launch_laplace_scatter(const double2 *res, const int *map, int ncols, double2 *out,
                                   cudaStream_t stream)
{
    laplace_scatter_kernel<<<(ncols + 127) / 128, 128, 0, stream>>>(res, map, ncols, out);
    return cudaGetLastError();
}

cudaError_t launch_sampler_pmf(const LapParams &P, int ncp1, const double2 *U, int d,
                               double *pmf, cudaStream_t stream)
{
    sampler_pmf_kernel<<<P.nprob, 128, 0, stream>>>(P, ncp1, U, d, pmf);
    return cudaGetLastError();
}

cudaError_t launch_sampler_draw(const double *pmf, int n, int d, const double *u, int *index,
                                cudaStream_t stream)
{
    sampler_draw_kernel<<<(n + 127) / 128, 128, 0, stream>>>(pmf, n, d, u, index);
    return cudaGetLastError();
}
#endif

} // namespace pqperm
