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
// -DPQ_LAP_PART=1, 2 (mode 2 only): the one-lane-per-segment instantiations of the
// batched permanents for 9..20 and 21..32 columns, translation units of their own.
#ifndef PQ_LAP_PART
#define PQ_LAP_PART 0
#endif
#define PQ_CONCAT3_(a, b, c, d, e, f) a##b##c##d##e##f
#define PQ_CONCAT3(a, b, c, d, e, f) PQ_CONCAT3_(a, b, c, d, e, f)
#define PQ_LAP_LAUNCHER PQ_CONCAT3(launch_laplace_u, PQ_LAP_UNIT, _m, PQ_LAP_MODE, _p, PQ_LAP_PART)

namespace pqperm {

template <int NCL, int S>
static cudaError_t launch_one(const LapParams &P, int total_blocks, size_t smem,
                              cudaStream_t stream)
{
    auto kernel = laplace_walk_kernel<NCL, S, PQ_LAP_UNIT != 0, PQ_LAP_MODE>;
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

cudaError_t PQ_LAP_LAUNCHER(
    int S, int NCL, const LapParams &P, int total_blocks, size_t smem, cudaStream_t stream)
{
#define PQ_CASE(SS, NN)                                                                 \
    if (S == SS && NCL == NN)                                                           \
        return launch_one<NN, SS>(P, total_blocks, smem, stream);
#if PQ_LAP_PART == 1
    PQ_CASE(1, 9) PQ_CASE(1, 10) PQ_CASE(1, 11) PQ_CASE(1, 12) PQ_CASE(1, 13) PQ_CASE(1, 14)
    PQ_CASE(1, 15) PQ_CASE(1, 16) PQ_CASE(1, 17) PQ_CASE(1, 18) PQ_CASE(1, 19) PQ_CASE(1, 20)
#elif PQ_LAP_PART == 2
    PQ_CASE(1, 21) PQ_CASE(1, 22) PQ_CASE(1, 23) PQ_CASE(1, 24) PQ_CASE(1, 25) PQ_CASE(1, 26)
    PQ_CASE(1, 27) PQ_CASE(1, 28) PQ_CASE(1, 29) PQ_CASE(1, 30) PQ_CASE(1, 31) PQ_CASE(1, 32)
#else
    PQ_CASE(1, 1) PQ_CASE(1, 2) PQ_CASE(1, 3) PQ_CASE(1, 4) PQ_CASE(1, 5) PQ_CASE(1, 6)
    PQ_CASE(1, 7) PQ_CASE(1, 8)
    PQ_CASE(2, 5) PQ_CASE(2, 6) PQ_CASE(2, 7) PQ_CASE(2, 8) PQ_CASE(2, 9) PQ_CASE(2, 10) PQ_CASE(2, 11) PQ_CASE(2, 12)
    PQ_CASE(2, 13)
    PQ_CASE(4, 5) PQ_CASE(4, 6) PQ_CASE(4, 7) PQ_CASE(4, 8) PQ_CASE(4, 9) PQ_CASE(4, 10) PQ_CASE(4, 11) PQ_CASE(4, 12)
    PQ_CASE(4, 13) PQ_CASE(4, 14) PQ_CASE(4, 15) PQ_CASE(4, 16)
    PQ_CASE(32, 3) PQ_CASE(32, 4) PQ_CASE(32, 5) PQ_CASE(32, 6) PQ_CASE(32, 7) PQ_CASE(32, 8)
#endif
#undef PQ_CASE
    return cudaErrorInvalidValue;
}

#if PQ_LAP_UNIT == 1 && PQ_LAP_MODE == 0 && PQ_LAP_PART == 0
#define PQ_DECL(u, m, p)                                                                \
    cudaError_t launch_laplace_u##u##_m##m##_p##p(int, int, const LapParams &, int, size_t,  \
                                                  cudaStream_t);
PQ_DECL(0, 0, 0) PQ_DECL(0, 1, 0) PQ_DECL(0, 2, 0) PQ_DECL(1, 1, 0) PQ_DECL(1, 2, 0)
PQ_DECL(0, 2, 1) PQ_DECL(0, 2, 2) PQ_DECL(1, 2, 1) PQ_DECL(1, 2, 2)
#undef PQ_DECL

cudaError_t launch_laplace(int S, int NCL, bool unitcols, int mode, const LapParams &P,
                           int total_blocks, size_t smem, cudaStream_t stream)
{
    // batched permanents, one lane per segment, more than 8 columns (perm_variant)
    const int part = (mode == 2 && S == 1 && NCL > 8) ? (NCL <= 20 ? 1 : 2) : 0;
#define PQ_GO(u, m, p)                                                                  \
    if ((unitcols ? 1 : 0) == u && mode == m && part == p)                              \
        return launch_laplace_u##u##_m##m##_p##p(S, NCL, P, total_blocks, smem, stream);
    PQ_GO(0, 0, 0) PQ_GO(0, 1, 0) PQ_GO(0, 2, 0) PQ_GO(1, 0, 0) PQ_GO(1, 1, 0) PQ_GO(1, 2, 0)
    PQ_GO(0, 2, 1) PQ_GO(0, 2, 2) PQ_GO(1, 2, 1) PQ_GO(1, 2, 2)
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
