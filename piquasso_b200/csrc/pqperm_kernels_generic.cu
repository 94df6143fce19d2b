// pqperm_kernels_generic.cu -- instantiations of the generic n-ary walk and the
// DFMA throughput probe.
#include "pqperm_launch_impl.cuh"

namespace pqperm {

template <int NC>
static cudaError_t launch_generic_nc(bool binary, bool unitcols, const WalkParams &P,
                                     int num_sms, int max_grid, cudaStream_t stream,
                                     LaunchInfo *info)
{
    constexpr int NT = NC <= 16 ? 128 : 64;
    // matrix, the two low-digit vectors, one common-seed buffer per warp
    const size_t smem = (size_t)(P.D + 1 + 2 + NT / 32) * NC * sizeof(double2);
    if (binary && unitcols)
        return launch_walk(perm_walk_generic<NC, true, true, NT>, P, P, NT, smem, num_sms,
                           max_grid, stream, info);
    if (unitcols) // n-ary rows, unit (or expanded) columns: unrolled product chains
        return launch_walk(perm_walk_generic<NC, false, true, NT>, P, P, NT, smem, num_sms,
                           max_grid, stream, info);
    return launch_walk(perm_walk_generic<NC, false, false, NT>, P, P, NT, smem, num_sms,
                       max_grid, stream, info);
}

cudaError_t launch_generic(int ncp, bool binary, bool unitcols, const WalkParams &P,
                           int num_sms, int max_grid, cudaStream_t stream, LaunchInfo *info)
{
    switch (ncp) {
#define PQ_CASE(N)                                                                      \
    case N:                                                                             \
        return launch_generic_nc<N>(binary, unitcols, P, num_sms, max_grid, stream, info);
        PQ_CASE(4)
        PQ_CASE(8)
        PQ_CASE(12)
        PQ_CASE(16)
        PQ_CASE(20)
        PQ_CASE(24)
        PQ_CASE(28)
        PQ_CASE(32)
        PQ_CASE(36)
        PQ_CASE(40)
        PQ_CASE(44)
        PQ_CASE(48)
        PQ_CASE(52)
        PQ_CASE(56)
        PQ_CASE(60)
        PQ_CASE(64)
#undef PQ_CASE
    default:
        return cudaErrorInvalidValue;
    }
}

// ---- DFMA probe: the FP64 roofline denominator, measured -------------------
// 8 independent FMA chains per thread, no memory traffic.
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double *sink)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 1.0 - 1e-12, c = 1e-13;
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, m, c);
        a1 = __fma_rn(a1, m, c);
        a2 = __fma_rn(a2, m, c);
        a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c);
        a5 = __fma_rn(a5, m, c);
        a6 = __fma_rn(a6, m, c);
        a7 = __fma_rn(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678)
        sink[0] = s;
}

cudaError_t launch_dfma_probe(int num_sms, int iters, double *sink, cudaStream_t stream,
                              double *flops)
{
    const int grid = num_sms * 8;
    dfma_probe_kernel<<<grid, 256, 0, stream>>>(iters, sink);
    *flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)grid;
    return cudaGetLastError();
}

} // namespace pqperm
