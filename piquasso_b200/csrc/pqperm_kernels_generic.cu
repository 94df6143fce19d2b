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

#ifdef PQ_TRACE
extern "C" int pq_debug_trace_read(unsigned long long *out, int n)
{
    return (int)cudaMemcpyFromSymbol(out, pq_trace_buf, (size_t)n * 8);
}
#endif

// ---- DFMA probe: the FP64 roofline denominator, measured -------------------
// 16 independent FMA chains per thread, every instruction with three distinct
// register operands, no memory traffic: 62-63 of the 64 FMA/clk/SM on B200
// (tools/dfma_operands.cu; a loop whose chains share two operands, as round 1 used,
// stops at 56-58 -- the denominator must be the best the pipe can do).
__global__ void __launch_bounds__(64) dfma_probe_kernel(int iters, double *sink,
                                                        const double *in)
{
    double x[16], y[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        x[i] = in[i];
        y[i] = in[16 + i] + threadIdx.x * 1e-12; // per-thread: stays in vector registers
        z[i] = in[32 + i];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            x[i] = __fma_rn(y[i], z[i], x[i]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        s += x[i];
    if (s == 12345.678)
        sink[0] = s;
}

// `sink` must hold 64 doubles: [0] is the never-taken output, [8..56) the operands.
cudaError_t launch_dfma_probe(int num_sms, int iters, double *sink, cudaStream_t stream,
                              double *flops)
{
    static double h_in[48];
    for (int i = 0; i < 48; i++)
        h_in[i] = 1.0 + 1e-9 * i;
    cudaError_t e = cudaMemcpyAsync(sink + 8, h_in, sizeof(h_in), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
        return e;
    const int grid = num_sms * 8; // 8 CTAs of 64 threads per SM: 16 warps
    dfma_probe_kernel<<<grid, 64, 0, stream>>>(iters, sink, sink + 8);
    *flops = 2.0 * 16.0 * (double)iters * 64.0 * (double)grid;
    return cudaGetLastError();
}

} // namespace pqperm
