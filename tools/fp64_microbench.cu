// Development microbenchmark: FP64 pipe throughput vs resident warps and ILP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(int iters, double *sink, double m, double c)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = __fma_rn(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

// complex product chain like the permanent: p *= s_j over NCOL register-resident values, CH chains
template <int NCOL, int CH>
__global__ void kc(int iters, double *sink, double m)
{
    double sr[NCOL], si[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; j++) { sr[j] = 1.0 + 1e-6 * (threadIdx.x + j); si[j] = 1e-6 * j; }
    double accr = 0, acci = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NCOL; j++) { sr[j] = __fma_rn(m, 1e-9, sr[j]); si[j] = __fma_rn(m, -1e-9, si[j]); }
        double cr[CH], ci[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int j0 = NCOL * c / CH, j1 = NCOL * (c + 1) / CH;
            cr[c] = sr[j0]; ci[c] = si[j0];
#pragma unroll
            for (int j = j0 + 1; j < j1; j++) {
                double nr = __fma_rn(cr[c], sr[j], -(ci[c] * si[j]));
                double ni = __fma_rn(cr[c], si[j], ci[c] * sr[j]);
                cr[c] = nr; ci[c] = ni;
            }
        }
#pragma unroll
        for (int c = 1; c < CH; c++) {
            double nr = __fma_rn(cr[0], cr[c], -(ci[0] * ci[c]));
            double ni = __fma_rn(cr[0], ci[c], ci[0] * cr[c]);
            cr[0] = nr; ci[0] = ni;
        }
        accr += cr[0]; acci += ci[0];
        m = -m;
    }
    if (accr + acci == 12345.678) sink[0] = accr;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    double *sink; cudaMalloc(&sink, 8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d clock %d kHz\n", sms, clk);
    const int iters = 1 << 15;
    printf("DFMA: warps/SM x ILP -> FMA per clk per SM (peak 64) assuming %.3f GHz\n", clk * 1e-6);
    int wps[] = {4, 8, 12, 16, 32, 64};
    for (int w : wps) {
        // one block of 32*? threads per SM slot: use blocks of 64 threads (2 warps)
        int blocks = sms * (w / 2);
        auto rate = [&](float ms, int ilp) { return (double)iters * ilp * 64.0 * blocks / (ms * 1e-3) / (clk * 1e3) / sms; };
        float t1 = timeit([&] { k<1><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        float t2 = timeit([&] { k<2><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        float t4 = timeit([&] { k<4><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        float t8 = timeit([&] { k<8><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        float t16 = timeit([&] { k<16><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        printf("warps/SM %2d: ILP1 %.1f ILP2 %.1f ILP4 %.1f ILP8 %.1f ILP16 %.1f\n", w, rate(t1, 1), rate(t2, 2), rate(t4, 4), rate(t8, 8), rate(t16, 16));
    }
    printf("complex chain kernel (NCOL=40): FP64 instr per clk per SM (peak 64)\n");
    for (int w : {8, 16}) {
        int blocks = sms * (w / 2);
        const int it2 = 1 << 12;
        auto rate = [&](float ms) { return (double)it2 * (6 * 40 - 2 + 2) * 64.0 * blocks / (ms * 1e-3) / (clk * 1e3) / sms; };
        float a = timeit([&] { kc<40, 1><<<blocks, 64>>>(it2, sink, 1.0); });
        float b = timeit([&] { kc<40, 2><<<blocks, 64>>>(it2, sink, 1.0); });
        float c = timeit([&] { kc<40, 4><<<blocks, 64>>>(it2, sink, 1.0); });
        float d = timeit([&] { kc<40, 8><<<blocks, 64>>>(it2, sink, 1.0); });
        printf("warps/SM(req) %2d: CH1 %.1f CH2 %.1f CH4 %.1f CH8 %.1f\n", w, rate(a), rate(b), rate(c), rate(d));
    }
    return 0;
}
