// Development microbenchmark: does the FP64 pipe care where a DFMA's operands come from?
//   A: x[i] = fma(y, z, x[i])        -- two operands shared by all chains (register reuse cache)
//   B: x[i] = fma(y[i], z[i], x[i])  -- three distinct registers per instruction, no reuse
//   C: x[i] = y[i] * z[i] + rotating -- DMUL with two distinct registers
//   D: complex FMA pattern of the Laplace leaf accumulation
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_operands tools/dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void kA(int iters, double *sink, double y, double z)
{
    double x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __fma_rn(y, z, x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i];
    if (s == 12345.678) sink[0] = s;
}

template <int N>
__global__ void kB(int iters, double *sink, const double *in)
{
    double x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[i]; y[i] = in[N + i] + threadIdx.x * 1e-12; z[i] = in[2 * N + i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = __fma_rn(y[i], z[i], x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i];
    if (s == 12345.678) sink[0] = s;
}

template <int N>
__global__ void kC(int iters, double *sink, const double *in)
{
    double x[N], y[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = in[i]; y[i] = in[N + i] + threadIdx.x * 1e-12; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = x[i] * y[i];
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += x[i];
    if (s == 12345.678) sink[0] = s;
}

// N accumulators (re, im), each fed by its own (p, s) pair: acc += p * s (complex), 4 DFMA
template <int N>
__global__ void kD(int iters, double *sink, const double *in)
{
    double ar[N], ai[N], pr[N], pi[N], sr[N], si[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        ar[i] = ai[i] = 0.0;
        pr[i] = in[i] + threadIdx.x * 1e-12; pi[i] = in[N + i]; sr[i] = in[2 * N + i]; si[i] = in[3 * N + i];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            ar[i] = __fma_rn(pr[i], sr[i], __fma_rn(-pi[i], si[i], ar[i]));
            ai[i] = __fma_rn(pr[i], si[i], __fma_rn(pi[i], sr[i], ai[i]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += ar[i] + ai[i];
    if (s == 12345.678) sink[0] = s;
}

template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    double *sink, *in; cudaMalloc(&sink, 8); cudaMalloc(&in, 64 * 8);
    double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 1 << 15;
    printf("FP64 instructions per clk per SM (peak 64); N = independent chains per thread\n");
    for (int w : {8, 12, 16}) {
        int blocks = sms * (w / 2);
        auto rate = [&](float ms, double per_iter) { return (double)iters * per_iter * 64.0 * blocks / (ms * 1e-3) / (clk * 1e3) / sms; };
        float a = timeit([&] { kA<8><<<blocks, 64>>>(iters, sink, 0.999999, 1e-7); });
        float b = timeit([&] { kB<8><<<blocks, 64>>>(iters, sink, in); });
        float c = timeit([&] { kC<8><<<blocks, 64>>>(iters, sink, in); });
        float d = timeit([&] { kD<8><<<blocks, 64>>>(iters, sink, in); });
        float b16 = timeit([&] { kB<16><<<blocks, 64>>>(iters, sink, in); });
        printf("warps/SM %2d: A shared operands %.1f | B three distinct registers %.1f (16 chains %.1f) | C DMUL two distinct %.1f | D complex FMA %.1f\n",
               w, rate(a, 8), rate(b, 8), rate(b16, 16), rate(c, 8), rate(d, 32));
    }
    return 0;
}
