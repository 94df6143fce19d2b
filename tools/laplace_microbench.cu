// Development microbenchmark: the dependency structure of one Laplace step
// (update, chunked suffix chains, chunk starts, chunked prefix chains, accumulate)
// on registers only -- which pipe utilisation does the STRUCTURE allow?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cmul(double &pr, double &pi, double sr, double si)
{
    const double nr = __fma_rn(pr, sr, -(pi * si));
    const double ni = __fma_rn(pr, si, pi * sr);
    pr = nr; pi = ni;
}

template <int NCL, int CH, bool SHFL>
__global__ void __launch_bounds__(128) k(int iters, double *sink, double dr, double di)
{
    constexpr int CLEN = (NCL + CH - 1) / CH;
    double sr[NCL], si[NCL], accr[NCL], acci[NCL];
#pragma unroll
    for (int j = 0; j < NCL; j++) { sr[j] = 1.0 + 1e-3 * (threadIdx.x + j); si[j] = 1e-3 * j; accr[j] = acci[j] = 0; }
    double fullr = 0, fulli = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NCL; j++) { sr[j] = __fma_rn(dr, 1e-9, sr[j]); si[j] = __fma_rn(di, 1e-9, si[j]); }
        dr = -dr;
        double sufr[NCL], sufi[NCL];
#pragma unroll
        for (int i = CLEN - 1; i >= 0; i--)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int j0 = (NCL * c) / CH, j1 = (NCL * (c + 1)) / CH, j = j0 + i;
                if (j < j1) {
                    double tr = sr[j], ti = si[j];
                    if (j + 1 < j1) cmul(tr, ti, sufr[j + 1], sufi[j + 1]);
                    sufr[j] = tr; sufi[j] = ti;
                }
            }
        double lr = sufr[0], li = sufi[0];
#pragma unroll
        for (int c = 1; c < CH; c++) cmul(lr, li, sufr[(NCL * c) / CH], sufi[(NCL * c) / CH]);
        double olr = lr, oli = li;
        if (SHFL) { olr = __shfl_xor_sync(0xffffffffu, lr, 1); oli = __shfl_xor_sync(0xffffffffu, li, 1); }
        double prer[CH], prei[CH];
        double lor = olr, loi = oli;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            double hr = 1.0, hi = 0.0; bool hh = false;
#pragma unroll
            for (int c2 = CH - 1; c2 > c; c2--) {
                const int f = (NCL * c2) / CH;
                if (!hh) { hr = sufr[f]; hi = sufi[f]; hh = true; } else cmul(hr, hi, sufr[f], sufi[f]);
            }
            prer[c] = lor; prei[c] = loi;
            if (hh) cmul(prer[c], prei[c], hr, hi);
            cmul(lor, loi, sufr[(NCL * c) / CH], sufi[(NCL * c) / CH]);
        }
        fullr += lor; fulli += loi;
#pragma unroll
        for (int i = 0; i < CLEN; i++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int j0 = (NCL * c) / CH, j1 = (NCL * (c + 1)) / CH, j = j0 + i;
                if (j < j1) {
                    double pr = prer[c], pi = prei[c];
                    if (j + 1 < j1) {
                        double nr = pr, ni = pi; cmul(nr, ni, sr[j], si[j]); prer[c] = nr; prei[c] = ni;
                        cmul(pr, pi, sufr[j + 1], sufi[j + 1]);
                    }
                    accr[j] = __fma_rn(dr, pr, accr[j]); acci[j] = __fma_rn(dr, pi, acci[j]);
                }
            }
    }
    double s = fullr + fulli;
#pragma unroll
    for (int j = 0; j < NCL; j++) s += accr[j] + acci[j];
    if (s == 12345.678) sink[0] = s;
}

template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

template <int NCL, int CH, bool SHFL> void run(const char *name, int sms, int clk, double *sink)
{
    const int iters = 4096;
    for (int bps : {2, 3}) {
        int blocks = sms * bps;
        float ms = timeit([&] { k<NCL, CH, SHFL><<<blocks, 128>>>(iters, sink, 1.0, -1.0); });
        // FP64 per iteration: 2*NCL update + 4 * (3*NCL cmuls approx) + 2*NCL acc = 16*NCL
        double instr = (double)iters * 16.0 * NCL * 128.0 * blocks;
        printf("%s NCL=%d CH=%d shfl=%d blocks/SM=%d: %.3f ms  %.1f FP64 instr/clk/SM (16*NCL model)\n", name, NCL, CH, (int)SHFL, bps, ms,
               instr / (ms * 1e-3) / (clk * 1e3) / sms);
    }
}

int main()
{
    double *sink; cudaMalloc(&sink, 8);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    run<13, 1, false>("lap", sms, clk, sink);
    run<13, 2, false>("lap", sms, clk, sink);
    run<13, 3, false>("lap", sms, clk, sink);
    run<13, 4, false>("lap", sms, clk, sink);
    run<13, 6, false>("lap", sms, clk, sink);
    run<13, 3, true>("lap", sms, clk, sink);
    run<7, 2, true>("lap", sms, clk, sink);
    return 0;
}
