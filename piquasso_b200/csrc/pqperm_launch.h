// pqperm_launch.h -- internal interface between the kernel translation units
// and the C-ABI host code (pqperm_api.cu).  Not installed.
#pragma once

#include <cuda_runtime.h>

#include "pqperm_device.cuh"

namespace pqperm {

struct LaunchInfo {
    int grid = 0;
    int block = 0;
    size_t smem = 0;
    int blocks_per_sm = 0;
};

// Width the generic kernel is instantiated for: smallest multiple of 4 >= nc.
inline int generic_padded_cols(int nc) { return nc <= 4 ? 4 : (nc + 3) / 4 * 4; }

// Launch the generic walk for P (P.partials must hold max_grid * 4 doubles).
cudaError_t launch_generic(int ncp, bool binary, bool unitcols, const WalkParams &P,
                           int num_sms, int max_grid, cudaStream_t stream,
                           LaunchInfo *info);

// Binary hypercube walk (kernel 2).  With `h_A2` (host, (P.D+1) x nc double2) and
// nc <= kBinMaxParamCols the matrix rides in the kernel's parameter block;
// otherwise `d_A2` (device, same layout) is copied into the kernel's __constant__
// matrix on `stream` first.
cudaError_t launch_binary(int nc, int B, const WalkParams &P, const double *h_A2,
                          const double2 *d_A2, int num_sms, int max_grid, cudaStream_t stream,
                          LaunchInfo *info);

// DFMA throughput probe: returns flops executed, time via events by caller.
cudaError_t launch_dfma_probe(int num_sms, int iters, double *sink, cudaStream_t stream,
                              double *flops);

} // namespace pqperm

namespace pqperm {

// mode: 0 leave-one-out sums only, 1 plus the full product, 2 full product only
cudaError_t launch_laplace(int S, int NCL, bool unitcols, int mode, const LapParams &P,
                           int total_blocks, size_t smem, cudaStream_t stream);
// batched permanents, hypercube flavour (unit columns, the three lowest digits binary)
cudaError_t launch_perm_hyper(int nc, const LapParams &P, int total_blocks, size_t smem,
                              cudaStream_t stream);
cudaError_t launch_laplace_reduce(const LapParams &P, int ncp1, cudaStream_t stream);
// out[j] = res[map[j]] for j < ncols (all device pointers)
cudaError_t launch_laplace_scatter(const double2 *res, const int *map, int ncols, double2 *out,
                                   cudaStream_t stream);
// pmf rows of the Clifford-Clifford sampler from the Laplace results of P (device)
cudaError_t launch_sampler_pmf(const LapParams &P, int ncp1, const double2 *U, int d,
                               double *pmf, cudaStream_t stream);
// index[i] = numpy's Generator.choice(d, p = pmf row i normalised) for the variate u[i]
cudaError_t launch_sampler_draw(const double *pmf, int n, int d, const double *u, int *index,
                                cudaStream_t stream);

} // namespace pqperm
