/*
 * pqperm.h -- C ABI of libpqperm.so: the B200 (sm_100a) implementation of
 * piquasso's permanent hot path.
 *
 * Every entry point takes plain pointers and sizes (no torch / numpy / C++
 * types) and returns 0 on success or a PQ_ERR_* code; pq_last_error() gives the
 * message for the calling thread.  No exception crosses this boundary and
 * there is NO CPU fallback: without a usable CUDA device the compute calls
 * return PQ_ERR_NO_DEVICE.
 *
 * "Reference" below is Budapest-Quantum-Computing-Group/piquasso 8.0.1; each
 * entry cites the reference interface it replaces (paths relative to the
 * reference tree).
 *
 * Layout conventions (same as the reference's Matrix<std::complex<T>>,
 * src/matrix.hpp:152-260): matrices are row-major, `R` rows by `C` columns,
 * complex entries interleaved (re, im); `rows` has R and `cols` has C int32
 * multiplicities.  Inputs are borrowed for the duration of the call and never
 * written.
 */
#ifndef PQPERM_H
#define PQPERM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQ_OK 0
#define PQ_ERR_SUM_MISMATCH 1 /* sum(rows) != sum(cols): the reference throws, src/permanent.cpp:97-104 */
#define PQ_ERR_BAD_ARG 2
#define PQ_ERR_NO_DEVICE 3    /* no CUDA device / driver: there is no CPU fallback */
#define PQ_ERR_CUDA 4
#define PQ_ERR_TOO_LARGE 5    /* beyond the limits below */

/* Limits (the reference has none, but its run time is exponential in the rows):
 * rows with non-zero multiplicity after the split <= PQ_MAX_DIGITS, multiplicities
 * <= 254, term space <= 2^62.  Columns with non-zero multiplicity: <= PQ_MAX_COLS
 * for the partitioned entries (pq_perm_partial_c128, pq_perm_job_*, pq_perm_plan);
 * <= PQ_MAX_COLS_WIDE for pq_perm_c128 / pq_perm_c64 / pq_perm_batch_c128, the
 * permanent_laplace entries and the sampler steps -- beyond PQ_MAX_COLS a whole
 * warp walks each Gray segment, and (active rows + 1) x columns x 16 bytes plus
 * 36 KB of accumulators must fit in 224 KB of shared memory. */
#define PQ_MAX_COLS 64
#define PQ_MAX_COLS_WIDE 256
#define PQ_MAX_DIGITS 64

/* Message of the last failing call made by this thread ("" if none). */
const char *pq_last_error(void);

/* Number of usable CUDA devices (0 when there is no driver/GPU). */
int pq_device_count(void);

/* Devices used by the host-buffer entry points below; default = device 0 only.
 * With n > 1 a single permanent's term space is split over the devices and
 * the partial sums are combined (single process; the multi-process path is
 * pq_perm_partial_c128 + an NCCL all-gather done by the caller). */
int pq_set_devices(const int *device_ids, int n);

/* ---------------------------------------------------------------------
 * permanent(matrix, rows, cols) -> complex scalar
 * Replaces permanent_cpp<double> / permanent_cpp<float>
 * (src/permanent.hpp:32-34, src/permanent.cpp:49-264) as bound by
 * permanent_np<T> (piquasso/_math/permanent.cpp:26-41).
 * The c64 entry converts to double, computes in FP64 and rounds once.
 * ------------------------------------------------------------------- */
int pq_perm_c128(const double *A, int R, int C, const int32_t *rows,
                 const int32_t *cols, double out[2]);
int pq_perm_c64(const float *A, int R, int C, const int32_t *rows,
                const int32_t *cols, float out[2]);

/* ---------------------------------------------------------------------
 * permanent_laplace(matrix, rows, cols) -> complex vector
 * Replaces permanent_laplace_cpp<T> (src/permanent_laplace.hpp:32-34,
 * src/permanent_laplace.cpp:43-240) as bound by permanent_laplace_np<T>
 * (piquasso/_math/permanent.cpp:43-58).  `out` must hold 2*max(C,1) values;
 * *out_len receives the number of complex results: 1 on the reference's
 * early-out (src/permanent_laplace.cpp:52-57), otherwise C.
 * ------------------------------------------------------------------- */
int pq_perm_laplace_c128(const double *A, int R, int C, const int32_t *rows,
                         const int32_t *cols, double *out, int *out_len);
int pq_perm_laplace_c64(const float *A, int R, int C, const int32_t *rows,
                        const int32_t *cols, float *out, int *out_len);

/* ---------------------------------------------------------------------
 * One LARGE permanent_laplace over several GPUs (the
 * term-space split of a single Laplace call): rank `part` of `nparts` walks the
 * contiguous share [nseg*part/nparts, nseg*(part+1)/nparts) of the problem's
 * Gray-code segments -- the rule of src/permanent.cpp:158-164 applied to
 * permanent_laplace_cpp's loop (src/permanent_laplace.cpp:120-134) -- and
 * returns its partial sums, already scaled by 2^-(sum_rows-1), in the layout of
 * pq_perm_laplace_c128.  The results of all parts add up to
 * permanent_laplace(A, rows, cols); the exchange step is one sum (all-gather +
 * add, or all-reduce) of C complex numbers per rank.  On the reference's
 * early-out part 0 returns [1] and the others [0].
 * ------------------------------------------------------------------- */
int pq_perm_laplace_partial_c128(const double *A, int R, int C, const int32_t *rows,
                                 const int32_t *cols, int part, int nparts, double *out,
                                 int *out_len);

/* The same with the rank's partial sums LEFT ON THE DEVICE: d_out (memory of
 * `device`, 2*C doubles, layout of pq_perm_laplace_c128) is written by the library's
 * stream and the call returns after that stream has been synchronised, so the
 * caller can hand d_out straight to the collective (ncclAllGather / all-reduce)
 * without a trip through the host.  *out_len = C, or 1 on the reference's early-out;
 * trivial[2] is NaN unless the problem was that early-out, in which case it holds the
 * value (part 0: 1, others: 0) and d_out is untouched. */
int pq_perm_laplace_partial_dev_c128(const double *A, int R, int C, const int32_t *rows,
                                     const int32_t *cols, int part, int nparts, int device,
                                     double *d_out, double *trivial, int *out_len);

/* ---------------------------------------------------------------------
 * Batch of independent permanent_laplace problems in one call (what the
 * Clifford-Clifford sampler issues once per photon per shot,
 * piquasso/_simulators/passive/sampling.py:723-734).  Problem b has
 * R[b] x C[b] matrix at A + 2*a_off[b] doubles, multiplicities at
 * rows + r_off[b] / cols + c_off[b]; its C[b] (or 1) results are written at
 * out + 2*o_off[b], and out_len[b] is set as above.
 * ------------------------------------------------------------------- */
int pq_perm_laplace_batch_c128(int nprob, const double *A, const int64_t *a_off,
                               const int32_t *R, const int32_t *C,
                               const int32_t *rows, const int64_t *r_off,
                               const int32_t *cols, const int64_t *c_off,
                               double *out, const int64_t *o_off,
                               int32_t *out_len);

/* ---------------------------------------------------------------------
 * Many permanents of ONE matrix with different multiplicity vectors in one
 * call: out[b] = permanent(A, row_mult[b*R ..], col_mult[b*C ..]).  This is
 * the batched form of connector.permanent(interferometer, cols=input,
 * rows=output) (piquasso/_simulators/passive/utils.py:131-138,
 * probabilities.py:26-54) for tables of detection amplitudes: the matrix is
 * uploaded once and every CTA gathers its own minor.  A sum mismatch in any
 * problem fails the whole call with PQ_ERR_SUM_MISMATCH.  Any number of problems:
 * the batch is planned (on host threads), staged and launched in chunks of 2^17.
 * ------------------------------------------------------------------- */
int pq_perm_batch_c128(const double *A, int R, int C, int nprob, const int32_t *row_mult,
                       const int32_t *col_mult, double *out);

/* ---------------------------------------------------------------------
 * One photon step of the Clifford-Clifford sampler for `nshots` independent
 * shots sharing one d x d interferometer U (row-major complex128).  For shot s,
 * with the output occupations placed so far out_occ[s*d .. ] and the input
 * occupations grown so far in_occ[s*d .. ], computes what _calculate_pmf
 * (piquasso/_simulators/passive/sampling.py:723-749) computes before
 * normalisation:   pmf[s*d + m] = | sum_j in_j * partial_j * U[m, nz_j] |^2,
 * partial = permanent_laplace(U[out>0][:, in>0], out[out>0], in[in>0]).
 * Zero filtering, the batched Laplace walk and the pmf assembly all happen
 * inside the call (the pmf on the device, as an epilogue of the walk); the
 * host keeps the per-shot RNG and draws from the returned rows.
 * ------------------------------------------------------------------- */
int pq_sampler_pmf_c128(const double *U, int d, int nshots, const int32_t *out_occ,
                        const int32_t *in_occ, double *pmf);

/* ---------------------------------------------------------------------
 * The same photon step including the draw (_calculate_pmf's normalisation and
 * _sample_from_pmf, piquasso/_simulators/passive/sampling.py:736-753): the pmf
 * rows stay on the device and index[s] is the output mode that
 * rng.choice(arange(d), p=pmf) returns when the shot's generator yields the
 * uniform variate u[s] -- numpy's Generator.choice(a, p=p) is
 * searchsorted(cumsum(p) / cumsum(p)[-1], rng.random(), side="right"), so the
 * caller draws u[s] = rng.random() from shot s's own generator (which does not
 * depend on the pmf) and this call repeats numpy's arithmetic in numpy's
 * order.  index[s] = -1 marks a row numpy would reject (NaN probabilities).
 * ------------------------------------------------------------------- */
int pq_sampler_draw_c128(const double *U, int d, int nshots, const int32_t *out_occ,
                         const int32_t *in_occ, const double *u, int32_t *index);

/* The same on an explicit CUDA device.  Steps issued by different host threads
 * for DIFFERENT devices run concurrently (each device has its own lock, scratch
 * and stream): shots are independent (sampling.py:149-194 gives every shot its
 * own generator), so a single process shards them over the GPUs of a box with
 * one thread per device and no exchange step. */
int pq_sampler_draw_dev_c128(int device, const double *U, int d, int nshots,
                             const int32_t *out_occ, const int32_t *in_occ, const double *u,
                             int32_t *index);
int pq_sampler_pmf_dev_c128(int device, const double *U, int d, int nshots,
                            const int32_t *out_occ, const int32_t *in_occ, double *pmf);

/* Host helper of the sampler (no GPU involved): the raw 64-bit streams of the numpy
 * generators the reference gives its shots, np.random.default_rng(seed0 + i) for
 * i in [0, n) (piquasso/_simulators/passive/sampling.py:149-194), `draws` outputs
 * each: out[i * draws + k] == np.random.PCG64(seed0 + i).random_raw(draws)[k].
 * Restates numpy's SeedSequence + PCG64 seeding for integer seeds below 2^64. */
int pq_pcg64_streams(uint64_t seed0, int64_t n, int draws, uint64_t *out);

/* Where the calling thread's last pq_sampler_* call spent its wall time, in
 * milliseconds: out_ms[0] planning on the host (zero filtering, problem
 * descriptors), [1] waiting for the device's lock, [2] the device phase (uploads,
 * kernels, download, scatter), [3] the kernels alone (CUDA events).  Diagnostic
 * only; the reference has no counterpart. */
void pq_last_sampler_profile(double out_ms[4]);

/* Finer split of the device phase of the calling thread's last pq_sampler_* call
 * (ms): [0] scratch growth, [1] staging descriptors into pinned memory, [2]
 * enqueueing copies and launches, [3] waiting for the stream, [4] scattering the
 * results, [5] uploading the interferometer; [6..7] reserved.  Diagnostic only. */
void pq_last_sampler_detail(double out_ms[8]);

/* Work done by all pq_sampler_* steps of this process since the last reset:
 * out[0] = Gray-code terms walked, out[1] = algorithmic flops (22 k per term of
 * a k-column Laplace problem).  For roofline reports. */
void pq_sampler_work(double out[2]);
void pq_sampler_work_reset(void);

/* ---------------------------------------------------------------------
 * Partitioned permanent: the piece of one permanent that rank `part` of
 * `nparts` owns.  The term space [0, idx_max) (src/permanent.cpp:131-142) is
 * cut into equal-length Gray-code segments (one per GPU thread; the
 * hierarchical form of src/permanent.cpp:158-164) and rank `part` walks the
 * contiguous segment range [nseg*part/nparts, nseg*(part+1)/nparts).
 *
 * The UNSCALED partial sum is left in DEVICE memory as four doubles
 * (re_hi, re_lo, im_hi, im_lo; value = hi + lo) at d_partial on `device`,
 * enqueued on `stream` (a cudaStream_t, NULL = the library's own stream,
 * synchronised before returning).  The caller gathers the four doubles of
 * every rank (one NCCL all-gather), sums them with pq_perm_combine and calls
 * pq_perm_finish.
 *
 * The walk kernels of one device share scratch memory (the staged matrix, the
 * per-CTA partials, the segment dispenser).  The library orders its own use of
 * that scratch on the device: a launch that follows one enqueued on a DIFFERENT
 * stream first waits for it (cudaStreamWaitEvent), so callers need not
 * synchronise their stream before the next library call.
 *
 * *status (host int, may be NULL) receives 0 when a partial was enqueued or
 * 1 when the problem is one of the reference's trivial cases and `trivial`
 * already holds the final value (src/permanent.cpp:106-122).
 * ------------------------------------------------------------------- */
int pq_perm_partial_c128(const double *A, int R, int C, const int32_t *rows,
                         const int32_t *cols, int part, int nparts, int device,
                         void *stream, double *d_partial, int *status,
                         double trivial[2]);

/* Error-free sum of n partial quadruples (re_hi, re_lo, im_hi, im_lo) into
 * one, in index order.  Rank partials can be orders of magnitude larger than
 * their sum, so they must not be added component-wise in plain doubles: gather
 * them (one NCCL all-gather of 4 doubles per rank) and combine here. */
int pq_perm_combine(const double *quads, int n, double out4[4]);

/* hi/lo quadruple + sum(rows) -> permanent: (hi+lo) * 2^-(sum_rows-1)
 * (src/permanent.cpp:259). */
int pq_perm_finish(const double partial[4], int sum_rows, double out[2]);

/* ---------------------------------------------------------------------
 * Introspection used by the parity tests and the bench.
 * ------------------------------------------------------------------- */
typedef struct pq_plan_info {
    int64_t idx_max;       /* number of Gray-code terms, prod(limits) */
    int64_t seg_len;       /* terms per segment (per GPU thread) */
    int64_t nseg;          /* number of segments = idx_max / seg_len */
    int32_t active_rows;   /* Gray digits with radix >= 2 */
    int32_t active_cols;   /* columns with multiplicity > 0 */
    int32_t low_digits;    /* digits walked inside a segment */
    int32_t kernel;        /* 1 = generic n-ary walk, 2 = binary constant-bank walk
                              (pq_perm_batch_plan: 3 / 4, see there) */
    int32_t cols_padded;   /* register-resident row sums per thread */
    int32_t sum_rows;
    int32_t trivial;       /* 1 = handled by a reference early-out */
    double flops_per_term; /* 2*C + 6*M + 2 */
} pq_plan_info;

/* ---------------------------------------------------------------------
 * Resident jobs: the same partitioned permanent with the inputs uploaded
 * ONCE (create) and the kernels launched any number of times (launch), e.g.
 * to time the kernels with the inputs already in HBM.  One job is resident
 * per device; a job evicted by another call re-uploads itself on launch.
 * status / trivial as for pq_perm_partial_c128 (*job stays NULL when the
 * problem is trivial).
 * ------------------------------------------------------------------- */
typedef struct pq_perm_job pq_perm_job;
int pq_perm_job_create_c128(const double *A, int R, int C, const int32_t *rows,
                            const int32_t *cols, int part, int nparts, int device,
                            pq_perm_job **job, int *status, double trivial[2]);
int pq_perm_job_launch(pq_perm_job *job, void *stream, double *d_partial);
int pq_perm_job_info(const pq_perm_job *job, pq_plan_info *info);
int64_t pq_perm_job_terms(const pq_perm_job *job); /* Gray-code terms this rank walks */
int pq_perm_job_destroy(pq_perm_job *job);

/* Plan that pq_perm_c128 would use for these multiplicities (no GPU needed). */
int pq_perm_plan(int R, int C, const int32_t *rows, const int32_t *cols,
                 pq_plan_info *info);

/* Plan that pq_perm_batch_c128 would use for ONE problem of a batch of `nprob`
 * (no GPU needed).  kernel: 3 = batched walk, term by term (one lane per Gray
 * segment up to 32 columns, lane-split beyond), 4 = batched walk, hypercube
 * flavour (three rows of multiplicity 1 are the low digits, blocks of 8 terms).
 * cols_padded = columns the kernel holds after column multiplicities were written
 * out as unit columns (when they fit); seg_len * nseg = idx_max always. */
int pq_perm_batch_plan(int R, int C, const int32_t *rows, const int32_t *cols, int nprob,
                       pq_plan_info *info);

/* Gray digits (reference digit order, one per row after the split, i.e. R
 * entries for digits 0..R-1 of the split problem) that the GPU path assigns
 * to `offset`, computed by the same host/device code path the kernels use for
 * seeding (no GPU needed): enumeration parity against
 * src/n_aryGrayCodeCounter.hpp:170-194. */
int pq_perm_gray_of_offset(int R, const int32_t *rows, int64_t offset,
                           int32_t *gray);

/* Per-segment unscaled partial sums of one permanent, segments
 * [seg_begin, seg_begin+nseg), 2 doubles (re, im) each, written to host `out`:
 * partition-indexing parity (segment s covers offsets
 * [s*seg_len, (s+1)*seg_len)). */
int pq_perm_segment_sums_c128(const double *A, int R, int C,
                              const int32_t *rows, const int32_t *cols,
                              int64_t seg_begin, int64_t nseg, double *out);

/* Kernel-only duration (ms, CUDA events on the library stream) of the last
 * pq_perm_c128 / pq_perm_laplace* call made on `device`; -1 if none. */
double pq_last_kernel_ms(int device);

/* Durations (ms, CUDA events on the launching stream) of the most recent
 * walk+reduce launches on `device`, newest first; returns how many were
 * written.  The stream they ran on must have been synchronised. */
int pq_kernel_ms_history(int device, double *out, int max);

/* Number of kernels the library has launched on all devices since load. */
int64_t pq_launch_count(void);

/* ACCURACY ARBITER (tests, accuracy reports -- not a fast path): the permanent
 * of pq_perm_c128's arguments with every operation in double-double arithmetic
 * (~106-bit mantissa) on the library's first device, returned as
 * out = {re_hi, re_lo, im_hi, im_lo} (value = hi + lo).  Its Gray-code counter
 * restates the reference's class literally (src/n_aryGrayCodeCounter.hpp:170-254)
 * with 64-bit offsets and shares no device code with the production walks; it
 * is what arbitrates Haar-random matrices beyond n = 32, where the reference
 * itself is wrong (:179).  ~100x slower than pq_perm_c128.  At most 64 columns. */
int pq_perm_arbiter_c128(const double *A, int R, int C, const int32_t *rows,
                         const int32_t *cols, double out[4]);

/* Measured FP64 peak of `device`: a dependent-free DFMA loop over all SMs,
 * `iters` FMAs per thread; returns TFLOP/s (2 flop per FMA), <0 on error. */
double pq_fp64_peak_tflops(int device, int iters);

/* Select the permanent kernel: 0 = automatic, 1 = force the generic walk,
 * 2 = force the binary constant-bank walk where applicable (tests / bench). */
int pq_set_kernel_choice(int choice);

/* Kernel timing of the permanent entries: on (default) every walk launch is
 * bracketed by a CUDA event pair (pq_last_kernel_ms, pq_kernel_ms_history); off
 * saves three driver calls (~4 us) per permanent and pq_last_kernel_ms reports -1. */
int pq_set_timing(int on);

/* Override the segment length exponent (0 = automatic): tests only. */
int pq_set_seg_len_hint(int64_t seg_len);

#ifdef __cplusplus
}
#endif
#endif /* PQPERM_H */
