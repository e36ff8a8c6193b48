/* tfft -- B200-native fp16 complex-to-complex FFT, drop-in C ABI for the plan/execute path of
 * CPestka/Tensor-FFT (src/base).  The reference is header-only C++ with no C ABI of its own; each
 * entry point below names the reference interface it replaces (paths relative to the reference
 * root).  Conventions are the reference's: forward transform exp(-2*pi*i*n*k/N), result scaled
 * by 1/N, natural order in and out, PLANAR fp16 (one array of real parts, one of imaginary
 * parts), N a power of two >= 256 (src/base/Plan.h:85-96).
 *
 * All pointers are plain device (or, for the *_host entry points, host) pointers to IEEE
 * binary16 values; no CUDA or torch types appear in the signatures.  `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).  Functions return 0 on success, a negative
 * TFFT_E_* code for library errors and a positive cudaError_t value for CUDA errors.
 * There is no CPU fallback: without a CUDA device every exec call fails with an error.
 */
#ifndef TFFT_H_
#define TFFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfft_plan_s* tfft_plan_t;

enum {
  TFFT_OK = 0,
  TFFT_E_INVALID_SIZE = -1,   /* not a power of two, < 256 or > 2^30 (Plan.h:85-96 prints + nullopt) */
  TFFT_E_INVALID_ARG = -2,    /* null pointer, misaligned pointer or stride */
  TFFT_E_NO_DEVICE = -3,      /* no sm_100 device: the product has no CPU path */
  TFFT_E_UNSUPPORTED = -4,
  TFFT_E_NOMEM = -5,
  TFFT_E_NOT_IN_FILE = -6,    /* tuner file has no line for this length (Plan.h:238-254 prints + nullopt) */
  TFFT_E_TIMEOUT = -7         /* multi-GPU plans: a rank did not reach a phase barrier in time (tfft_mg_status) */
};

/* flags for tfft_plan_create */
enum {
  TFFT_DEFAULT = 0,
  /* multi-pass sizes (N > 32768) overwrite the input planes like the reference does
   * (src/base/ComputeFFT.h:89-145 ping-pongs between input and result buffers); set this flag
   * to make the plan own a scratch buffer instead and leave the input intact. */
  TFFT_PRESERVE_INPUT = 1,
  /* SURVEY.md 8f rank 2, the conventions of the other fp16 FFT the reference is compared with (cuFFT,
   * src/testing/unitTesting/CuFFTTest.h:25-57): inverse transform exp(+2*pi*i*n*k/N) (computed as the forward
   * transform with the real and imaginary planes exchanged on both sides: bit-identical arithmetic), and no 1/N
   * scaling (every stage's DFT matrix unscaled; intermediate values must stay inside the fp16 range). */
  TFFT_INVERSE = 2,
  TFFT_UNSCALED = 4,
  /* interleaved layout (cuFFT's half2 / the reference's CuFFTTest.h:25-57 buffers): in_re and out_re point to arrays of
   * (re, im) fp16 pairs, in_im / out_im are ignored, strides count complex elements.  Every size and the 2-D plans; these
   * plans load with 16-byte vector loads instead of TMA tiles (the split into planes happens in registers), multi-pass
   * 1-D sizes own a planar scratch buffer (the input is preserved), 2-D plans keep their intermediate in the output array. */
  TFFT_INTERLEAVED = 8
};

typedef struct tfft_plan_info_s {
  int64_t n;                 /* transform length                          (Plan.h:19 fft_length_) */
  int64_t batch;             /* transforms per exec                                                */
  int32_t r16_stages;        /* tensor-core radix-16 stages per pass                               */
  int32_t tail_radix;        /* 1, 2, 4 or 8: CUDA-core radix fused into the load phase            */
  int32_t passes;            /* HBM round trips: 1 (N <= 2^15), 2 (<= 2^23), 3 from 2^24 on          */
  int32_t results_in_results;/* always 1: results land in the output planes (Plan.h:25)            */
  int32_t amount_of_r16_steps; /* reference-compatible: log2(N)/4 - 1   (Plan.h:99)                */
  int32_t amount_of_r2_steps;  /* reference-compatible: log2(N) % 4     (Plan.h:100)               */
  int32_t transforms_per_cta;  /* pass 1 (or the only pass)                                        */
  int32_t smem_bytes;          /* dynamic shared memory per CTA, largest pass                      */
  int32_t tmem_columns;        /* tensor-memory columns per CTA, largest pass                      */
  int64_t grid;                /* CTAs of the first pass                                           */
  int64_t workspace_bytes;     /* device scratch owned by the plan                                 */
  int64_t algorithmic_bytes;   /* 8 * n * batch * passes                            */
} tfft_plan_info_t;

/* Replaces CreatePlan(fft_length, ...) (src/base/Plan.h:77-194) + PlanWorksOnDevice
 * (Plan.h:257-296).  `batch` replaces the amount_of_ffts_ of DataBatchHandler
 * (src/base/DataHandler.h:88-89).  The plan is immutable afterwards; exec calls on one plan
 * may be issued from several host threads / streams (except with workspace-owning plans). */
int tfft_plan_create(tfft_plan_t* plan, int64_t n, int64_t batch, uint32_t flags);

/* Replaces CreatePlan(fft_length, tuner_results_file) (src/base/Plan.h:197-255): the plan for length n is configured
 * from the line of `path` that starts with n.  Line format: the reference's `N mode base_warps r16_warps r2_block`
 * (written by src/testing/FileWriter.h:250-269; those four columns are accepted and ignored) optionally followed by
 * `key=value` knobs of the B200 kernels: tma, pipe, two_slot, prefetch (0/1), lg1 (log2 of the four-step column-pass
 * length), tma_col (0: cp.async, 1: TMA column tiles, 2: never the 64- / 32-column tiles), cluster (1: units of 2^16 elements shared by a CTA pair through distributed shared memory --
 * N = 65536 in ONE HBM pass, 16-column units for 4096-point column passes; off by default), ring
 * (landing-ring kernel for 32K-element units: 1 = the 4096-point column pass of four-step plans, the default; 0 = off;
 * 2 = also N = 32768 and the 4096-point row pass, where it does not pay).  For n = 2^24 a line that names lg1 or cluster
 * keeps the two-pass (four-step) plan instead of the default three passes of 256.
 * tools/tune.py measures and writes such a file.  TFFT_E_NOT_IN_FILE when no line matches.
 * The environment variable TFFT_TUNER_FILE makes tfft_plan_create consult a file the same way. */
int tfft_plan_create_from_file(tfft_plan_t* plan, int64_t n, int64_t batch, uint32_t flags, const char* path);

/* Plan for `batch` 2-D transforms of ny rows x nx columns (row-major planar images), scale
 * 1/(ny*nx).  The reference has no 2-D path. */
int tfft_plan_create_2d(tfft_plan_t* plan, int64_t ny, int64_t nx, int64_t batch, uint32_t flags);

/* Optional: do every one-time device initialisation of the plan now, on the current device (kernel module loading and
 * the > 48 KiB shared-memory opt-in, upload of the constant tables), so that later tfft_exec calls only enqueue work.
 * tfft_exec does the same lazily at first use; call this before capturing execs into a CUDA graph or before running
 * them next to kernels that spin-wait on other streams (module loading can synchronise the device). */
int tfft_plan_prepare(tfft_plan_t plan);

int tfft_plan_info(tfft_plan_t plan, tfft_plan_info_t* info);
int tfft_plan_destroy(tfft_plan_t plan);

/* Replaces ComputeFFT(plan, DataHandler) and ComputeFFT(plan, DataBatchHandler)
 * (src/base/ComputeFFT.h:54-151, :162-293).  Transform b reads in_re + b*in_stride,
 * in_im + b*in_stride and writes out_re + b*out_stride, out_im + b*out_stride (strides in
 * elements, multiples of 8; pointers 16-byte aligned).  The reference batch layout
 * [RE_0|IM_0|RE_1|IM_1|...] (DataHandler.h:105-114) is in_im = in_re + n, stride = 2n.
 * Asynchronous on `stream`; one kernel launch per pass for the whole batch. */
int tfft_exec(tfft_plan_t plan, const void* in_re, const void* in_im, void* out_re, void* out_im,
              int64_t in_stride, int64_t out_stride, void* stream);

/* Batched transform fused with the inter-pass twiddle of a four-/six-step decomposition: transform b
 * additionally multiplies its output k by exp(-2*pi*i * k * (first_col + b) / 2^log2_total).  Used by
 * the multi-GPU 1-D transform (one huge transform = columns pass + all-to-all + rows pass; the
 * reference has no multi-GPU path).  Only for n <= 32768. */
int tfft_exec_twiddled(tfft_plan_t plan, const void* in_re, const void* in_im, void* out_re, void* out_im,
                       int64_t in_stride, int64_t out_stride, int32_t log2_total, int64_t first_col, void* stream);

/* tfft_exec_twiddled for an input whose transforms are SEGMENTED: every transform is cut into `segments` equal pieces and
 * piece q of transform b starts at in + q*segment_stride + b*in_stride (the exchange buffers of the multi-GPU transform are
 * source-rank major: one block per sending rank).  The gather happens inside the TMA load of the transform (a 5-D tensor
 * map), not in a separate pass.  log2_total = 0: no twiddle.  Only for single-pass plans with n >= 2048 and `segments`
 * dividing the first radix (16 or 32); TFFT_E_UNSUPPORTED otherwise (callers then gather with a copy). */
int tfft_exec_segmented(tfft_plan_t plan, const void* in_re, const void* in_im, void* out_re, void* out_im,
                        int64_t in_stride, int64_t out_stride, int32_t segments, int64_t segment_stride,
                        int32_t log2_total, int64_t first_col, void* stream);

/* In-place execution: for single-pass sizes (n <= 32768) out_re == in_re and out_im == in_im (same strides) is legal --
 * every CTA reads its transforms completely before it writes them.  Multi-pass sizes run in place on the INPUT planes
 * by design (see TFFT_PRESERVE_INPUT) but their final pass is a transposition: out must not alias in.  2-D plans: out
 * must not alias in. */

/* Whole reference call sequence with HOST buffers: CopyDataHostToDevice -> ComputeFFT ->
 * CopyResultsDeviceToHost (src/base/DataHandler.h:45-70,124-153).  host_in / host_out hold,
 * per transform, [RE(n) | IM(n)] halves (2*n*batch values each).  Uses plan-owned device
 * buffers (a ring of four 16 MiB slots per direction, not the whole batch) and three plan-owned streams: upload,
 * transform and download of consecutive chunks overlap.  Synchronises before returning like the reference's batch
 * overload does.  Pass PINNED host memory (cudaHostAlloc / cudaHostRegister): with pageable buffers every chunk copy
 * blocks the calling thread and the three-stage pipeline degenerates to serial copies (results are the same).
 * Calls on one plan are serialised internally. */
int tfft_exec_host(tfft_plan_t plan, const void* host_in, void* host_out);

/* Pack / unpack kernels of the all-to-all exchanges of the multi-GPU 1-D transform (nothing in the
 * reference).  tfft_transpose_blocks: dst[b][c][r] = src[b][r][c] for nb0*nb1 fp16 matrices of rows x cols (multiples
 * of 64); matrix (b0, b1) starts at src + b0*src_b0 + b1*src_b1 / dst + b0*dst_b0 + b1*dst_b1, row strides in
 * elements (multiples of 8).  tfft_copy_runs: copies n0*n1*n2 contiguous runs of `run` elements between two
 * three-level strided layouts.  Asynchronous on `stream`. */
int tfft_transpose_blocks(const void* src, void* dst, int64_t rows, int64_t cols, int64_t src_row_stride,
                          int64_t dst_row_stride, int64_t nb0, int64_t nb1, int64_t src_b0, int64_t src_b1,
                          int64_t dst_b0, int64_t dst_b1, void* stream);
int tfft_copy_runs(const void* src, void* dst, int64_t run, int64_t n0, int64_t n1, int64_t n2, int64_t s0, int64_t s1,
                   int64_t s2, int64_t d0, int64_t d1, int64_t d2, void* stream);

/* ---- one transform sharded over the GPUs of a node (BASELINE config C4) ------------------------
 * The reference's multi-GPU entry points are commented-out replica code (src/base/ComputeFFT.h:295-557,
 * DataHandler.h:168-403); these are new.  One process (or host thread) per GPU, `world` a power of two <= 16, n = 2^16
 * ... 2^30 with n1/world and n2/world >= 64 (n = n1*n2, n1 = 2^ceil(lg/2)).  Rank r holds elements
 * [r*n/world, (r+1)*n/world) of the planar input and receives the same slice of the spectrum (natural order, 1/n).
 *
 *   tfft_mg_plan_create   allocates this rank's exchange buffers on the current device
 *   tfft_mg_plan_handle   fills TFFT_MG_HANDLE_BYTES that identify them (a CUDA IPC handle inside); the caller moves
 *                         the handles between the ranks with whatever transport it has (all-gather)
 *   tfft_mg_plan_connect  maps the peers' buffers; `handles` = world * TFFT_MG_HANDLE_BYTES in rank order
 *   tfft_mg_exec          transpose -> n2/world transforms of length n1 (twiddle fused) -> transpose -> n1/world
 *                         transforms of length n2 -> transpose.  Each transpose is ONE kernel that stores 64x64 tiles
 *                         straight into the owning peer's buffer over NVLink, followed by a flag barrier between the
 *                         ranks; there is no NCCL call and no pack/unpack pass.  Asynchronous on `stream`.  The result
 *                         stays in plan-owned planes (tfft_mg_plan_info: result_re / result_im, valid until the next
 *                         exec); pass out_re / out_im to have it copied out, or NULL for zero-copy use.
 *   tfft_mg_status        TFFT_E_TIMEOUT if some barrier gave up waiting for a peer (default 10 s; results invalid)
 *   tfft_mg_exec_phase    one phase of tfft_mg_exec WITHOUT its flag barrier, for callers that order the ranks themselves
 *                         (one host thread driving several GPUs with events; the single-GPU tests, which run the ranks
 *                         of a phase one after the other on one stream): phase 0 = exchange 1, 1 = transforms over the
 *                         first factor + exchange 2, 2 = transforms over the second factor + exchange 3, 3 = copy out.
 *                         Every rank must have finished phase p before any rank starts phase p + 1.
 * All ranks must call tfft_mg_exec the same number of times.  Destroy a plan only after every rank has finished its last
 * exec (synchronise the streams, then a barrier of the caller's transport): peers store into each other's buffers. */
#define TFFT_MG_HANDLE_BYTES 128
typedef struct tfft_mg_plan_s* tfft_mg_plan_t;
typedef struct tfft_mg_info_s {
  int64_t n, n1, n2;             /* n = n1 * n2                                                              */
  int64_t local_elems;           /* n / world complex elements per rank                                      */
  int32_t rank, world;
  int32_t exchanges;             /* 3                                                                        */
  int32_t reserved;
  int64_t exchange_bytes_per_rank; /* fp16 planar bytes this rank sends over NVLink per exec:
                                      exchanges * (world-1)/world * 4*n/world            */
  int64_t device_bytes;          /* plan-owned device memory on this rank                                    */
  void* result_re;               /* plan-owned result planes (local_elems halves each)                       */
  void* result_im;
} tfft_mg_info_t;
int tfft_mg_plan_create(tfft_mg_plan_t* plan, int64_t n, int32_t rank, int32_t world, uint32_t flags);
int tfft_mg_plan_handle(tfft_mg_plan_t plan, void* handle);
int tfft_mg_plan_connect(tfft_mg_plan_t plan, const void* handles);
int tfft_mg_plan_info(tfft_mg_plan_t plan, tfft_mg_info_t* info);
int tfft_mg_exec(tfft_mg_plan_t plan, const void* in_re, const void* in_im, void* out_re, void* out_im, void* stream);
int tfft_mg_exec_phase(tfft_mg_plan_t plan, int32_t phase, const void* in_re, const void* in_im, void* out_re,
                       void* out_im, void* stream);
int tfft_mg_status(tfft_mg_plan_t plan);
int tfft_mg_set_timeout_ms(tfft_mg_plan_t plan, int64_t milliseconds);
int tfft_mg_plan_destroy(tfft_mg_plan_t plan);

/* Device-side harness helpers.
 * tfft_fixture_sine: the reference's test signal (CreateSineSuperpostionKernel,
 * src/testing/TestingDataCreation.h:89-117) for `batch` transforms, written as planar fp16 at re/im + b*stride:
 * x_b[t] = sum_{i<cutoff} w_b[i] * sinf(2*pi*i*t/n); w_re / w_im are HOST arrays of batch*cutoff floats (the
 * reference draws them with GetRandomWeights, TestingDataCreation.h:15-27).  Synchronises `stream`.
 * tfft_error_stats: deviation statistics of a planar fp16 device result against planar fp64 device values over
 * the 2*count real and imaginary parts, as src/testing/AccuracyCalculator.h:86-148 computes them on the host:
 * out4 (HOST) = {largest deviation, average deviation, sigma of the deviation, relative L2 error}. */
int tfft_fixture_sine(void* re, void* im, int64_t n, int64_t batch, int64_t stride, const float* w_re,
                      const float* w_im, int32_t cutoff, void* stream);
int tfft_error_stats(const void* a_re, const void* a_im, const double* b_re, const double* b_im, int64_t count,
                     double* out4, void* stream);

const char* tfft_error_string(int code);
int tfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TFFT_H_ */
