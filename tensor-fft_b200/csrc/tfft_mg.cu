// Multi-GPU entry points of the C ABI (include/tfft.h, tfft_mg_*): ONE 1-D transform of length n = N1 * N2 slab-distributed
// over `world` GPUs of one node, one process (or thread) per GPU.  Six-step: transpose, N2/G transforms of length N1 fused
// with the twiddle exp(-2*pi*i*k1*i2/n), transpose, N1/G transforms of length N2, transpose (natural order out).
// The three exchanges are single kernels that store straight into the peers' buffers over NVLink (mg_kernels.cuh); the
// buffers are cudaMalloc'ed by the plan and shared through CUDA IPC handles that the caller passes between the ranks
// (any transport: torch.distributed, MPI, a file).  No NCCL on the data path.  The reference has no multi-GPU path
// (src/base/ComputeFFT.h:295-557 is commented out); SURVEY.md 8e / 8b name these entry points.
#include <cuda_runtime.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/tfft.h"
#include "mg_kernels.cuh"

using namespace tfft;

namespace {

constexpr uint64_t kBlobMagic = 0x74666674'6d673031ull;   // "tfftmg01"
struct Blob {   // what a rank publishes (<= TFFT_MG_HANDLE_BYTES)
  uint64_t magic;
  int64_t n;
  int32_t rank, world, pid, device;
  int32_t staged, reserved;   // exchange layout of the exporting rank: every rank of a transform must use the same one
  uint64_t base;    // device pointer in the exporting process (used directly by ranks living in the same process)
  uint64_t bytes;
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(Blob) <= TFFT_MG_HANDLE_BYTES, "handle blob too large");

int cuda_rc(cudaError_t e) {
  if (e == cudaSuccess) return TFFT_OK;
  cudaGetLastError();
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return TFFT_E_NO_DEVICE;
  if (e == cudaErrorMemoryAllocation) return TFFT_E_NOMEM;
  return static_cast<int>(e);
}

}  // namespace

struct tfft_mg_plan_s {
  int64_t n = 0, n1 = 0, n2 = 0, local = 0;   // local = n / world elements per rank and plane
  int lg = 0, rank = 0, world = 1, device = 0;
  tfft_plan_t fft1 = nullptr, fft2 = nullptr;
  char* base = nullptr;                        // [S0_re | S0_im | S1_re | S1_im | W_re | W_im | C_re | C_im | flags]
                                               // S0/S1: staging planes the peers store into (source-rank major); W: this
                                               // rank's working matrix (row major, transforms run in place); C: the result
  size_t bytes = 0, flags_off = 0;
  char* peer[kMgMaxRanks] = {};
  bool opened[kMgMaxRanks] = {};
  bool connected = false;
  uint32_t epoch = 0;
  int fuse1 = -1, fuse2 = -1;                  // staged plans: the transform gathers its input from the staging planes in its
                                               // TMA load (tfft_exec_segmented); -1 = not tried yet, 0 = unsupported shape
  bool staged = false;                         // exchanges go through source-rank-major staging planes + a local unpack (world > 2)
  uint32_t parity = 0;                         // execs alternate the staging planes: exchange e of an exec lands in S[(e + parity) & 1],
                                               // so that exchange 1 of the next exec never targets the planes a slower
                                               // peer is still unpacking from exchange 3 of this one
  int* status_host = nullptr;
  int* status_dev = nullptr;
  unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
  __half* plane(int r, int which) const {      // which: 0..7 as in the layout above, on rank r
    return reinterpret_cast<__half*>(peer[r] + static_cast<size_t>(which) * local * sizeof(__half));
  }
};

extern "C" {

int tfft_mg_plan_create(tfft_mg_plan_t* out, int64_t n, int32_t rank, int32_t world, uint32_t flags) {
  if (!out) return TFFT_E_INVALID_ARG;
  *out = nullptr;
  if (world < 1 || world > kMgMaxRanks || (world & (world - 1)) || rank < 0 || rank >= world || flags != 0) return TFFT_E_INVALID_ARG;
  int lg = 0;
  while ((int64_t(1) << lg) < n) ++lg;
  if (n <= 0 || (int64_t(1) << lg) != n || lg < 16 || lg > 30) return TFFT_E_INVALID_SIZE;
  const int lg1 = (lg + 1) / 2, lg2 = lg - lg1;
  const int64_t n1 = int64_t(1) << lg1, n2 = int64_t(1) << lg2;
  // every exchange moves 64 x 64 tiles between slabs: both factors must split into whole tiles per rank
  if (n1 / world < 64 || n2 / world < 64) return TFFT_E_INVALID_SIZE;
  tfft_mg_plan_s* p = new (std::nothrow) tfft_mg_plan_s;
  if (!p) return TFFT_E_NOMEM;
  p->n = n; p->n1 = n1; p->n2 = n2; p->lg = lg; p->rank = rank; p->world = world; p->local = n / world;
  int rc = cuda_rc(cudaGetDevice(&p->device));
  if (rc == TFFT_OK) rc = tfft_plan_create(&p->fft1, n1, n2 / world, 0);
  if (rc == TFFT_OK) rc = tfft_plan_create(&p->fft2, n2, n1 / world, 0);
  // everything lazy happens NOW: kernel loading can synchronise the device, which would deadlock against a peer's
  // barrier kernel that is already spinning on this GPU (several ranks per GPU) or serialise the ranks' first exec
  if (rc == TFFT_OK) rc = tfft_plan_prepare(p->fft1);
  if (rc == TFFT_OK) rc = tfft_plan_prepare(p->fft2);
  if (rc == TFFT_OK) {
    cudaFuncAttributes fa;
    rc = cuda_rc(cudaFuncGetAttributes(&fa, mg_transpose_send));
    if (rc == TFFT_OK) rc = cuda_rc(cudaFuncGetAttributes(&fa, mg_barrier));
    if (rc == TFFT_OK) rc = cuda_rc(cudaFuncGetAttributes(&fa, mg_unpack));
  }
  if (rc == TFFT_OK) {
    p->flags_off = 8 * static_cast<size_t>(p->local) * sizeof(__half);
    p->bytes = p->flags_off + 4096;
    rc = cuda_rc(cudaMalloc(&p->base, p->bytes));
  }
  if (rc == TFFT_OK) rc = cuda_rc(cudaMemset(p->base + p->flags_off, 0, 4096));
  if (rc == TFFT_OK) rc = cuda_rc(cudaHostAlloc(&p->status_host, sizeof(int), cudaHostAllocMapped));
  if (rc == TFFT_OK) {
    *p->status_host = 0;
    rc = cuda_rc(cudaHostGetDevicePointer(&p->status_dev, p->status_host, 0));
  }
  if (rc == TFFT_OK) rc = cuda_rc(cudaDeviceSynchronize());
  if (rc != TFFT_OK) {
    tfft_mg_plan_destroy(p);
    return rc;
  }
  p->peer[rank] = p->base;
  // direct row-major peer stores for 1-2 ranks (measured at 2 GPUs: 1.71 ms against 2.26 ms staged), staging above
  // (measured at 8 GPUs: the scattered stores cost 0.13 ms per exchange); TFFT_MG_STAGED=0/1 (developer) forces one
  p->staged = world > 2;
  if (const char* e = getenv("TFFT_DEVELOPER") ? getenv("TFFT_MG_STAGED") : nullptr) p->staged = atoi(e) != 0;
  if (world == 1) p->connected = true;
  *out = p;
  return TFFT_OK;
}

int tfft_mg_plan_handle(tfft_mg_plan_t p, void* handle) {
  if (!p || !handle) return TFFT_E_INVALID_ARG;
  Blob b;
  std::memset(&b, 0, sizeof(b));
  b.magic = kBlobMagic; b.n = p->n; b.rank = p->rank; b.world = p->world; b.pid = static_cast<int32_t>(getpid());
  b.device = p->device; b.base = reinterpret_cast<uint64_t>(p->base); b.bytes = p->bytes;
  b.staged = p->staged ? 1 : 0;
  const int rc = cuda_rc(cudaIpcGetMemHandle(&b.ipc, p->base));
  if (rc != TFFT_OK) return rc;
  std::memset(handle, 0, TFFT_MG_HANDLE_BYTES);
  std::memcpy(handle, &b, sizeof(b));
  return TFFT_OK;
}

int tfft_mg_plan_connect(tfft_mg_plan_t p, const void* handles) {
  if (!p || !handles) return TFFT_E_INVALID_ARG;
  if (p->connected) return TFFT_OK;
  const char* h = static_cast<const char*>(handles);
  for (int r = 0; r < p->world; ++r) {
    Blob b;
    std::memcpy(&b, h + static_cast<size_t>(r) * TFFT_MG_HANDLE_BYTES, sizeof(b));
    if (b.magic != kBlobMagic || b.rank != r || b.world != p->world || b.n != p->n || b.bytes != p->bytes ||
        b.staged != (p->staged ? 1 : 0))
      return TFFT_E_INVALID_ARG;
    if (r == p->rank) continue;
    if (b.pid == static_cast<int32_t>(getpid())) {
      // ranks that live in this process (one host thread per GPU, or several ranks on one GPU in the tests)
      if (b.device != p->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, p->device, b.device);
        if (!can) return TFFT_E_UNSUPPORTED;
        const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_rc(e);
        cudaGetLastError();
      }
      p->peer[r] = reinterpret_cast<char*>(b.base);
    } else {
      void* ptr = nullptr;
      const int rc = cuda_rc(cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess));
      if (rc != TFFT_OK) return rc;
      p->peer[r] = static_cast<char*>(ptr);
      p->opened[r] = true;
    }
  }
  p->connected = true;
  return TFFT_OK;
}

int tfft_mg_plan_info(tfft_mg_plan_t p, tfft_mg_info_t* info) {
  if (!p || !info) return TFFT_E_INVALID_ARG;
  std::memset(info, 0, sizeof(*info));
  info->n = p->n; info->n1 = p->n1; info->n2 = p->n2; info->local_elems = p->local;
  info->rank = p->rank; info->world = p->world; info->exchanges = 3;
  info->exchange_bytes_per_rank = 3 * static_cast<int64_t>(p->world - 1) * 4 * p->local / p->world;
  info->device_bytes = static_cast<int64_t>(p->bytes);
  info->result_re = p->base + 6 * static_cast<size_t>(p->local) * sizeof(__half);
  info->result_im = p->base + 7 * static_cast<size_t>(p->local) * sizeof(__half);
  return TFFT_OK;
}

// stage planes (source-rank major: [q][c][r]) -> row-major matrix [c][q*rows_local + r]; the slab that was sent had
// rows_local rows of cols columns on every rank, so this rank received cl = cols/world rows of world*rows_local elements
static int mg_unpack_stage(tfft_mg_plan_t p, int stage, int dst, int64_t rows_local, int64_t cols, cudaStream_t s) {
  const int64_t cl = cols / p->world;
  const dim3 grid(static_cast<unsigned>((cl + 7) / 8), static_cast<unsigned>(p->world), 2);
  mg_unpack<<<grid, 256, 0, s>>>(p->plane(p->rank, stage), p->plane(p->rank, stage + 1), p->plane(p->rank, dst),
                                 p->plane(p->rank, dst + 1), static_cast<int>(rows_local), static_cast<int>(cl), p->world);
  return cuda_rc(cudaGetLastError());
}

// One exchange.  Staged: transpose-send my slab into the peers' staging planes `stage` (0: S0, 2: S1), flag barrier, then
// -- if unpack_dst >= 0 -- unpack my own staging planes into the row-major matrix unpack_dst (4 = W, 6 = C).  Direct:
// transpose-send straight into the peers' row-major planes `stage` (0, 2 or 6), flag barrier.
static int mg_exchange(tfft_mg_plan_t p, const __half* src_re, const __half* src_im, int stage, int unpack_dst,
                       int64_t rows_local, int64_t cols, bool barrier, cudaStream_t s) {
  MgPeers peers;
  for (int r = 0; r < p->world; ++r) { peers.re[r] = p->plane(r, stage); peers.im[r] = p->plane(r, stage + 1); }
  const dim3 grid(static_cast<unsigned>(rows_local / 64), static_cast<unsigned>((cols / 32 + 7) / 8), 2);
  if (grid.y > 65535) return TFFT_E_UNSUPPORTED;
  mg_transpose_send<<<grid, 256, 0, s>>>(src_re, src_im, peers, static_cast<int>(rows_local), static_cast<int>(cols),
                                         p->rank, p->world, cols, p->staged ? 1 : 0);
  int rc = cuda_rc(cudaGetLastError());
  if (rc != TFFT_OK) return rc;
  if (barrier) {
    MgFlags f;
    for (int r = 0; r < p->world; ++r) f.flags[r] = reinterpret_cast<uint32_t*>(p->peer[r] + p->flags_off);
    mg_barrier<<<1, 32, 0, s>>>(f, p->rank, p->world, ++p->epoch, p->timeout_ns, p->status_dev);
    rc = cuda_rc(cudaGetLastError());
    if (rc != TFFT_OK) return rc;
    if (p->staged && unpack_dst >= 0) return mg_unpack_stage(p, stage, unpack_dst, rows_local, cols, s);
  }
  return TFFT_OK;
}

// The transforms of a staged plan: gather the input from the staging planes `stage` inside the TMA load when the shape
// allows (tfft_exec_segmented), else unpack into W first and transform in place.  rows_local / cols describe the slab the
// PEERS sent (this rank received cols/world transforms of world*rows_local elements).
static int mg_staged_fft(tfft_mg_plan_t p, tfft_plan_t fft, int* fuse, int stage, int64_t rows_local, int64_t cols, int64_t n,
                         int log2_total, int64_t first_col, cudaStream_t s) {
  __half *w_re = p->plane(p->rank, 4), *w_im = p->plane(p->rank, 5);
  const int64_t cl = cols / p->world;
  if (*fuse != 0) {
    const int rc = tfft_exec_segmented(fft, p->plane(p->rank, stage), p->plane(p->rank, stage + 1), w_re, w_im, rows_local, n,
                                       p->world, cl * rows_local, log2_total, first_col, s);
    if (rc != TFFT_E_UNSUPPORTED) { *fuse = 1; return rc; }
    *fuse = 0;
  }
  int rc = mg_unpack_stage(p, stage, 4, rows_local, cols, s);
  if (rc != TFFT_OK) return rc;
  return log2_total ? tfft_exec_twiddled(fft, w_re, w_im, w_re, w_im, n, n, log2_total, first_col, s)
                    : tfft_exec(fft, w_re, w_im, w_re, w_im, n, n, s);
}

// phase 0: exchange 1;  phase 1: transforms over i1 + exchange 2;  phase 2: transforms over i2 + exchange 3;  phase 3:
// result in place (+ copy out if asked).
// Staged plans (world > 2): the peers store into staging planes S[(e + parity) & 1] of exchange e; the transforms read them
// (gathering in their TMA load) and write the working matrix W, which the next exchange sends; exchange 3 is unpacked
// into the result C (after its barrier; for phase-stepped callers in phase 3).
// Direct plans (1-2 ranks): exchange 1 stores the row-major matrix into the peers' S0, transforms in place there, exchange 2
// into S1, exchange 3 into C.
static int mg_phase(tfft_mg_plan_t p, int phase, const void* in_re, const void* in_im, void* out_re, void* out_im,
                    bool barrier, cudaStream_t s) {
  const int64_t g = p->world, r1 = p->n1 / g, r2 = p->n2 / g;
  __half *c_re = p->plane(p->rank, 6), *c_im = p->plane(p->rank, 7);
  int rc = TFFT_OK;
  if (phase == 0) p->parity ^= 1u;
  const int sa = p->staged ? static_cast<int>(p->parity & 1u) * 2 : 0, sb = 2 - sa;
  switch (phase) {
    case 0:   // my n1/g rows of n2 -> every rank gets its n2/g columns, transposed
      return mg_exchange(p, static_cast<const __half*>(in_re), static_cast<const __half*>(in_im), sa, -1, r1, p->n2, barrier, s);
    case 1:   // n2/g transforms over i1, times exp(-2*pi*i*k1*i2/n); then M1[i2_local][k1] -> peers
      if (p->staged) {
        rc = mg_staged_fft(p, p->fft1, &p->fuse1, sa, r1, p->n2, p->n1, p->lg, p->rank * r2, s);
        if (rc == TFFT_OK) rc = mg_exchange(p, p->plane(p->rank, 4), p->plane(p->rank, 5), sb, -1, r2, p->n1, barrier, s);
      } else {
        rc = tfft_exec_twiddled(p->fft1, p->plane(p->rank, 0), p->plane(p->rank, 1), p->plane(p->rank, 0), p->plane(p->rank, 1),
                                p->n1, p->n1, p->lg, p->rank * r2, s);
        if (rc == TFFT_OK) rc = mg_exchange(p, p->plane(p->rank, 0), p->plane(p->rank, 1), 2, -1, r2, p->n1, barrier, s);
      }
      return rc;
    case 2:   // n1/g transforms over i2; then M2[k1_local][k2] -> C[k2_local][k1] = X[k1 + n1*k2]: this rank's n/g slice
      if (p->staged) {
        rc = mg_staged_fft(p, p->fft2, &p->fuse2, sb, r2, p->n1, p->n2, 0, 0, s);
        if (rc == TFFT_OK) rc = mg_exchange(p, p->plane(p->rank, 4), p->plane(p->rank, 5), sa, 6, r1, p->n2, barrier, s);
      } else {
        rc = tfft_exec(p->fft2, p->plane(p->rank, 2), p->plane(p->rank, 3), p->plane(p->rank, 2), p->plane(p->rank, 3), p->n2,
                       p->n2, s);
        if (rc == TFFT_OK) rc = mg_exchange(p, p->plane(p->rank, 2), p->plane(p->rank, 3), 6, -1, r1, p->n2, barrier, s);
      }
      return rc;
    case 3:
      if (p->staged && !barrier) rc = mg_unpack_stage(p, sa, 6, r1, p->n2, s);   // phase-stepped callers
      if (rc == TFFT_OK && out_re && out_re != c_re)
        rc = cuda_rc(cudaMemcpyAsync(out_re, c_re, p->local * sizeof(__half), cudaMemcpyDeviceToDevice, s));
      if (rc == TFFT_OK && out_im && out_im != c_im)
        rc = cuda_rc(cudaMemcpyAsync(out_im, c_im, p->local * sizeof(__half), cudaMemcpyDeviceToDevice, s));
      return rc;
    default:
      return TFFT_E_INVALID_ARG;
  }
}

int tfft_mg_exec(tfft_mg_plan_t p, const void* in_re, const void* in_im, void* out_re, void* out_im, void* stream_) {
  if (!p || !in_re || !in_im) return TFFT_E_INVALID_ARG;
  if (!p->connected) return TFFT_E_INVALID_ARG;
  if ((out_re == nullptr) != (out_im == nullptr)) return TFFT_E_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int rc = TFFT_OK;
  for (int phase = 0; phase < 4 && rc == TFFT_OK; ++phase) rc = mg_phase(p, phase, in_re, in_im, out_re, out_im, true, s);
  return rc;
}

int tfft_mg_exec_phase(tfft_mg_plan_t p, int32_t phase, const void* in_re, const void* in_im, void* out_re, void* out_im,
                       void* stream_) {
  if (!p || !p->connected || phase < 0 || phase > 3) return TFFT_E_INVALID_ARG;
  if (phase == 0 && (!in_re || !in_im)) return TFFT_E_INVALID_ARG;
  if ((out_re == nullptr) != (out_im == nullptr)) return TFFT_E_INVALID_ARG;
  return mg_phase(p, phase, in_re, in_im, out_re, out_im, false, static_cast<cudaStream_t>(stream_));
}

int tfft_mg_status(tfft_mg_plan_t p) {
  if (!p) return TFFT_E_INVALID_ARG;
  return *static_cast<volatile int*>(p->status_host) ? TFFT_E_TIMEOUT : TFFT_OK;
}

int tfft_mg_set_timeout_ms(tfft_mg_plan_t p, int64_t ms) {
  if (!p || ms < 1) return TFFT_E_INVALID_ARG;
  p->timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
  return TFFT_OK;
}

int tfft_mg_plan_destroy(tfft_mg_plan_t p) {
  if (!p) return TFFT_E_INVALID_ARG;
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
  if (p->fft1) tfft_plan_destroy(p->fft1);
  if (p->fft2) tfft_plan_destroy(p->fft2);
  if (p->base) cudaFree(p->base);
  if (p->status_host) cudaFreeHost(p->status_host);
  cudaGetLastError();
  delete p;
  return TFFT_OK;
}

}  // extern "C"
