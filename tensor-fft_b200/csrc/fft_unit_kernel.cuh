// Fused shared-memory-resident FFT kernel for sm_100a.  One CTA transforms one *unit*
// (U transforms of length L = T*16^s, E = U*L <= 32768 complex elements) and touches HBM
// exactly once: planar fp16 in, planar fp16 out.
//
//   load    16-byte LDG -> radix-T butterfly + twiddle in fp32 registers -> 16-byte STS into
//           the stage-1 tensor-core operand layout (replaces the reference's digit-reversal
//           gather src/base/TensorFFT256.cu:125-171 and its Radix2Kernel launches
//           src/base/Radix2.cu:20-77, src/base/ComputeFFT.h:123-145)
//   stage   radix-16 DFT as tcgen05.mma (M=128 rows of data x N=32 [re|im] x K=16), fp32
//           accumulation in tensor memory; epilogue tcgen05.ld -> fp32 twiddle -> fp16 ->
//           16-byte STS into the next stage's operand layout (replaces the wmma stages of
//           src/base/TensorFFT256.cu:191-275 and src/base/TensorRadix16.cu:101-213, which
//           accumulate in fp16 and pay one HBM round trip per radix-16 step)
//   store   16-byte LDS -> 8x8 in-register transpose -> 16-byte coalesced STG.
// All index maps come from the host-side UnitPlan (unit_plan.h).  Scaling: 1/T in the load
// phase and 1/16 folded into the fp16 DFT matrix (exact powers of two), total 1/L like the
// reference's "sequential scaling" (TensorFFT256.cu:167-171, TensorRadix16.cu:133-136).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "sm100_ptx.cuh"
#include "unit_plan.h"

namespace tfft {

constexpr int kThreads = 256;
constexpr uint32_t kBMatBytes = 1024;   // one 16 x 32 fp16 B operand

struct Cplx {
  float re, im;
};
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) {
  return {fmaf(a.re, b.re, -a.im * b.im), fmaf(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ Cplx cadd(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cmul_mi(Cplx a) { return {a.im, -a.re}; }  // a * (-i)
// exp(-2*pi*i * x / 2^log2n) for an integer phase x (already reduced mod 2^log2n or small)
__device__ __forceinline__ Cplx twiddle(uint32_t x, uint32_t log2n) {
  float s, c;
  // sincospif is exact in its range reduction; -2x/2^log2n is an exact fp32 product
  sincospif(-static_cast<float>(x) * __uint_as_float((128u - log2n) << 23), &s, &c);  // 2^(1-log2n)
  return {c, s};
}

// in-register forward DFT of size T (T = 2, 4, 8), unscaled
template <int T>
__device__ __forceinline__ void small_dft(Cplx (&x)[T]) {
  if constexpr (T == 2) {
    Cplx a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  } else if constexpr (T == 4) {
    Cplx a = cadd(x[0], x[2]), b = csub(x[0], x[2]), c = cadd(x[1], x[3]), d = cmul_mi(csub(x[1], x[3]));
    x[0] = cadd(a, c);
    x[1] = cadd(b, d);
    x[2] = csub(a, c);
    x[3] = csub(b, d);
  } else if constexpr (T == 8) {
    Cplx e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
    small_dft<4>(e);
    small_dft<4>(o);
    const float h = 0.70710678118654752f;
    Cplx o1 = {h * (o[1].re + o[1].im), h * (o[1].im - o[1].re)};    // * exp(-i*pi/4)
    Cplx o2 = cmul_mi(o[2]);                                          // * (-i)
    Cplx o3 = {h * (o[3].im - o[3].re), -h * (o[3].re + o[3].im)};   // * exp(-3i*pi/4)
    x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
    x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
    x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
    x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
  }
}

__device__ __forceinline__ uint32_t bit_sum(uint32_t q, const uint32_t* contrib, int first, int count) {
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < count; ++i)
    if ((q >> (first + i)) & 1u) s += contrib[first + i];
  return s;
}

__device__ __forceinline__ uint4 ldg128(const __half* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg128(__half* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}

// Dynamic shared memory carve-up (bytes): [plane_re | plane_im | B1 | B2 | mbar | tmem slot]
struct SmemLayout {
  uint32_t plane_stride, b1_off, b2_off, bar_off, slot_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(const UnitPlan& p) {
  SmemLayout l;
  uint32_t pb = p.plane_bytes > p.stage_plane_bytes ? p.plane_bytes : p.stage_plane_bytes;
  l.plane_stride = (pb + 127u) & ~127u;
  l.b1_off = 2 * l.plane_stride;
  l.b2_off = l.b1_off + kBMatBytes;
  l.bar_off = l.b2_off + kBMatBytes;
  l.slot_off = l.bar_off + 8;
  l.total = l.slot_off + 8;
  return l;
}
__host__ __device__ inline uint32_t tmem_cols(const UnitPlan& p) {
  uint32_t need = p.n_tiles * 32, c = 32;
  while (c < need) c <<= 1;
  return c;
}

template <int LOG2T>
__global__ void __launch_bounds__(kThreads) fft_unit_kernel(const __grid_constant__ UnitPlan P,
                                                            const __half* __restrict__ in_re,
                                                            const __half* __restrict__ in_im,
                                                            __half* __restrict__ out_re,
                                                            __half* __restrict__ out_im) {
  using namespace ptx;
  constexpr int T = 1 << LOG2T;
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout SL = smem_layout(P);
  uint8_t* plane_re = smem;
  uint8_t* plane_im = smem + SL.plane_stride;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SL.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SL.slot_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t unit = blockIdx.x;
  const uint32_t ub = unit / P.units_per_batch, uu = unit % P.units_per_batch;
  const int64_t in_base = static_cast<int64_t>(ub) * P.in_batch_stride + static_cast<int64_t>(uu) * P.in_unit_stride +
                          static_cast<int64_t>(blockIdx.y) * P.in_outer_stride;
  const int64_t out_base = static_cast<int64_t>(ub) * P.out_batch_stride +
                           static_cast<int64_t>(uu) * P.out_unit_stride +
                           static_cast<int64_t>(blockIdx.y) * P.out_outer_stride;
  // row/row passes with a ragged batch: transforms past the end are loaded as zeros, never stored
  const uint32_t u_limit = P.n_transforms ? P.n_transforms - min(P.n_transforms, unit << P.log2_units) : 0xFFFFFFFFu;
  const uint32_t ncols = tmem_cols(P);

  // ------------------------------------------------------------------ setup
  if (warp == 0) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  {
    // B operands, K-major SWIZZLE_NONE: Bmath[k][n] at (n>>3)*256 + (k>>3)*128 + (n&7)*16 + (k&7)*2
    // B1 = [Fr | Fi] / 16, B2 = [-Fi | Fr] / 16, F[k][n] = exp(-2*pi*i*k*n/16)
    __half* b1 = reinterpret_cast<__half*>(smem + SL.b1_off);
    __half* b2 = reinterpret_cast<__half*>(smem + SL.b2_off);
    for (int e = tid; e < 512; e += kThreads) {
      int k = e >> 5, n = e & 31, nn = n & 15;
      float s, c;
      sincospif(-static_cast<float>((k * nn) & 15) * 0.125f, &s, &c);
      float fr = c * 0.0625f, fi = s * 0.0625f;
      uint32_t off = (n >> 3) * 128 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);  // in halves
      b1[off] = __float2half_rn(n < 16 ? fr : fi);
      b2[off] = __float2half_rn(n < 16 ? -fi : fr);
    }
  }

  // ------------------------------------------------------------------ load phase
  {
    const uint32_t n_items = 1u << P.load_item_bits;
    const __half* gre = in_re + in_base;
    const __half* gim = in_im + in_base;
    Cplx delta = {1.f, 0.f};
    if (T > 1 && P.load_estep) delta = twiddle(1u, P.log2_len);
    for (uint32_t q = tid; q < n_items; q += kThreads) {
      const uint32_t g = bit_sum(q, P.load_gofs, 0, kMaxItemBits);
      const uint32_t so = bit_sum(q, P.load_sofs, 0, kMaxItemBits);
      const bool live = bit_sum(q, P.load_uval, 0, kMaxItemBits) < u_limit;
      if constexpr (T == 1) {
        uint4 vr = make_uint4(0, 0, 0, 0), vi = vr;
        if (live) {
          vr = ldg128(gre + g);
          vi = ldg128(gim + g);
        }
        *reinterpret_cast<uint4*>(plane_re + so) = vr;
        *reinterpret_cast<uint4*>(plane_im + so) = vi;
      } else {
        const uint32_t r0 = bit_sum(q, P.load_rval, 0, kMaxItemBits);
        uint4 vr[T], vi[T];
#pragma unroll
        for (int j = 0; j < T; ++j) {
          vr[j] = vi[j] = make_uint4(0, 0, 0, 0);
          if (live) {
            vr[j] = ldg128(gre + g + j * P.load_gj);
            vi[j] = ldg128(gim + g + j * P.load_gj);
          }
        }
        uint32_t ore[T][4], oim[T][4];
        Cplx w1 = twiddle(r0, P.log2_len);  // exp(-2*pi*i*r/L) of chunk element 0
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {      // two chunk elements per iteration (one packed register)
          float yre[T][2], yim[T][2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            Cplx x[T];
#pragma unroll
            for (int j = 0; j < T; ++j) {
              const uint32_t wr = reinterpret_cast<const uint32_t*>(&vr[j])[e2];
              const uint32_t wi = reinterpret_cast<const uint32_t*>(&vi[j])[e2];
              float2 fr = unpack_half2(wr), fi = unpack_half2(wi);
              x[j] = {h ? fr.y : fr.x, h ? fi.y : fi.x};
            }
            small_dft<T>(x);
            Cplx wk = {P.load_scale, 0.f};
            const Cplx w1s = w1;
#pragma unroll
            for (int k0 = 0; k0 < T; ++k0) {
              Cplx y = cmul(x[k0], wk);
              yre[k0][h] = y.re;
              yim[k0][h] = y.im;
              wk = cmul(wk, w1s);
            }
            w1 = cmul(w1, delta);
          }
#pragma unroll
          for (int k0 = 0; k0 < T; ++k0) {
            ore[k0][e2] = pack_half2(yre[k0][0], yre[k0][1]);
            oim[k0][e2] = pack_half2(yim[k0][0], yim[k0][1]);
          }
        }
#pragma unroll
        for (int k0 = 0; k0 < T; ++k0) {
          const uint32_t o = so + P.load_sk0[k0];
          *reinterpret_cast<uint4*>(plane_re + o) = make_uint4(ore[k0][0], ore[k0][1], ore[k0][2], ore[k0][3]);
          *reinterpret_cast<uint4*>(plane_im + o) = make_uint4(oim[k0][0], oim[k0][1], oim[k0][2], oim[k0][3]);
        }
      }
    }
  }

  // ------------------------------------------------------------------ radix-16 stages
  const uint32_t lane_row = static_cast<uint32_t>((warp & 3) * 32 + lane);
  uint32_t taddr = 0;
  for (uint32_t t = 0; t < P.stages; ++t) {
    fence_proxy_async_smem();   // generic-proxy operand stores -> visible to the tensor core
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    taddr = *tmem_slot;
    if (tid == 0) {
      const uint32_t idesc = make_idesc_f16(128, 32, /*a_mn=*/1, /*b_mn=*/0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t db1 = make_smem_desc(sbase + SL.b1_off, 128, 256);
      const uint64_t db2 = make_smem_desc(sbase + SL.b2_off, 128, 256);
      for (uint32_t tile = 0; tile < P.n_tiles; ++tile) {
        const uint32_t a_off = tile * 16 * kRowChunkStride;
        const uint64_t da_re = make_smem_desc(sbase + a_off, kKGroupStride, kRowChunkStride);
        const uint64_t da_im = make_smem_desc(sbase + SL.plane_stride + a_off, kKGroupStride, kRowChunkStride);
        umma_f16_ss(taddr + tile * 32, da_re, db1, idesc, 0);
        umma_f16_ss(taddr + tile * 32, da_im, db2, idesc, 1);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, t & 1);
    tc_fence_after_sync();

    const UnitPlan::Epi& E = P.epi[t];
    const uint32_t dst_lo = bit_sum(lane_row, E.dst, 0, 7);
    const uint32_t aux_lo = bit_sum(lane_row, E.aux, 0, 7);
    const uint32_t col_lo = bit_sum(lane_row, E.col, 0, 7);
    for (uint32_t tile = warp >> 2; tile < P.n_tiles; tile += kThreads / 128) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(taddr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + tile * 32, acc);
      const uint32_t dst = dst_lo + bit_sum(tile, E.dst + 7, 0, kMaxRowBits - 7);
      const uint32_t aux = aux_lo + bit_sum(tile, E.aux + 7, 0, kMaxRowBits - 7);
      Cplx tw = {1.f, 0.f}, step = {1.f, 0.f};
      if (E.tw_mode == 1) {
        step = twiddle(aux, E.tw_log2n);
      } else if (E.tw_mode == 2) {
        const uint32_t col = col_lo + bit_sum(tile, E.col + 7, 0, kMaxRowBits - 7) + (uu / P.col_div) * P.col_base_stride;
        const uint32_t mask = (1u << E.tw_log2n) - 1u;
        tw = twiddle((aux * col) & mask, E.tw_log2n);
        step = twiddle((E.tw_kw * col) & mask, E.tw_log2n);
      }
      tmem_ld_wait();
      uint32_t pre[8], pim[8];
      if (E.tw_mode == 0) {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          pre[k >> 1] = pack_half2(__uint_as_float(acc[k]), __uint_as_float(acc[k + 1]));
          pim[k >> 1] = pack_half2(__uint_as_float(acc[16 + k]), __uint_as_float(acc[17 + k]));
        }
      } else {
        float vre[16], vim[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          Cplx y = cmul({__uint_as_float(acc[k]), __uint_as_float(acc[16 + k])}, tw);
          vre[k] = y.re;
          vim[k] = y.im;
          tw = cmul(tw, step);
        }
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          pre[k >> 1] = pack_half2(vre[k], vre[k + 1]);
          pim[k >> 1] = pack_half2(vim[k], vim[k + 1]);
        }
      }
      *reinterpret_cast<uint4*>(plane_re + dst) = make_uint4(pre[0], pre[1], pre[2], pre[3]);
      *reinterpret_cast<uint4*>(plane_re + dst + E.dst_khi) = make_uint4(pre[4], pre[5], pre[6], pre[7]);
      *reinterpret_cast<uint4*>(plane_im + dst) = make_uint4(pim[0], pim[1], pim[2], pim[3]);
      *reinterpret_cast<uint4*>(plane_im + dst + E.dst_khi) = make_uint4(pim[4], pim[5], pim[6], pim[7]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();

  // ------------------------------------------------------------------ store phase
  {
    const uint32_t n_items = 1u << P.store_item_bits;
    __half* gre = out_re + out_base;
    __half* gim = out_im + out_base;
    for (uint32_t q = tid; q < n_items; q += kThreads) {
      const uint32_t so = bit_sum(q, P.store_sofs, 0, kMaxRowBits);
      const uint32_t g = bit_sum(q, P.store_gofs, 0, kMaxRowBits);
      if (bit_sum(q, P.store_uval, 0, kMaxRowBits) >= u_limit) continue;
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        const uint8_t* sp = plane ? plane_im : plane_re;
        __half* gp = plane ? gim : gre;
        uint4 a[8];
#pragma unroll
        for (int x = 0; x < 8; ++x)
          a[x] = *reinterpret_cast<const uint4*>(sp + so + bit_sum(static_cast<uint32_t>(x), P.store_xs, 0, 3));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t sel = (c & 1) ? 0x7632u : 0x5410u;
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t lo = reinterpret_cast<const uint32_t*>(&a[2 * i])[c >> 1];
            const uint32_t hi = reinterpret_cast<const uint32_t*>(&a[2 * i + 1])[c >> 1];
            w[i] = __byte_perm(lo, hi, sel);
          }
          stg128(gp + g + bit_sum(static_cast<uint32_t>(c), P.store_cg, 0, 3), make_uint4(w[0], w[1], w[2], w[3]));
        }
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  if (warp == 0) tmem_dealloc(taddr, ncols);
}

}  // namespace tfft
