// TEST INFRASTRUCTURE ONLY -- CPU interpreter of a UnitPlan (tensor-fft_b200/csrc/unit_plan.h).
// Walks the exact index maps the CUDA kernel uses (load items, operand layouts, MMA tiles /
// TMEM lanes, epilogue destinations, staging, store items) in double precision, optionally
// rounding to fp16 where the kernel does, and counts shared-memory bank conflicts of every
// 16-byte access pattern.  Lets `pytest -m "not gpu"` prove the plan algebra == DFT.
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../../tensor-fft_b200/csrc/unit_plan.h"

using namespace tfft;
typedef std::complex<double> cd;
static const double kPi = 3.14159265358979323846264338327950288;
static inline double rh(double v, bool on) { return on ? (double)(_Float16)v : v; }
static inline cd twd(int64_t e, int64_t n) {
  e %= n; if (e < 0) e += n;
  double a = -2.0 * kPi * (double)e / (double)n;
  return cd(std::cos(a), std::sin(a));
}
static uint32_t bitsum(uint32_t q, const uint32_t* c, int nbits) {
  uint32_t s = 0;
  for (int i = 0; i < nbits; ++i) if (q >> i & 1) s += c[i];
  return s;
}
// Excess wavefronts of one warp-wide LDS.64 (32 lanes, 8 bytes each, byte offsets `off` into the table): the access is
// served in as many wavefronts as the busiest of the 32 four-byte banks has DISTINCT words; ideal = ceil(distinct words / 32)
#include <set>
#include <map>
static int lds64_excess(const uint32_t* off) {
  std::map<int, std::set<uint32_t>> banks;
  std::set<uint32_t> words;
  for (int l = 0; l < 32; ++l)
    for (int hw = 0; hw < 2; ++hw) {
      const uint32_t w = (off[l] >> 2) + hw;
      banks[w & 31].insert(w);
      words.insert(w);
    }
  size_t worst = 0;
  for (auto& b : banks) worst = std::max(worst, b.second.size());
  return (int)worst - (int)((words.size() + 31) / 32);
}
// conflicts of one quarter-warp of 16-byte accesses: max lanes per 16-byte bank group - 1
static int qw_conflict(const uint32_t* addr) {
  int cnt[8] = {0};
  int worst = 0;
  for (int i = 0; i < 8; ++i) { int g = (addr[i] >> 4) & 7; if (++cnt[g] > worst) worst = cnt[g]; }
  return worst - 1;
}

extern "C" {

// Runs `n_units` units.  in/out are planar double arrays (element offsets as the plan says).
// Returns the total number of bank conflicts found (negative = plan error).
// conflicts[0..3] = load stores, epilogue stores, store-phase loads, (unused)
// ext3 (may be null): {kron_bits, kron_log2n, col_base_stride, out_hi_from, out_hi_stride} for Kronecker (2-D row pass)
// units; the last two (0 = off) describe a tiled row-mode output
int plansim_run_ex(int log2_len, int log2_units, int in_mode_flags, int out_mode, const int64_t* strides9,
                   int pass1_log2n, int n_units, const double* in_re, const double* in_im,
                   double* out_re, double* out_im, int emulate_fp16, int* conflicts, const int64_t* ext3) {
  UnitShape shape; shape.log2_len = log2_len; shape.log2_units = log2_units;
  if (ext3) shape.kron_bits = (int)ext3[0];
  const int in_mode = in_mode_flags & 1;
  shape.in_mode = (AxisMode)in_mode; shape.out_mode = (AxisMode)out_mode; shape.tma_load = (in_mode_flags & 2) != 0;
  shape.pipe_stage2 = (in_mode_flags & 4) != 0;
  shape.cluster = (in_mode_flags & 8) != 0;   // CTA-pair unit: both ranks are simulated, stage-1 stores cross between them
  shape.no_col64 = (in_mode_flags & 32) != 0; // column tiles: 16-column SWIZZLE_32B tiles even for >= 64 columns
  shape.ring = (in_mode_flags & 16) != 0;     // landing-ring unit: stage-1 operand outside the planes, dense staging for 8-column output
  UnitPlan P; PlanBuildInfo info;
  if (!build_unit_plan(shape, &P, &info)) { fprintf(stderr, "plan error: %s\n", info.error.c_str()); return -1; }
  UnitStrides st;
  st.in_tstride = strides9[0]; st.in_nstride = strides9[1]; st.out_tstride = strides9[2]; st.out_nstride = strides9[3];
  st.in_batch_stride = strides9[4]; st.in_unit_stride = strides9[5]; st.out_batch_stride = strides9[6];
  st.out_unit_stride = strides9[7]; st.units_per_batch = (uint32_t)strides9[8]; st.col_base_stride = 1u << log2_units;
  st.n_units = (uint32_t)n_units;
  st.pass1_log2n = pass1_log2n;
  if (ext3) { st.kron_log2n = (uint32_t)ext3[1]; st.col_base_stride = (uint32_t)ext3[2]; st.out_hi_from = (int)ext3[3]; st.out_hi_stride = ext3[4]; }
  fill_strides(st, info, &P);
  const bool h = emulate_fp16 != 0;
  const int s = P.stages;
  const int64_t L = int64_t(1) << P.log2_len;
  conflicts[0] = conflicts[1] = conflicts[2] = conflicts[3] = 0;
  int lookups = 0;   // warp-wide LDS.64 table lookups modelled (for reference; conflicts[3] is their excess wavefronts)
  (void)lookups;
  // ring units keep the stage-1 operand in the landing ring (a unit = 64 KiB per plane), not in the planes
  const uint32_t smem_halves = (P.ring ? (P.plane_bytes > 65536u ? P.plane_bytes : 65536u) : P.plane_bytes) / 2;
  if (P.ring) {   // the kernel's carve-up (smem_ring_layout): planes + 64 KiB ring + tables + barriers within 227 KiB
    uint32_t tables = P.stages == 3 ? 0u : 4608u;
    for (uint32_t t = 0; t < P.stages; ++t) {
      bool seen = false;
      for (uint32_t u = 0; u < t; ++u) seen = seen || P.log2_radix[u] == P.log2_radix[t];
      if (!seen) tables += 6u << (2 * P.log2_radix[t]);
    }
    if (2u * ((P.plane_bytes + 1023u) & ~1023u) + 65536u + tables + 128u > 227u * 1024u) { fprintf(stderr, "ring unit does not fit shared memory\n"); return -8; }
  }
  for (int unit = 0; unit < n_units; ++unit) {
    const int64_t ibase = (unit / P.units_per_batch) * P.in_batch_stride + (unit % P.units_per_batch) * P.in_unit_stride;
    const int64_t obase = (unit / P.units_per_batch) * P.out_batch_stride + (unit % P.units_per_batch) * P.out_unit_stride;
    const uint32_t col_base = ((unit % P.units_per_batch) / P.col_div) * P.col_base_stride;
    const int nranks = P.cluster ? 2 : 1;
    std::vector<double> pre_[2], pim_[2];   // operand planes of the (up to two) CTAs
    for (int rk = 0; rk < nranks; ++rk) { pre_[rk].assign(smem_halves, NAN); pim_[rk].assign(smem_halves, NAN); }
    // ---------------- load (per CTA): TMA tensor tile ... or 16-byte copies
    for (int rk = 0; rk < nranks; ++rk) {
    std::vector<double>& sre = pre_[rk];
    std::vector<double>& sim = pim_[rk];
    const int R0 = 1 << P.log2_radix[0];
    const int64_t Mfull = L / R0, Mloc = Mfull >> (P.cluster ? 1 : 0), m0 = rk * Mloc;   // a cluster CTA loads one half of m
    if (P.tma_load == 6) {   // column mode, 32-column tiles as SWIZZLE_64B atoms: row = (u&31) + 32*(m + M*(u>>5))
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t u = 0; u < U; ++u)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap) {
            const int64_t row = (u & 31) + 32 * (m + Mloc * (u >> 5));
            uint32_t off = (uint32_t)((row >> 5) * 64 * R + kap * 64 + (row & 31) * 2);
            off ^= ((off >> 7) & 3u) << 4;
            const int64_t a = ibase + u + (kap * Mfull + m0 + m) * strides9[1];
            sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
          }
    } else if (P.tma_load == 5) {   // column mode, 64-column tiles as SWIZZLE_128B atoms: row = (u&63) + 64*(m + M*(u>>6))
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t u = 0; u < U; ++u)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap) {
            const int64_t row = (u & 63) + 64 * (m + Mloc * (u >> 6));
            const uint32_t off = (uint32_t)((row >> 6) * 128 * R + (kap >> 3) * 1024 + (kap & 7) * 128 +
                                            ((((row >> 3) & 7) ^ (kap & 7)) << 4) + (row & 7) * 2);
            const int64_t a = ibase + u + (kap * Mfull + m0 + m) * strides9[1];
            sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
          }
    } else if (P.tma_load == 4) {   // column mode, 16-column tiles as SWIZZLE_32B atoms: row = (u&15) + 16*(m + M*(u>>4))
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t u = 0; u < U; ++u)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap) {
            const int64_t row = (u & 15) + 16 * (m + Mloc * (u >> 4));
            uint32_t off = (uint32_t)((row >> 4) * 32 * R + kap * 32 + (row & 15) * 2);
            off ^= ((off >> 7) & 1u) << 4;
            const int64_t a = ibase + u + (kap * Mfull + m0 + m) * strides9[1];
            sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
          }
    } else if (P.tma_load == 2) {   // column mode: tiles {8 columns, R kappa, M rows} per 8-column group, dense, no swizzle
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t ug = 0; ug < U / 8; ++ug)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap)
            for (int cc = 0; cc < 8; ++cc) {
              const uint32_t off = (uint32_t)((((ug * Mloc + m) * R + kap) * 8 + cc) * 2);
              const int64_t a = ibase + ug * 8 + cc + (kap * Mfull + m0 + m) * strides9[1];
              sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
            }
    } else if (P.tma_load == 3) {   // SWIZZLE_32B atoms of 16 rows: dense [atom][kappa][16 rows], byte bit 4 ^= bit 7
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t u = 0; u < U; ++u)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap) {
            const int64_t row = u * Mloc + m;
            uint32_t off = (uint32_t)((row >> 4) * 32 * R + kap * 32 + (row & 15) * 2);
            off ^= ((off >> 7) & 1u) << 4;
            const int64_t a = ibase + u * strides9[0] + kap * Mfull + m0 + m;
            sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
          }
    } else if (P.tma_load) {
      const int R = R0;
      const int64_t U = int64_t(1) << P.log2_units;
      for (int64_t u = 0; u < U; ++u)
        for (int64_t m = 0; m < Mloc; ++m)
          for (int kap = 0; kap < R; ++kap) {
            const int64_t row = u * Mloc + m;
            const uint32_t off = (uint32_t)((row >> 6) * 128 * R + (kap >> 3) * 1024 + (kap & 7) * 128 +
                                            ((((row >> 3) & 7) ^ (kap & 7)) << 4) + (row & 7) * 2);
            const int64_t a = ibase + u * strides9[0] + kap * Mfull + m0 + m;
            sre[off / 2] = rh(in_re[a], h); sim[off / 2] = rh(in_im[a], h);
          }
    }
    // ... or a pure copy of 16-byte chunks (cp.async)
    const uint32_t n_items = P.tma_load ? 0u : (1u << P.load_item_bits);
    for (uint32_t q0 = 0; q0 < n_items; q0 += 8) {
      uint32_t addr[8];
      for (uint32_t dq = 0; dq < 8; ++dq) {
        uint32_t q = q0 + dq;
        uint32_t g = bitsum(q, P.load_gofs, P.load_item_bits) + rk * P.cl_load_gofs;
        uint32_t so = bitsum(q, P.load_sofs, P.load_item_bits);
        addr[dq] = so;
        if (so + 16 > P.plane_bytes) { fprintf(stderr, "load dst out of range\n"); return -2; }
        for (int e = 0; e < 8; ++e) {
          if (!std::isnan(sre[so / 2 + e])) { fprintf(stderr, "load: destination written twice\n"); return -3; }
          sre[so / 2 + e] = rh(in_re[ibase + g + e], h);
          sim[so / 2 + e] = rh(in_im[ibase + g + e], h);
        }
      }
      conflicts[0] += qw_conflict(addr);
    }
    }   // rank (load)
    // ---------------- MMA stages
    for (int t = 1; t <= s; ++t) {
      const UnitPlan::Epi& E = P.epi[t - 1];
      const int rho = P.log2_radix[t - 1], R = 1 << rho;
      const int rowbits = P.log2_elems - rho;
      const uint32_t S = P.chunk_stride[t - 1];
      std::vector<double> nre_[2], nim_[2];
      for (int rk = 0; rk < nranks; ++rk) { nre_[rk].assign(smem_halves, NAN); nim_[rk].assign(smem_halves, NAN); }
      const uint32_t rows = 1u << rowbits;
      if (rows != P.n_tiles[t - 1] * 128) { fprintf(stderr, "tile count mismatch\n"); return -5; }
      // Twiddle-table lookups of the epilogue (2-stage plans only; 3-stage plans keep their seeds in registers): per
      // item a warp of 32 consecutive rows reads TWlo[x & 63], TWhi[x >> 6] for x = aux << tw_shift and, for radix 32 / 64,
      // for x * 16 g.  conflicts[3] accumulates the EXCESS wavefronts of these LDS.64 (0 = conflict free).
      if (t < s && s == 2 && E.tw_mode == 1 && unit == 0) {
        for (uint32_t row0 = 0; row0 < rows; row0 += 32)
          for (int gq = 0; gq < R / 16; ++gq) {   // one item = (32 rows of a tile, column group g)
            uint32_t lo[32], hi[32], lo2[32], hi2[32];
            for (int l = 0; l < 32; ++l) {
              const uint32_t x = bitsum(row0 + l, E.aux, rowbits) << E.tw_shift;        // seed w1
              const uint32_t x2 = (x * 16u * gq) & ((1u << P.log2_len) - 1u);           // seed t0 = w1^(16 g)
              lo[l] = 8u * (x & 63u);
              hi[l] = 8u * (64u + (x >> 6));
              lo2[l] = 8u * (x2 & 63u);
              hi2[l] = 8u * (64u + (x2 >> 6));
            }
            conflicts[3] += lds64_excess(lo) + lds64_excess(hi);
            lookups += 2;
            if (R > 16) { conflicts[3] += lds64_excess(lo2) + lds64_excess(hi2); lookups += 2; }
          }
      }
      for (int rk = 0; rk < nranks; ++rk) {
      const std::vector<double>& sre = pre_[rk];
      const std::vector<double>& sim = pim_[rk];
      for (uint32_t row0 = 0; row0 < rows; row0 += 8) {
        uint32_t addr[8];
        for (uint32_t dr = 0; dr < 8; ++dr) {
          uint32_t row = row0 + dr;
          // A operand read exactly as the UMMA descriptors address it
          std::vector<cd> a(R), y(R);
          for (int kap = 0; kap < R; ++kap) {
            uint32_t off = (row >> 3) * S + (kap >> 3) * kKGroupStride + (kap & 7) * 16 + (row & 7) * 2;
            if (t == 1 && (P.tma_load == 3 || P.tma_load == 4)) {
              off = (row >> 4) * 32 * R + kap * 32 + (row & 15) * 2;
              off ^= ((off >> 7) & 1u) << 4;
            }
            if (t == 1 && P.tma_load == 6) {
              off = (row >> 5) * 64 * R + kap * 64 + (row & 31) * 2;
              off ^= ((off >> 7) & 3u) << 4;
            }
            if (t == 1 && (P.tma_load == 1 || P.tma_load == 5))
              off = (row >> 6) * 128 * R + (kap >> 3) * 1024 + (kap & 7) * 128 + ((((row >> 3) & 7) ^ (kap & 7)) << 4) + (row & 7) * 2;
            a[kap] = cd(sre[off / 2], sim[off / 2]);
            if (std::isnan(sre[off / 2])) { fprintf(stderr, "stage %d reads an unwritten operand slot\n", t); return -6; }
          }
          for (int k = 0; k < R; ++k) {
            cd acc = 0;
            for (int kap = 0; kap < R; ++kap) {
              cd f = twd((int64_t)kap * k, R) / (double)R;
              if (P.kron_bits && t == s) {   // F_x (x) F_y: K index = (kappa_x, kappa_y), column = (k_x, k_y)
                const int Rx = R >> P.kron_bits, Ry = 1 << P.kron_bits;
                f = twd((int64_t)(kap % Rx) * (k % Rx), Rx) * twd((int64_t)(kap / Rx) * (k / Rx), Ry) / (double)R;
              }
              if (h) f = cd(rh(f.real(), true), rh(f.imag(), true));
              acc += a[kap] * f;
            }
            y[k] = acc;
          }
          uint32_t dst = bitsum(row, E.dst, rowbits);
          uint32_t aux = bitsum(row, E.aux, rowbits);
          uint32_t col = bitsum(row, E.col, rowbits);
          if (P.cluster && rk) {   // the CTA's rank is an index bit (input half in stage 1, k_1 half afterwards)
            if (t == 1) { dst += P.cl_in_dst; aux += P.cl_in_aux; }
            else if (t == s) aux += P.cl_out_aux;
          }
          addr[dr] = dst;
          // MMA / epilogue overlap of stage 2: first-half rows may only write into the consumed first half
          if (P.pipe_stage2 && t == 2 && row < rows / 2) {
            const uint32_t half_bytes = (rows / 16) * S;
            uint32_t hi = dst;
            for (int j = 0; j < 3; ++j) hi += E.dst_k[j];
            if (hi + 16 > half_bytes) { fprintf(stderr, "pipeline hazard: first-half epilogue writes beyond the consumed half\n"); return -7; }
          }
          for (int k = 0; k < R; ++k) {
            cd v = y[k];
            if (E.tw_mode == 1) v *= twd(((int64_t)aux << E.tw_shift) * k, L);
            if (E.tw_mode == 2) v *= twd(((int64_t)aux + (int64_t)k * E.tw_kw) * (int64_t)(col_base + col), int64_t(1) << E.tw_log2n);
            if (E.tw_mode == 3) v *= twd((int64_t)(k >> E.tw_shift) * (int64_t)col_base, int64_t(1) << E.tw_log2n);
            uint32_t o = dst + bitsum((uint32_t)(k >> 3), E.dst_k, 3);
            if (o + 16 > P.plane_bytes) { fprintf(stderr, "dst out of range\n"); return -2; }
            o = o / 2 + (k & 7);
            // cluster stage 1: the top k_1 bit picks the CTA that receives the chunk
            const int tgt = (P.cluster && t == 1) ? (int)(((uint32_t)(k >> 3) >> (rho - 4)) & 1u) : rk;
            if (!std::isnan(nre_[tgt][o])) { fprintf(stderr, "stage %d: destination written twice\n", t); return -3; }
            nre_[tgt][o] = rh(v.real(), h); nim_[tgt][o] = rh(v.imag(), h);
          }
        }
        conflicts[1] += qw_conflict(addr);
      }
      }   // rank (stage)
      for (int rk = 0; rk < nranks; ++rk) { pre_[rk].swap(nre_[rk]); pim_[rk].swap(nim_[rk]); }
    }
    // ---------------- store
    for (int rk = 0; rk < nranks; ++rk) {
    const std::vector<double>& sre = pre_[rk];
    const std::vector<double>& sim = pim_[rk];
    const uint32_t n_sitems = 1u << P.store_item_bits;
    for (uint32_t q0 = 0; q0 < n_sitems; q0 += 8) {
      for (int x = 0; x < 8; ++x) {
        uint32_t addr[8];
        for (uint32_t dq = 0; dq < 8; ++dq) {
          uint32_t q = q0 + dq;
          uint32_t so = bitsum(q, P.store_sofs, P.store_item_bits) + bitsum(x, P.store_xs, 3);
          uint32_t g = bitsum(q, P.store_gofs, P.store_item_bits) + rk * P.cl_out_gofs;
          addr[dq] = so;
          for (int c = 0; c < 8; ++c) {
            int64_t a = obase + g + x + bitsum(c, P.store_cg, 3);
            double vr = sre[so / 2 + c], vi = sim[so / 2 + c];
            if (std::isnan(vr)) { fprintf(stderr, "store reads an unwritten staging slot\n"); return -4; }
            out_re[a] = vr; out_im[a] = vi;
          }
        }
        conflicts[2] += qw_conflict(addr);
      }
    }
    }   // rank (store)
  }
  return conflicts[0] + conflicts[1] + conflicts[2];
}

int plansim_run(int log2_len, int log2_units, int in_mode_flags, int out_mode, const int64_t* strides9,
                int pass1_log2n, int n_units, const double* in_re, const double* in_im,
                double* out_re, double* out_im, int emulate_fp16, int* conflicts) {
  return plansim_run_ex(log2_len, log2_units, in_mode_flags, out_mode, strides9, pass1_log2n, n_units, in_re, in_im,
                        out_re, out_im, emulate_fp16, conflicts, nullptr);
}

int plansim_plan_bytes() { return (int)sizeof(UnitPlan); }

}  // extern "C"
