// Host-side planning for the fused "unit" FFT kernel (pure C++17, no CUDA needed).
//
// A *unit* is the amount of work one CTA keeps resident in shared memory: E = U * L complex
// elements = U transforms of length L = T * 16^s (tail radix T in {1,2,4,8}, s in {2,3}
// tensor-core radix-16 stages).  The kernel runs, per unit:
//   load    : 16-byte global loads, radix-T butterfly + twiddle in fp32 registers, 16-byte
//             shared stores into the stage-1 operand layout
//   stage t : tcgen05.mma  D[128 x 32] = A_re * [Fr|Fi] + A_im * [-Fi|Fr]   (A = data, MN-major
//             SWIZZLE_NONE canonical layout, fp32 accumulators in tensor memory), then an
//             epilogue  tcgen05.ld -> twiddle (fp32) -> fp16 -> 16-byte shared stores into
//             the NEXT stage's operand layout (or the output staging layout)
//   store   : 16-byte shared loads, 8x8 in-register transpose, 16-byte global stores.
// Decimation in frequency, most significant input digit first:
//   n = j*(L/T) + r,  r = (n_1 n_2 .. n_s) base 16 (n_s least significant)
//   o = k_0 + T*(k_1 + 16*k_2 + 256*k_3)
// Every index map the kernel needs (item -> global offset, item -> shared offset, MMA row ->
// destination chunk, MMA row -> twiddle index ...) is *bit-linear*: a sum of per-bit
// contributions.  This file computes those contributions; the kernel just adds them up.
//
// Shared-memory operand layout of stage t (one plane = all real or all imaginary parts):
//   element (row, kappa) lives at   (row>>3)*kRowChunkStride + (kappa>>3)*128 + (kappa&7)*16
//   + (row&7)*2 bytes: 8 consecutive rows x 8 consecutive K values form a 128-byte core
//   matrix (UMMA SWIZZLE_NONE, MN-major: SBO = kRowChunkStride, LBO = 128; verified on a
//   B200 by probe/umma_probe.cu).  kRowChunkStride = 272 = 2 core matrices + 16 bytes of
//   padding, which makes every 16-byte store pattern below bank-conflict free.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace tfft {

constexpr int kMaxRowBits = 12;    // rows per unit = E/16 <= 4096
constexpr int kMaxItemBits = 13;   // load items per unit = E/(8T) <= 8192
constexpr int kMaxStages = 3;
constexpr uint32_t kRowChunkStride = 272;  // bytes between consecutive 8-row chunks (SBO)
constexpr uint32_t kKGroupStride = 128;    // bytes between the two 8-wide K groups (LBO)

enum AxisMode : uint32_t { kRowMode = 0, kColMode = 1 };

struct UnitShape {
  int log2_len = 0;     // log2(L)
  int log2_units = 0;   // log2(U): transforms per unit
  AxisMode in_mode = kRowMode;   // kRowMode: transform elements contiguous; kColMode: 8+ transforms interleaved
  AxisMode out_mode = kRowMode;
};

// Device-visible description of one kernel pass (passed by value as a kernel parameter).
struct UnitPlan {
  // ---- shape
  uint32_t log2_len, log2_units, log2_tail, stages;  // L, U, T, s
  uint32_t log2_elems;                               // log2(E)
  uint32_t in_mode, out_mode;
  uint32_t n_tiles;        // E / 16 / 128
  uint32_t plane_bytes;    // bytes of one operand plane (rows/8 * kRowChunkStride)
  uint32_t stage_plane_bytes;  // bytes of one staging plane
  // ---- load phase: item q (bit-linear) -> offsets
  uint32_t load_item_bits;
  uint32_t load_gofs[kMaxItemBits];   // global element offset contribution of item bit i
  uint32_t load_sofs[kMaxItemBits];   // shared byte offset contribution (k_0 = 0 chunk)
  uint32_t load_rval[kMaxItemBits];   // contribution to r (remaining index) of chunk element 0
  uint32_t load_estep;                // r increment between the 8 elements of a chunk (1 or 0)
  uint32_t load_gj;                   // global element stride of the tail digit j
  uint32_t load_sk0[8];               // shared byte offset of output k_0
  uint32_t load_uval[kMaxItemBits];   // contribution to the unit-local transform index u
  // ---- epilogues
  struct Epi {
    uint32_t dst[kMaxRowBits];   // byte contribution of row bit i to the destination chunk
    uint32_t aux[kMaxRowBits];   // contribution to the twiddle integer (m_t; o_row for the last stage)
    uint32_t col[kMaxRowBits];   // last stage: contribution to the unit-local transform index u
    uint32_t dst_khi;            // byte offset of the k in [8,16) chunk
    uint32_t tw_log2n;           // twiddle = exp(-2*pi*i * x / 2^tw_log2n)
    uint32_t tw_mode;            // 0: none; 1: x = aux*k; 2: x = (aux + k*tw_kw) * (col_base + col)
    uint32_t tw_kw;
  } epi[kMaxStages];
  // ---- store phase: item q (bit-linear) -> offsets
  uint32_t store_item_bits;
  uint32_t store_sofs[kMaxRowBits];   // staging byte offset contribution of item bit i
  uint32_t store_gofs[kMaxRowBits];   // global element offset contribution
  uint32_t store_xs[3];               // staging byte offsets of the 3 transposed bits
  uint32_t store_cg[3];               // global element offsets of chunk-internal bits k_s[0..2]
  uint32_t store_uval[kMaxRowBits];   // contribution to the unit-local transform index u
  uint32_t n_transforms;              // != 0 (row/row passes): transforms >= n_transforms are masked
  // ---- global addressing (elements)
  int64_t in_batch_stride, in_unit_stride;     // unit base = (unit / upb) * batch_stride + (unit % upb) * unit_stride
  int64_t out_batch_stride, out_unit_stride;
  uint32_t units_per_batch;
  uint32_t col_base_stride;   // tw_mode 2: col_base = ((unit % upb) / col_div) * col_base_stride
  uint32_t col_div;
  int64_t in_outer_stride, out_outer_stride;   // blockIdx.y level (2-D images)
  float load_scale;           // 1/T
};

// ------------------------------------------------------------------------------------------
// Logical bits of a unit: U(b) transform-in-unit, R(i) bit i of the remaining index r,
// K(t, i) bit i of output digit k_t (t = 0 is the tail digit).
struct LBit {
  enum Kind : uint8_t { U, R, K } kind;
  uint8_t stage;  // for K
  uint8_t idx;
  bool operator==(const LBit& o) const { return kind == o.kind && stage == o.stage && idx == o.idx; }
};

struct PlanBuildInfo {   // host-only by-products, used by the CPU simulator in tests/
  std::vector<LBit> row_bits[kMaxStages];   // row bit position -> logical bit, per MMA stage
  std::vector<LBit> load_bits;              // load item bit -> logical bit
  std::vector<LBit> stage_chunk_bits;       // staging dense chunk bit -> logical bit
  std::vector<LBit> store_bits;             // store item bit -> logical bit
  LBit store_x[3];
  std::string error;
};

namespace detail {

inline int find_bit(const std::vector<LBit>& v, const LBit& b) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == b) return static_cast<int>(i);
  return -1;
}

// byte contribution of a row at row-bit position p (p >= 3) of an operand plane
inline uint32_t row_pos_bytes(int p) { return (1u << (p - 3)) * kRowChunkStride; }
// byte contribution of bit p of the K index
inline uint32_t k_bit_bytes(int p) { return p < 3 ? (16u << p) : kKGroupStride; }
// 16-byte-chunk residue (mod 8) of the two
inline int row_pos_res(int p) { return p - 3 <= 2 ? (1 << (p - 3)) : 0; }
inline int k_bit_res(int p) { return p < 3 ? (1 << p) : 0; }

}  // namespace detail

// Builds the plan for one pass.  Addressing strides are filled in by the caller afterwards
// (fill_strides below) because they do not influence the layout decisions.
inline bool build_unit_plan(const UnitShape& shape, UnitPlan* plan, PlanBuildInfo* info) {
  using namespace detail;
  std::memset(plan, 0, sizeof(*plan));
  const int lg = shape.log2_len;
  const int tau = lg % 4;
  const int s = lg / 4;
  const int ups = shape.log2_units;
  const int eps = lg + ups;
  if (s < 2 || s > 3) { info->error = "length must be 2^8 .. 2^15"; return false; }
  if (eps < 11 || eps > 15) { info->error = "unit must hold 2^11 .. 2^15 elements"; return false; }
  if ((shape.in_mode == kColMode || shape.out_mode == kColMode) && ups < 3) {
    info->error = "column modes need >= 8 transforms per unit"; return false;
  }
  const int rbits = 4 * s;  // bits of r
  plan->log2_len = lg; plan->log2_units = ups; plan->log2_tail = tau; plan->stages = s;
  plan->log2_elems = eps; plan->in_mode = shape.in_mode; plan->out_mode = shape.out_mode;
  const int rowbits = eps - 4;
  plan->n_tiles = (1u << rowbits) / 128;
  plan->plane_bytes = (1u << (rowbits - 3)) * kRowChunkStride;
  plan->load_scale = 1.0f / static_cast<float>(1 << tau);

  auto is_kbit_of_stage = [&](const LBit& b, int t, int* p) {  // is b a bit of n_t (t = 1..s)?
    if (b.kind != LBit::R) return false;
    int lo = 4 * (s - t);
    if (b.idx >= lo && b.idx < lo + 4) { *p = b.idx - lo; return true; }
    return false;
  };

  // ---------------- row orders of the s operand layouts
  // in-chunk bits (row positions 0..2) are forced by whoever writes the layout with 16-byte
  // stores; positions 3..5 are chosen so that the writer's 8 quarter-warp lanes hit 8
  // different 16-byte bank groups.
  std::vector<LBit> writer_varying;  // logical bits that vary across the writer's quarter warp
  // load-phase item bit order (decides writer_varying for stage 1)
  std::vector<LBit>& lb = info->load_bits;
  lb.clear();
  if (shape.in_mode == kRowMode) {
    for (int i = 3; i < rbits; ++i) lb.push_back({LBit::R, 0, (uint8_t)i});
    for (int b = 0; b < ups; ++b) lb.push_back({LBit::U, 0, (uint8_t)b});
  } else {
    for (int b = 3; b < ups; ++b) lb.push_back({LBit::U, 0, (uint8_t)b});
    for (int i = 0; i < rbits; ++i) lb.push_back({LBit::R, 0, (uint8_t)i});
  }
  plan->load_item_bits = static_cast<uint32_t>(lb.size());
  writer_varying.assign(lb.begin(), lb.begin() + 3);

  for (int t = 1; t <= s; ++t) {
    std::vector<LBit>& rb = info->row_bits[t - 1];
    rb.clear();
    // forced in-chunk bits
    if (t == 1) {
      for (int i = 0; i < 3; ++i)
        rb.push_back(shape.in_mode == kRowMode ? LBit{LBit::R, 0, (uint8_t)i} : LBit{LBit::U, 0, (uint8_t)i});
    } else {
      for (int i = 0; i < 3; ++i) rb.push_back({LBit::K, (uint8_t)(t - 1), (uint8_t)i});
    }
    // all row bits of this stage
    std::vector<LBit> all;
    for (int i = 0; i < 4 * (s - t); ++i) all.push_back({LBit::R, 0, (uint8_t)i});
    for (int i = 0; i < tau; ++i) all.push_back({LBit::K, 0, (uint8_t)i});
    for (int tt = 1; tt < t; ++tt)
      for (int i = 0; i < 4; ++i) all.push_back({LBit::K, (uint8_t)tt, (uint8_t)i});
    for (int b = 0; b < ups; ++b) all.push_back({LBit::U, 0, (uint8_t)b});
    std::vector<LBit> rest;
    for (auto& b : all)
      if (find_bit(rb, b) < 0) rest.push_back(b);
    // residues (16-byte bank group mod 8) already taken by writer-varying bits that are K-line
    // bits of this stage; the other writer-varying bits must sit at row positions 3..5
    bool res_used[3] = {false, false, false};
    bool slot_has[3] = {false, false, false};
    LBit slot[3] = {};
    for (auto& v : writer_varying) {
      int p;
      if (is_kbit_of_stage(v, t, &p) && p < 3) res_used[p] = true;
    }
    for (auto& v : writer_varying) {
      int p;
      if (is_kbit_of_stage(v, t, &p) || find_bit(rb, v) >= 0) continue;
      p = 0;
      while (p < 3 && (res_used[p] || slot_has[p])) ++p;
      if (p == 3) { info->error = "no conflict-free slot"; return false; }
      slot_has[p] = true;
      slot[p] = v;
      rest.erase(rest.begin() + find_bit(rest, v));
    }
    for (int p = 0; p < 3; ++p) {
      if (slot_has[p]) { rb.push_back(slot[p]); continue; }
      if (rest.empty()) { info->error = "unit too small"; return false; }
      rb.push_back(rest.front());
      rest.erase(rest.begin());
    }
    for (auto& b : rest) rb.push_back(b);
    if ((int)rb.size() != rowbits) { info->error = "row bit count mismatch"; return false; }
    writer_varying.assign(rb.begin(), rb.begin() + 3);  // the epilogue of this stage writes the next layout
  }

  // ---------------- load maps
  {
    const std::vector<LBit>& rb = info->row_bits[0];
    auto smem_contrib = [&](const LBit& b) -> uint32_t {
      int p;
      if (is_kbit_of_stage(b, 1, &p)) return k_bit_bytes(p);
      int pos = find_bit(rb, b);
      return pos >= 3 ? row_pos_bytes(pos) : 0u;  // in-chunk bits are inside the 16-byte chunk
    };
    for (size_t i = 0; i < lb.size(); ++i) {
      plan->load_sofs[i] = smem_contrib(lb[i]);
      plan->load_rval[i] = lb[i].kind == LBit::R ? (1u << lb[i].idx) : 0u;
      plan->load_uval[i] = lb[i].kind == LBit::U ? (1u << lb[i].idx) : 0u;
      // gofs filled by fill_strides
    }
    plan->load_estep = shape.in_mode == kRowMode ? 1u : 0u;
    for (int k0 = 0; k0 < (1 << tau); ++k0) {
      uint32_t o = 0;
      for (int i = 0; i < tau; ++i)
        if (k0 >> i & 1) o += smem_contrib({LBit::K, 0, (uint8_t)i});
      plan->load_sk0[k0] = o;
    }
  }

  // ---------------- staging layout (output of the last epilogue)
  // chunk = 8 consecutive k_s[0..2]; chunk bits = row bits of stage s plus k_s[3].
  std::vector<LBit> obits;  // logical bits of the output index o, LSB first
  for (int i = 0; i < tau; ++i) obits.push_back({LBit::K, 0, (uint8_t)i});
  for (int t = 1; t <= s; ++t)
    for (int i = 0; i < 4; ++i) obits.push_back({LBit::K, (uint8_t)t, (uint8_t)i});
  std::vector<LBit> addr_bits;  // output address bits, fastest first (excluding nothing)
  if (shape.out_mode == kRowMode) {
    addr_bits = obits;
    for (int b = 0; b < ups; ++b) addr_bits.push_back({LBit::U, 0, (uint8_t)b});
  } else {
    for (int b = 0; b < ups; ++b) addr_bits.push_back({LBit::U, 0, (uint8_t)b});
    for (auto& b : obits) addr_bits.push_back(b);
  }
  for (int i = 0; i < 3; ++i) info->store_x[i] = addr_bits[i];
  std::vector<LBit> chunk_logical = info->row_bits[s - 1];
  chunk_logical.push_back({LBit::K, (uint8_t)s, 3});
  {
    std::vector<LBit>& sc = info->stage_chunk_bits;
    sc.clear();
    // S: the last epilogue's quarter-warp-varying bits
    for (int i = 0; i < 3; ++i) sc.push_back(info->row_bits[s - 1][i]);
    // G: the store phase's lane-varying bits = next output address bits (not X, not in-chunk k_s[0..2])
    std::vector<LBit> g;
    for (size_t i = 3; i < addr_bits.size() && g.size() < 3; ++i) {
      const LBit& b = addr_bits[i];
      if (b.kind == LBit::K && b.stage == s && b.idx < 3) continue;
      g.push_back(b);
    }
    // residues used by G bits already in S
    bool used[3] = {false, false, false};
    std::vector<LBit> gnew;
    for (auto& b : g) {
      int p = find_bit(sc, b);
      if (p >= 0) used[p] = true; else gnew.push_back(b);
    }
    // positions 3,4,5 have residues 1,2,4 (chunk address = d + (d >> 3)); place new G bits on free residues
    LBit mid[3]; bool mid_set[3] = {false, false, false};
    for (auto& b : gnew) {
      int p = 0;
      while (p < 3 && (used[p] || mid_set[p])) ++p;
      if (p == 3) { info->error = "no staging slot"; return false; }
      mid[p] = b; mid_set[p] = true;
    }
    std::vector<LBit> rest;
    for (auto& b : chunk_logical) {
      if (find_bit(sc, b) >= 0) continue;
      bool in_mid = false;
      for (int p = 0; p < 3; ++p) if (mid_set[p] && mid[p] == b) in_mid = true;
      if (!in_mid) rest.push_back(b);
    }
    for (int p = 0; p < 3; ++p) {
      if (mid_set[p]) { sc.push_back(mid[p]); continue; }
      sc.push_back(rest.front());
      rest.erase(rest.begin());
    }
    for (auto& b : rest) sc.push_back(b);
    const uint32_t nchunks = 1u << sc.size();
    plan->stage_plane_bytes = (nchunks + (nchunks >> 3)) * 16;
  }
  auto staging_contrib = [&](const LBit& b) -> uint32_t {
    int pd = find_bit(info->stage_chunk_bits, b);
    if (pd < 0) return 0;
    uint32_t c = (1u << pd) + (pd >= 3 ? (1u << (pd - 3)) : 0u);
    return c * 16;
  };

  // ---------------- epilogue maps
  for (int t = 1; t <= s; ++t) {
    UnitPlan::Epi& e = plan->epi[t - 1];
    const std::vector<LBit>& rb = info->row_bits[t - 1];
    for (int i = 0; i < rowbits; ++i) {
      const LBit& b = rb[i];
      if (t < s) {
        int p;
        if (is_kbit_of_stage(b, t + 1, &p)) e.dst[i] = k_bit_bytes(p);
        else e.dst[i] = row_pos_bytes(find_bit(info->row_bits[t], b));
        e.aux[i] = b.kind == LBit::R ? (1u << b.idx) : 0u;
      } else {
        e.dst[i] = staging_contrib(b);
        int po = find_bit(obits, b);
        e.aux[i] = po >= 0 ? (1u << po) : 0u;
        e.col[i] = b.kind == LBit::U ? (1u << b.idx) : 0u;
      }
    }
    if (t < s) {
      e.dst_khi = row_pos_bytes(find_bit(info->row_bits[t], {LBit::K, (uint8_t)t, 3}));
      e.tw_mode = 1;
      e.tw_log2n = 4 * (s - t + 1);
    } else {
      e.dst_khi = staging_contrib({LBit::K, (uint8_t)s, 3});
      e.tw_mode = 0;
      e.tw_log2n = 0;
      e.tw_kw = 1u << (tau + 4 * (s - 1));  // weight of k_s in o
    }
  }

  // ---------------- store maps
  {
    std::vector<LBit>& sb = info->store_bits;
    sb.clear();
    // item bits: output address order, skipping X (first 3 address bits) and in-chunk k_s[0..2]
    for (size_t i = 3; i < addr_bits.size(); ++i) {
      const LBit& b = addr_bits[i];
      if (b.kind == LBit::K && b.stage == s && b.idx < 3) continue;
      sb.push_back(b);
    }
    plan->store_item_bits = static_cast<uint32_t>(sb.size());
    for (size_t i = 0; i < sb.size(); ++i) {
      plan->store_sofs[i] = staging_contrib(sb[i]);
      plan->store_uval[i] = sb[i].kind == LBit::U ? (1u << sb[i].idx) : 0u;
    }
    for (int i = 0; i < 3; ++i) plan->store_xs[i] = staging_contrib(info->store_x[i]);
  }
  return true;
}

// Global addressing.  Row mode: element n of transform u of a unit at base + u*tstride + n.
// Column mode: base + u + n*nstride.  Unit base = (unit / upb)*batch_stride + (unit % upb)*unit_stride.
struct UnitStrides {
  int64_t in_tstride = 0, in_nstride = 1, out_tstride = 0, out_nstride = 1;
  int64_t in_batch_stride = 0, in_unit_stride = 0, out_batch_stride = 0, out_unit_stride = 0;
  uint32_t units_per_batch = 1;
  uint32_t col_base_stride = 0;
  uint32_t col_div = 1;
  bool col_from_u = true;    // tw_mode 2 column index includes the unit-local transform index u
  int64_t in_outer_stride = 0, out_outer_stride = 0;
  uint32_t n_transforms = 0;
  uint32_t pass1_log2n = 0;  // != 0: multiply outputs by exp(-2*pi*i*o*(col_base+u)/2^pass1_log2n)
};

inline void fill_strides(const UnitStrides& st, const PlanBuildInfo& info, UnitPlan* plan) {
  const int tau = plan->log2_tail, s = plan->stages;
  auto in_contrib = [&](const LBit& b) -> uint32_t {
    if (b.kind == LBit::U) return static_cast<uint32_t>((plan->in_mode == kRowMode ? st.in_tstride : 1) << b.idx);
    if (b.kind == LBit::R) return static_cast<uint32_t>((plan->in_mode == kRowMode ? 1 : st.in_nstride) << b.idx);
    return 0;
  };
  for (size_t i = 0; i < info.load_bits.size(); ++i) plan->load_gofs[i] = in_contrib(info.load_bits[i]);
  plan->load_gj = static_cast<uint32_t>((plan->in_mode == kRowMode ? 1 : st.in_nstride) << (4 * s));
  auto o_weight = [&](const LBit& b) -> int64_t {
    if (b.kind != LBit::K) return -1;
    return b.stage == 0 ? (int64_t(1) << b.idx) : (int64_t(1) << (tau + 4 * (b.stage - 1) + b.idx));
  };
  auto out_contrib = [&](const LBit& b) -> uint32_t {
    if (b.kind == LBit::U) return static_cast<uint32_t>((plan->out_mode == kRowMode ? st.out_tstride : 1) << b.idx);
    int64_t w = o_weight(b);
    return static_cast<uint32_t>(w * (plan->out_mode == kRowMode ? 1 : st.out_nstride));
  };
  for (size_t i = 0; i < info.store_bits.size(); ++i) plan->store_gofs[i] = out_contrib(info.store_bits[i]);
  for (int i = 0; i < 3; ++i) plan->store_cg[i] = out_contrib({LBit::K, (uint8_t)s, (uint8_t)i});
  plan->in_batch_stride = st.in_batch_stride; plan->in_unit_stride = st.in_unit_stride;
  plan->out_batch_stride = st.out_batch_stride; plan->out_unit_stride = st.out_unit_stride;
  plan->units_per_batch = st.units_per_batch;
  plan->col_base_stride = st.col_base_stride;
  plan->col_div = st.col_div ? st.col_div : 1;
  plan->in_outer_stride = st.in_outer_stride; plan->out_outer_stride = st.out_outer_stride;
  plan->n_transforms = st.n_transforms;
  if (!st.col_from_u)
    for (int i = 0; i < kMaxRowBits; ++i) plan->epi[s - 1].col[i] = 0;
  if (st.pass1_log2n) {
    plan->epi[s - 1].tw_mode = 2;
    plan->epi[s - 1].tw_log2n = st.pass1_log2n;
  }
}

}  // namespace tfft
