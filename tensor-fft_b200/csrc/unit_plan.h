// Host-side planning for the fused "unit" FFT kernel (pure C++17, no CUDA needed).
//
// A *unit* is the amount of work one CTA keeps resident in shared memory: E = U * L complex
// elements = U transforms of length L = R_1 * ... * R_s, s in {2,3} tensor-core stages of
// radix R_t in {16, 32, 64}.  Per unit the kernel runs
//   load    : 16-byte asynchronous global->shared copies straight into the stage-1 tensor-core
//             operand layout (no arithmetic, no register staging)
//   stage t : radix-R_t DFT as tcgen05.mma  D[128 x 2R] = A_re * [Fr|Fi] + A_im * [-Fi|Fr]
//             (A = 128 rows of data, MN-major SWIZZLE_NONE canonical layout; fp32 accumulators
//             in tensor memory), then an epilogue  tcgen05.ld -> twiddle (packed fp32) ->
//             fp16 -> 16-byte shared stores into the NEXT stage's operand layout (or the
//             output staging layout after the last stage)
//   store   : 16-byte shared loads, 8x8 in-register transpose, 16-byte coalesced global stores.
// Decimation in frequency, most significant input digit first:
//   n = (n_1 n_2 .. n_s) mixed radix, n_1 most significant;  o = k_1 + R_1*(k_2 + R_2*k_3)
//   after stage t the data is multiplied by exp(-2*pi*i * k_t * m_t / N_t), N_t = R_t*..*R_s,
//   m_t = the not yet transformed low digits.
// Every index map the kernel needs (item -> global offset, item -> shared offset, MMA row ->
// destination chunk, MMA row -> twiddle index ...) is *bit-linear*: a sum of per-bit
// contributions.  This file computes those contributions; the kernel just adds them up.
//
// Shared-memory operand layout of stage t (one plane = all real or all imaginary parts):
//   element (row, kappa) lives at (row>>3)*S_t + (kappa>>3)*128 + (kappa&7)*16 + (row&7)*2 bytes:
//   8 consecutive rows x 8 consecutive K values form a 128-byte core matrix (UMMA SWIZZLE_NONE,
//   MN-major: SBO = S_t, LBO = 128; verified on a B200 by probe/umma_probe.cu).
//   S_t = 16*R_t + 16: R_t/8 core matrices plus 16 bytes of padding, which makes every 16-byte
//   store pattern below bank-conflict free (consecutive row chunks land 16 bytes apart mod 128).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace tfft {

constexpr int kMaxRowBits = 12;    // rows per unit = E/16 <= 4096 (bit-linear maps of an MMA row)
constexpr int kMaxItemBits = 12;   // 16-byte chunks per unit and plane = E/8 <= 4096
constexpr int kMaxStages = 3;
constexpr int kMaxTiles = 32;     // 128-row tiles per stage = E / R / 128 <= 16 for the shapes built here
constexpr uint32_t kKGroupStride = 128;    // bytes between consecutive 8-wide K groups (LBO)
constexpr int kTwLoBits = 6;               // two-level twiddle table: phase = hi * 64 + lo

enum AxisMode : uint32_t { kRowMode = 0, kColMode = 1 };

struct UnitShape {
  int log2_len = 0;     // log2(L)
  int log2_units = 0;   // log2(U): transforms per unit
  AxisMode in_mode = kRowMode;   // kRowMode: transform elements contiguous; kColMode: 8+ transforms interleaved
  AxisMode out_mode = kRowMode;
  bool tma_load = false;   // stage-1 operand written by a TMA tensor load (SWIZZLE_128B, natural row order)
  int kron_bits = 0;         // 2-D row pass: the U = 2^kron_bits transforms of a unit are rows y_lo + u*(ny/U) of one image
                             // and the LAST tensor stage also transforms across them (DFT matrix = F_x (x) F_y), i.e.
                             // the unit computes a 2-D DFT of size U x L.  Needs kron_bits == log2_units, row modes.
  bool cluster = false;      // the unit holds 2^16 elements and is shared by a pair of CTAs (thread-block cluster of 2): CTA r
                             // loads the half of the rows whose highest not-yet-transformed index bit (the top bit of m_1)
                             // is r, runs stage 1 on it, and its stage-1 epilogue stores every output whose top k_1 bit is d
                             // into CTA d's shared memory (distributed shared memory); stages 2.. and the store are local
  bool no_col64 = false;     // column-mode TMA tiles: never the 64-column SWIZZLE_128B tiles (tma_load 5), developer A/B
  bool ring = false;         // 32K-element units with TMA input: the stage-1 operand lands in a separate ring of two quarter
                             // tiles (64 KiB) instead of the working planes, so that the loads and the stage-1 MMAs of
                             // unit q+1 run under the store phase of unit q (fft_unit_kernel_ring)
  bool pipe_stage2 = false;  // 3-stage plans: make the top row bit of stages 2 and 3 the same logical bit (k_1's
                             // top bit), so that the epilogue of the first half of stage 2's tiles only writes into
                             // the already consumed first half of the operand planes (MMA / epilogue overlap)
};

// Device-visible description of one kernel pass (passed by value as a kernel parameter).
struct UnitPlan {
  // ---- shape
  uint32_t log2_len, log2_units, stages;
  uint32_t log2_elems;                     // log2(E)
  uint32_t in_mode, out_mode;
  uint32_t log2_radix[kMaxStages];         // rho_t
  uint32_t n_tiles[kMaxStages];            // E / R_t / 128
  uint32_t chunk_stride[kMaxStages];       // S_t
  uint32_t plane_bytes;                    // bytes of one operand plane, max over stages and staging
  uint32_t tmem_cols;                      // power of two >= E/64
  uint32_t pipe_stage2;                    // 1: stage 2 may overlap its second-half MMAs with its first-half epilogue
  uint32_t kron_bits;                      // see UnitShape::kron_bits (0: plain 1-D units)
  uint32_t il_swap;                        // interleaved + inverse: the pairs are read / written as (im, re)
  uint32_t il_in, il_out;                  // TFFT_INTERLEAVED: this pass reads / writes half2 (re, im) elements (cp.async
                                           // load path only); the element offsets of the plan are doubled on the fly
  uint32_t cluster;                        // 1: CTA-pair unit (UnitShape::cluster); the fields below are the rank-1 constants
  uint32_t cl_in_dst, cl_in_aux;           //   stage-1 epilogue: destination bytes / twiddle integer of the input half bit
  uint32_t cl_out_aux;                     //   last stage: output-index weight of the top k_1 bit (tw_mode 2)
  uint32_t cl_load_c2;                     //   TMA loads: dim-2 tile coordinate of rank 1 (row tiles: M/128; column tiles: M/2)
  uint32_t cl_load_gofs;                   //   cp.async loads: global element offset of rank 1's rows (fill_strides)
  uint32_t cl_out_gofs;                    //   store: global element offset of rank 1's outputs (fill_strides)
  uint32_t tma_seg;                        // 1 (row tiles of 64-row atoms only): every input transform is split into equal
                                           // segments lying segment_stride apart (multi-GPU staging planes, source-rank
                                           // major); the tensor map is 5-D {64, kappa_lo, segment, M/64, transform} and the
                                           // tile coordinates are (0, 0, 0, c2, c3)
  uint32_t ring;                           // 1: landing-ring unit (UnitShape::ring): the unit is loaded as four parts (quarter
  uint32_t ring_c2_step, ring_c3_step;     //   of the stage-1 tiles each); part p starts at tile coordinates (c2, c3) + p * step
  uint32_t prefetch_next;                  // 1: pull the next unit's input into L2 during this unit's stages
  uint32_t tma_load;                       // 6: column-mode input, 32 columns per unit: tiles {32 columns, R kappa, M rows} as
                                           //    SWIZZLE_64B atoms of 32 rows (whole 64-byte pieces): row = (u&31) + 32*m, element
                                           //    (row, kappa) at (row>>5)*64R + kappa*64 + (row&31)*2, byte-address bits 4-5 ^= bits 7-8
                                           // 5: column-mode input, >= 64 columns per unit: tiles {64 columns, R kappa, M rows} per
                                           //    64-column group as SWIZZLE_128B atoms (whole 128-byte lines): row = (u&63) +
                                           //    64*(m + M*(u>>6)), element (row, kappa) as in mode 1
                                           // 4: column-mode input, >= 16 columns per unit: tiles {16 columns, R kappa, M rows} per
                                           //    16-column group as SWIZZLE_32B atoms: row = (u&15) + 16*(m + M*(u>>4)), element
                                           //    (row, kappa) as in mode 3
                                           // 3: row-mode input with 16 or 32 rows per K line: SWIZZLE_32B MN-major atoms of
                                           //    16 rows filled by TMA: element (row, kappa) at (row>>4)*32R + kappa*32 +
                                           //    (row&15)*2, byte-address bit 4 ^= bit 7
                                           // 2: column-mode input, stage-1 operand filled by TMA tiles {8 columns, R kappa,
                                           //    M rows} per 8-column group, no swizzle: dense chunks [group][m][kappa][8 cols],
                                           //    i.e. chunk_stride[0] = 16R without padding and natural row order (u&7, m, u>>3)
                                           // 1: stage-1 operand = SWIZZLE_128B MN-major atoms of 64 rows filled by TMA:
                                           //    element (row, kappa) at (row>>6)*128R + (kappa>>3)*1024 + (kappa&7)*128
                                           //    + ((((row>>3)&7) ^ (kappa&7))<<4) + (row&7)*2   (verified by probe/tma_probe.cu)
  // ---- load phase: chunk q (bit-linear) -> offsets
  uint32_t load_item_bits;
  uint32_t load_gofs[kMaxItemBits];   // global element offset contribution of item bit i
  uint32_t load_sofs[kMaxItemBits];   // shared byte offset contribution
  uint32_t load_uval[kMaxItemBits];   // contribution to the unit-local transform index u
  // ---- epilogues
  struct Epi {
    uint32_t dst[kMaxRowBits];   // byte contribution of row bit i to the destination chunk
    uint32_t aux[kMaxRowBits];   // contribution to the twiddle integer (m_t; o_row for the last stage)
    uint32_t col[kMaxRowBits];   // last stage: contribution to the unit-local transform index u
    uint32_t dst_k[3];           // byte offsets of the chunk-index bits k_t[3], k_t[4], k_t[5]
    uint32_t tw_mode;            // 0: none; 1: x = (aux << tw_shift) * k, unit angle 2*pi/L;
                                 // 2: x = (aux + k*tw_kw) * (col_base + col), unit angle 2*pi/2^tw_log2n
                                 // 3: x = (k >> tw_shift) * col_base, unit angle 2*pi/2^tw_log2n (Kronecker stage)
    uint32_t tw_shift;
    uint32_t tw_log2n;
    uint32_t tw_kw;
    // tw_mode 1: the twiddle integer of a row splits into a per-thread part (lane and warp-group bits, turned into a
    // complex seed once per kernel) and the part of the 128-row tile, which is the same for the whole warp and comes from
    // here (kernel-parameter space): tile_tw[t] = exp(-2*pi*i * x_t / L), tile_tw16[t] = exp(-2*pi*i * 16*x_t / L),
    // x_t = (sum of aux[] over the tile bits of t) << tw_shift
    float tile_tw[kMaxTiles][2];
    float tile_tw16[kMaxTiles][2];
  } epi[kMaxStages];
  // ---- store phase: item q (bit-linear) -> offsets
  uint32_t store_item_bits;
  uint32_t store_sofs[kMaxItemBits];   // staging byte offset contribution of item bit i
  uint32_t store_gofs[kMaxItemBits];   // global element offset contribution
  uint32_t store_uval[kMaxItemBits];   // contribution to the unit-local transform index u
  uint32_t store_xs[3];                // staging byte offsets of the 3 transposed bits
  uint32_t store_cg[3];                // global element offsets of chunk-internal bits k_s[0..2]
  // ---- global addressing (elements)
  int64_t in_batch_stride, in_unit_stride;     // unit base = (unit / upb) * batch_stride + (unit % upb) * unit_stride
  int64_t out_batch_stride, out_unit_stride;
  uint32_t units_per_batch;                    // always a power of two (or the sentinel 0x7FFFFFFF = "one batch")
  uint32_t upb_shift;                          // log2(units_per_batch) (31 for the sentinel): the kernel shifts and masks
  uint32_t col_shift;                          // log2(col_div)
  uint32_t tma_batch_step;                     // TMA loads: outermost tile coordinate = (unit >> upb_shift) * tma_batch_step
                                               //   + ((unit & mask) << log2_units)   (0: the unit index alone, 1-D batches)
  uint32_t n_units;                            // total units of the launch (persistent CTAs loop over them)
  uint32_t n_transforms;                       // != 0 (row/row passes): transforms >= n_transforms are masked
  // outer batch level (three-pass plans, whose own "batch" level is an index of the transform): unit = b3 << b3_shift | rest,
  // rest decomposes as above; b3_shift = 31: none
  uint32_t b3_shift;
  uint32_t tma_b3_step;                        // column tiles: 1: batch coordinate = unit / upb + b3 (4-D map); 2: 5-D map,
                                               // coordinates (unit / upb, b3).  Row tiles (tma_row5): transform
                                               // coordinate = (unit / upb) * U + b3 * tma_b3_step
  uint32_t tma_row5;                           // 1: row tiles through a 5-D map {rows, kappa, row atoms, unit % upb, transform}
                                               // (last pass of a three-pass plan: the unit's transforms lie a whole row of
                                               // the N1 x N2 matrix apart, unit % upb selects the column block)
  int64_t in_b3_stride, out_b3_stride;
  uint32_t col_base_stride;   // tw_mode 2: col_base = ((unit % upb) / col_div) * col_base_stride
  uint32_t col_div;
  uint32_t col_first;         // tw_mode 2: added to every column index (tfft_exec_twiddled)
};

// ------------------------------------------------------------------------------------------
// Logical bits of a unit: U(b) transform-in-unit, R(i) bit i of the input index n (within a
// transform), K(t, i) bit i of output digit k_t (t = 1..s).
struct LBit {
  enum Kind : uint8_t { U, R, K } kind;
  uint8_t stage;  // for K
  uint8_t idx;
  bool operator==(const LBit& o) const { return kind == o.kind && stage == o.stage && idx == o.idx; }
};

struct PlanBuildInfo {   // host-only by-products, used by fill_strides and the CPU simulator in tests/
  std::vector<LBit> row_bits[kMaxStages];   // row bit position -> logical bit, per MMA stage
  std::vector<LBit> load_bits;              // load item bit -> logical bit
  std::vector<LBit> stage_chunk_bits;       // staging dense chunk bit -> logical bit
  std::vector<LBit> store_bits;             // store item bit -> logical bit
  LBit store_x[3];
  int rho[kMaxStages] = {0, 0, 0};
  int rx[kMaxStages] = {0, 0, 0};           // bits of the transform index each stage consumes (== rho except for
                                            // a Kronecker last stage: rho - kron_bits)
  int kron_bits = 0;
  int lo_bit[kMaxStages] = {0, 0, 0};       // n_t = n bits [lo_bit, lo_bit + rho)
  bool cluster = false;
  LBit cin = {LBit::R, 0, 0}, cout = {LBit::K, 1, 0};   // cluster units: the input half bit and the output half bit
  std::string error;
};

// radix schedule: fewest stages, largest radix last (the last stage has no twiddle)
inline int radix_schedule(int log2_len, int* rho, int kron_bits = 0) {
  if (kron_bits == 1 && log2_len == 14) {   // 2 rows x 8192: the Kronecker stage as the small radix, so that stages 1 and
    rho[0] = 5; rho[1] = 5; rho[2] = 4;     // 2 share one DFT matrix and the two-slot kernel fits in shared memory
    return 3;
  }
  switch (log2_len) {
    case 8: rho[0] = 4; rho[1] = 4; return 2;
    case 9: rho[0] = 4; rho[1] = 5; return 2;
    case 10: rho[0] = 5; rho[1] = 5; return 2;
    case 11: rho[0] = 5; rho[1] = 6; return 2;
    case 12: rho[0] = 6; rho[1] = 6; return 2;
    case 13: rho[0] = 4; rho[1] = 4; rho[2] = 5; return 3;
    case 14: rho[0] = 4; rho[1] = 5; rho[2] = 5; return 3;
    case 15: rho[0] = 5; rho[1] = 5; rho[2] = 5; return 3;
    case 16: rho[0] = 4; rho[1] = 6; rho[2] = 6; return 3;   // cluster units only (one CTA pair per transform)
    default: return 0;
  }
}

namespace detail {

inline int find_bit(const std::vector<LBit>& v, const LBit& b) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == b) return static_cast<int>(i);
  return -1;
}
// byte contribution of bit p of the K index
inline uint32_t k_bit_bytes(int p) { return p < 3 ? (16u << p) : (kKGroupStride << (p - 3)); }

}  // namespace detail

inline bool build_unit_plan(const UnitShape& shape, UnitPlan* plan, PlanBuildInfo* info) {
  using namespace detail;
  std::memset(plan, 0, sizeof(*plan));
  const int lg = shape.log2_len;
  const int ups = shape.log2_units;
  const bool cl = shape.cluster;
  const int eps = lg + ups - (cl ? 1 : 0);   // elements per CTA
  int* rho = info->rho;
  const int kb = shape.kron_bits;
  if (cl && (kb || lg + ups != 16 || shape.in_mode != shape.out_mode)) {
    info->error = "cluster units hold 2^16 elements, same mode in and out, no Kronecker stage"; return false;
  }
  if (!cl && lg > 15) { info->error = "length must be 2^8 .. 2^15"; return false; }
  info->cluster = cl;
  plan->cluster = cl ? 1u : 0u;
  if (kb && (kb != ups || shape.in_mode != kRowMode || shape.out_mode != kRowMode || lg < 8)) {
    info->error = "Kronecker units need kron_bits == log2_units and row modes"; return false;
  }
  const int s = radix_schedule(kb ? eps : lg, rho, kb);
  if (s == 0) { info->error = "length must be 2^8 .. 2^15 (2^16 for cluster units)"; return false; }
  if (eps < 13 || eps > 15) { info->error = "unit must hold 2^13 .. 2^15 elements"; return false; }
  int* rx = info->rx;
  for (int t = 0; t < s; ++t) rx[t] = rho[t] - (t == s - 1 ? kb : 0);
  if (rx[s - 1] < 3) { info->error = "Kronecker stage keeps fewer than 3 transform bits"; return false; }
  info->kron_bits = kb;
  plan->kron_bits = static_cast<uint32_t>(kb);
  if ((shape.in_mode == kColMode || shape.out_mode == kColMode) && ups < 3) {
    info->error = "column modes need >= 8 transforms per unit"; return false;
  }
  plan->log2_len = lg; plan->log2_units = ups; plan->stages = s;
  plan->log2_elems = eps; plan->in_mode = shape.in_mode; plan->out_mode = shape.out_mode;
  // row-mode TMA tiles: SWIZZLE_128B atoms of 64 consecutive rows when a K line has M = L/R_1 >= 64 contiguous rows
  // (tma_load 1), SWIZZLE_32B atoms of 16 rows for M = 16 / 32 (tma_load 3; a 32-byte inner box under SWIZZLE_128B is
  // padded to 128-byte lines by the hardware, probe/tma_small_probe.cu)
  if (shape.tma_load && shape.in_mode == kRowMode && (lg - rho[0]) < 4) {
    info->error = "TMA load needs >= 16 contiguous rows per K line"; return false;
  }
  if (shape.tma_load && shape.in_mode == kColMode && (lg - rho[0]) > 8) {
    info->error = "column-mode TMA load: more than 256 rows per K line"; return false;
  }
  // column mode: tiles of 16 columns (full 32-byte sectors, SWIZZLE_32B atoms) when the unit has >= 16 columns
  plan->tma_load = shape.tma_load ? (shape.in_mode == kColMode ? ((ups >= 6 && !shape.no_col64 && !cl) ? 5u
                                                                  : (ups == 5 && !shape.no_col64 && !cl) ? 6u : ups >= 4 ? 4u : 2u)
                                                               : ((lg - rho[0]) < 6 ? 3u : 1u)) : 0u;
  if (shape.ring) {
    // parts = the top two row bits of the stage-1 operand (natural row order: m, then u)
    const uint32_t M = 1u << (lg - rho[0]);
    if (cl || kb || eps != 15 || (plan->tma_load != 1 && plan->tma_load != 2)) {
      info->error = "ring units: 2^15 elements, TMA tiles of 64-row atoms or 8-column tiles, no cluster / Kronecker stage"; return false;
    }
    plan->ring = 1;
    if (plan->tma_load == 2) {
      if (ups != 3 || M < 4) { info->error = "ring units: column tiles need exactly 8 columns"; return false; }
      plan->ring_c2_step = M / 4;
    } else if (ups >= 2) {
      plan->ring_c3_step = (1u << ups) / 4;
    } else if (ups == 0 && M >= 256) {
      plan->ring_c2_step = M / 256;
    } else { info->error = "ring units: 1 or >= 4 transforms per unit"; return false; }
  }
  // the first-half epilogue of stage 2 writes the first half of stage 3's layout, which must lie inside the half of
  // stage 2's layout that its first-half MMAs have consumed: padded plane sizes shrink with the radix, so R_3 >= R_2
  const bool pipe2 = shape.pipe_stage2 && s == 3 && rho[2] >= rho[1] && !cl;
  plan->pipe_stage2 = pipe2 ? 1u : 0u;
  {
    int lo = lg;
    for (int t = 0; t < s; ++t) { lo -= rx[t]; info->lo_bit[t] = lo; }
  }
  if (cl) {
    if (info->lo_bit[0] < 1) { info->error = "cluster unit: stage 1 leaves no index bit to split on"; return false; }
    info->cin = {LBit::R, 0, (uint8_t)(info->lo_bit[0] - 1)};
    info->cout = {LBit::K, 1, (uint8_t)(rho[0] - 1)};
  }
  const LBit cin = info->cin, cout = info->cout;
  uint32_t max_plane = 0;
  for (int t = 0; t < s; ++t) {
    plan->log2_radix[t] = rho[t];
    const uint32_t rows = 1u << (eps - rho[t]);
    if (rows < 128) { info->error = "unit too small for a 128-row tile"; return false; }
    plan->n_tiles[t] = rows / 128;
    plan->chunk_stride[t] = (16u << rho[t]) + ((t == 0 && plan->tma_load == 2) ? 0u : 16u);
    const uint32_t pb = (rows / 8) * plan->chunk_stride[t];
    if (pb > max_plane && !(shape.ring && t == 0)) max_plane = pb;   // ring units: stage 1 reads the landing ring
  }
  {
    uint32_t need = (1u << eps) / 64, c = 32;
    while (c < need) c <<= 1;
    plan->tmem_cols = c;
  }

  auto is_kbit_of_stage = [&](const LBit& b, int t, int* p) {  // is b a bit of the K index of stage t (t = 1..s)?
    if (kb && t == s && b.kind == LBit::U) { *p = rx[s - 1] + b.idx; return true; }
    if (b.kind != LBit::R) return false;
    const int lo = info->lo_bit[t - 1];
    if (b.idx >= lo && b.idx < lo + rx[t - 1]) { *p = b.idx - lo; return true; }
    return false;
  };
  auto row_pos_bytes = [&](int t, int p) { return (1u << (p - 3)) * plan->chunk_stride[t - 1]; };

  // ---------------- load-phase item order (decides the writer-varying bits of stage 1)
  std::vector<LBit>& lb = info->load_bits;
  lb.clear();
  if (shape.in_mode == kRowMode) {
    for (int i = 3; i < lg; ++i) lb.push_back({LBit::R, 0, (uint8_t)i});
    for (int b = 0; b < ups; ++b) lb.push_back({LBit::U, 0, (uint8_t)b});
  } else {
    for (int b = 3; b < ups; ++b) lb.push_back({LBit::U, 0, (uint8_t)b});
    for (int i = 0; i < lg; ++i) lb.push_back({LBit::R, 0, (uint8_t)i});
  }
  if (cl) lb.erase(lb.begin() + find_bit(lb, cin));   // the CTA's rank
  plan->load_item_bits = static_cast<uint32_t>(lb.size());
  std::vector<LBit> writer_varying(lb.begin(), lb.begin() + 3);

  // ---------------- row orders of the s operand layouts
  // in-chunk bits (row positions 0..2) are forced by whoever writes the layout with 16-byte
  // stores; positions 3..5 are chosen so that the writer's 8 quarter-warp lanes hit 8
  // different 16-byte bank groups.
  for (int t = 1; t <= s; ++t) {
    std::vector<LBit>& rb = info->row_bits[t - 1];
    rb.clear();
    if (t == 1) {
      for (int i = 0; i < 3; ++i)
        rb.push_back(shape.in_mode == kRowMode ? LBit{LBit::R, 0, (uint8_t)i} : LBit{LBit::U, 0, (uint8_t)i});
      if (plan->tma_load >= 4) rb.push_back({LBit::U, 0, 3});   // 16 columns = one atom of 16 rows
      if (plan->tma_load >= 5) rb.push_back({LBit::U, 0, 4});   // 32 columns = one atom of 32 rows
      if (plan->tma_load == 5) rb.push_back({LBit::U, 0, 5});   // 64 columns = one atom of 64 rows
    } else {
      for (int i = 0; i < 3; ++i) rb.push_back({LBit::K, (uint8_t)(t - 1), (uint8_t)i});
    }
    std::vector<LBit> all;
    for (int i = 0; i < info->lo_bit[t - 1]; ++i) all.push_back({LBit::R, 0, (uint8_t)i});
    for (int tt = 1; tt < t; ++tt)
      for (int i = 0; i < rho[tt - 1]; ++i) all.push_back({LBit::K, (uint8_t)tt, (uint8_t)i});
    if (cl) {   // stage 1 works on one half of m_1, the later stages on one half of k_1: that bit is the CTA's rank
      const int drop = find_bit(all, t == 1 ? cin : cout);
      if (drop < 0) { info->error = "cluster bit missing"; return false; }
      all.erase(all.begin() + drop);
    }
    if (!(kb && t == s))
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
      if (t == 1 && shape.tma_load) break;   // TMA writes the layout: natural row order (m, then u)
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
    if ((int)rb.size() != eps - rho[t - 1]) { info->error = "row bit count mismatch"; return false; }
    if (pipe2 && t >= 2) {
      const LBit hbit = {LBit::K, 1, (uint8_t)(rho[0] - 1)};
      const int pos = find_bit(rb, hbit);
      if (pos < 6) { info->error = "pipeline bit is not free"; return false; }
      rb.erase(rb.begin() + pos);
      rb.push_back(hbit);
    }
    writer_varying.assign(rb.begin(), rb.begin() + 3);  // the epilogue of this stage writes the next layout
  }

  // ---------------- load maps
  {
    const std::vector<LBit>& rb = info->row_bits[0];
    for (size_t i = 0; i < lb.size(); ++i) {
      int p;
      if (is_kbit_of_stage(lb[i], 1, &p)) plan->load_sofs[i] = k_bit_bytes(p);
      else {
        const int pos = find_bit(rb, lb[i]);
        plan->load_sofs[i] = pos >= 3 ? row_pos_bytes(1, pos) : 0u;
      }
      plan->load_uval[i] = lb[i].kind == LBit::U ? (1u << lb[i].idx) : 0u;
    }
  }

  // ---------------- staging layout (output of the last epilogue)
  // chunk = 8 consecutive k_s[0..2]; chunk bits = row bits of stage s plus k_s[3..rho_s).
  std::vector<LBit> obits;  // logical bits of the output index o, LSB first
  for (int t = 1; t <= s; ++t)
    for (int i = 0; i < rx[t - 1]; ++i) obits.push_back({LBit::K, (uint8_t)t, (uint8_t)i});
  std::vector<LBit> addr_bits;  // output address bits, fastest first
  if (kb) {   // output row = the high bits of the Kronecker stage's output digit
    addr_bits = obits;
    for (int b = 0; b < kb; ++b) addr_bits.push_back({LBit::K, (uint8_t)s, (uint8_t)(rx[s - 1] + b)});
  } else if (shape.out_mode == kRowMode) {
    addr_bits = obits;
    for (int b = 0; b < ups; ++b) addr_bits.push_back({LBit::U, 0, (uint8_t)b});
  } else {
    for (int b = 0; b < ups; ++b) addr_bits.push_back({LBit::U, 0, (uint8_t)b});
    for (auto& b : obits) addr_bits.push_back(b);
  }
  for (int i = 0; i < 3; ++i) info->store_x[i] = addr_bits[i];
  bool dense_staging = false;
  std::vector<LBit> chunk_logical = info->row_bits[s - 1];
  for (int i = 3; i < rho[s - 1]; ++i) chunk_logical.push_back({LBit::K, (uint8_t)s, (uint8_t)i});
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
      if (cl && b == cout) continue;
      g.push_back(b);
    }
    bool used[3] = {false, false, false};
    std::vector<LBit> gnew;
    for (auto& b : g) {
      int p = find_bit(sc, b);
      if (p >= 0) used[p] = true; else gnew.push_back(b);
    }
    // ring units: when every lane-varying bit of the store phase already is an in-run bit of the staging (column-mode
    // output of 8-column units: G = the three lowest output-index bits = the in-chunk row bits of the last stage) the
    // 16-byte padding per 128-byte run is not needed, and the planes must stay at 64 KiB + operand padding to leave
    // room for the landing ring
    dense_staging = shape.ring && gnew.empty();
    // positions 3,4,5 have residues 1,2,4 (chunk address = d + (d >> 3)); place new G bits on free residues
    LBit mid[3] = {}; bool mid_set[3] = {false, false, false};
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
    const uint32_t sb = (nchunks + (dense_staging ? 0u : (nchunks >> 3))) * 16;
    if (sb > max_plane) max_plane = sb;
  }
  plan->plane_bytes = (max_plane + 127u) & ~127u;
  auto staging_contrib = [&](const LBit& b) -> uint32_t {
    int pd = find_bit(info->stage_chunk_bits, b);
    if (pd < 0) return 0;
    uint32_t c = (1u << pd) + ((pd >= 3 && !dense_staging) ? (1u << (pd - 3)) : 0u);
    return c * 16;
  };

  // ---------------- epilogue maps
  for (int t = 1; t <= s; ++t) {
    UnitPlan::Epi& e = plan->epi[t - 1];
    const std::vector<LBit>& rb = info->row_bits[t - 1];
    for (size_t i = 0; i < rb.size(); ++i) {
      const LBit& b = rb[i];
      if (t < s) {
        int p;
        if (is_kbit_of_stage(b, t + 1, &p)) e.dst[i] = k_bit_bytes(p);
        else e.dst[i] = row_pos_bytes(t + 1, find_bit(info->row_bits[t], b));
        e.aux[i] = b.kind == LBit::R ? (1u << b.idx) : 0u;   // m_t = the remaining low digits
      } else {
        e.dst[i] = staging_contrib(b);
        int po = find_bit(obits, b);
        e.aux[i] = po >= 0 ? (1u << po) : 0u;
        e.col[i] = b.kind == LBit::U ? (1u << b.idx) : 0u;
      }
    }
    for (int i = 3; i < rho[t - 1]; ++i) {
      const LBit kb = {LBit::K, (uint8_t)t, (uint8_t)i};
      if (cl && kb == cout) { e.dst_k[i - 3] = 0; continue; }   // selects the destination CTA, not an address
      e.dst_k[i - 3] = t < s ? row_pos_bytes(t + 1, find_bit(info->row_bits[t], kb)) : staging_contrib(kb);
    }
    if (cl && t == 1) {
      int p;
      plan->cl_in_dst = is_kbit_of_stage(cin, 2, &p) ? k_bit_bytes(p) : row_pos_bytes(2, find_bit(info->row_bits[1], cin));
      plan->cl_in_aux = 1u << cin.idx;
    }
    if (cl && t == s) plan->cl_out_aux = 1u << find_bit(obits, cout);
    if (t < s) {
      e.tw_mode = 1;
      // N_t = 2^(lo_bit[t-1] + rho_t); unit angle 2*pi/L: x = m_t * (L / N_t) * k
      e.tw_shift = lg - (info->lo_bit[t - 1] + rx[t - 1]);
      const int tile_bits = static_cast<int>(rb.size()) - 7;
      if (tile_bits > 5) { info->error = "too many tiles per stage"; return false; }
      const double two_pi = 6.283185307179586476925286766559;
      const int64_t L = int64_t(1) << lg;
      for (int tile = 0; tile < (1 << tile_bits); ++tile) {
        int64_t x = 0;
        for (int j = 0; j < tile_bits; ++j)
          if ((tile >> j) & 1) x += e.aux[7 + j];
        x = (x << e.tw_shift) % L;
        auto unit = [&](int64_t ph, float* out) {   // exact on the axes
          ph %= L;
          double c = std::cos(-two_pi * static_cast<double>(ph) / static_cast<double>(L));
          double sn = std::sin(-two_pi * static_cast<double>(ph) / static_cast<double>(L));
          if (ph == 0) { c = 1; sn = 0; }
          else if (4 * ph == L) { c = 0; sn = -1; }
          else if (2 * ph == L) { c = -1; sn = 0; }
          else if (4 * ph == 3 * L) { c = 0; sn = 1; }
          out[0] = static_cast<float>(c); out[1] = static_cast<float>(sn);
        };
        unit(x, e.tile_tw[tile]);
        unit(16 * x, e.tile_tw16[tile]);
      }
    } else {
      e.tw_mode = 0;
      e.tw_kw = 1u << (lg - rx[s - 1]);  // weight of k_s in o
    }
  }

  if (cl) {
    const uint32_t M = 1u << info->lo_bit[0];   // rows per K line of stage 1
    plan->cl_load_c2 = shape.in_mode == kRowMode ? M / 128 : M / 2;
    if (shape.tma_load && shape.in_mode == kRowMode && M < 128) { info->error = "cluster row tiles need M >= 128"; return false; }
  }

  // ---------------- store maps
  {
    std::vector<LBit>& sb = info->store_bits;
    sb.clear();
    for (size_t i = 3; i < addr_bits.size(); ++i) {
      const LBit& b = addr_bits[i];
      if (b.kind == LBit::K && b.stage == s && b.idx < 3) continue;
      if (cl && b == cout) continue;
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
  uint32_t n_units = 0;
  uint32_t col_base_stride = 0;
  uint32_t col_div = 1;
  bool col_from_u = true;    // tw_mode 2 column index includes the unit-local transform index u
  uint32_t n_transforms = 0;
  uint32_t pass1_log2n = 0;  // != 0: multiply outputs by exp(-2*pi*i*o*(col_base+u)/2^pass1_log2n)
  uint32_t tma_batch_step = 0;
  uint32_t b3_units = 0;       // != 0 (power of two): units per outer batch element, which lie in/out_b3_stride apart
  int64_t in_b3_stride = 0, out_b3_stride = 0;
  uint32_t tma_b3_step = 0;
  bool tma_row5 = false;
  int64_t out_hi_stride = 0;   // tiled row-mode output (2-D row pass writing the column units' operand order):
  int out_hi_from = 0;         //   output-index bit i >= out_hi_from has stride out_hi_stride << (i - out_hi_from)
  uint32_t kron_log2n = 0;   // Kronecker units, != 0: multiply output row k_y by exp(-2*pi*i*k_y*col_base/2^kron_log2n)
};

inline void fill_strides(const UnitStrides& st, const PlanBuildInfo& info, UnitPlan* plan) {
  const int s = plan->stages;
  auto in_contrib = [&](const LBit& b) -> uint32_t {
    if (b.kind == LBit::U) return static_cast<uint32_t>((plan->in_mode == kRowMode ? st.in_tstride : 1) << b.idx);
    if (b.kind == LBit::R) return static_cast<uint32_t>((plan->in_mode == kRowMode ? 1 : st.in_nstride) << b.idx);
    return 0;
  };
  for (size_t i = 0; i < info.load_bits.size(); ++i) plan->load_gofs[i] = in_contrib(info.load_bits[i]);
  auto o_weight = [&](const LBit& b) -> int64_t {
    int w = 0;
    for (int t = 1; t < b.stage; ++t) w += info.rx[t - 1];
    return int64_t(1) << (w + b.idx);
  };
  auto out_contrib = [&](const LBit& b) -> uint32_t {
    if (info.kron_bits && b.kind == LBit::K && b.stage == s && b.idx >= info.rx[s - 1])   // output row bit
      return static_cast<uint32_t>(st.out_tstride << (b.idx - info.rx[s - 1]));
    if (b.kind == LBit::U) return static_cast<uint32_t>((plan->out_mode == kRowMode ? st.out_tstride : 1) << b.idx);
    const int64_t w = o_weight(b);
    // row-mode output into a tiled layout: output-index bits >= out_hi_from advance by out_hi_stride per unit step
    if (plan->out_mode == kRowMode && st.out_hi_stride && w >= (int64_t(1) << st.out_hi_from))
      return static_cast<uint32_t>((w >> st.out_hi_from) * st.out_hi_stride);
    return static_cast<uint32_t>(w * (plan->out_mode == kRowMode ? 1 : st.out_nstride));
  };
  for (size_t i = 0; i < info.store_bits.size(); ++i) plan->store_gofs[i] = out_contrib(info.store_bits[i]);
  for (int i = 0; i < 3; ++i) plan->store_cg[i] = out_contrib({LBit::K, (uint8_t)s, (uint8_t)i});
  if (info.cluster) {
    plan->cl_load_gofs = in_contrib(info.cin);
    plan->cl_out_gofs = out_contrib(info.cout);
  }
  plan->in_batch_stride = st.in_batch_stride; plan->in_unit_stride = st.in_unit_stride;
  plan->out_batch_stride = st.out_batch_stride; plan->out_unit_stride = st.out_unit_stride;
  plan->units_per_batch = st.units_per_batch;
  plan->upb_shift = 31;
  for (uint32_t i = 0; i < 31; ++i)
    if (st.units_per_batch == (1u << i)) plan->upb_shift = i;
  plan->col_shift = 0;
  for (uint32_t i = 0; i < 31; ++i)
    if (st.col_div == (1u << i)) plan->col_shift = i;
  plan->tma_batch_step = st.tma_batch_step;
  plan->b3_shift = 31;
  for (uint32_t i = 0; i < 31; ++i)
    if (st.b3_units == (1u << i)) plan->b3_shift = i;
  plan->in_b3_stride = st.in_b3_stride; plan->out_b3_stride = st.out_b3_stride;
  plan->tma_b3_step = st.tma_b3_step;
  plan->tma_row5 = st.tma_row5 ? 1u : 0u;
  plan->n_units = st.n_units;
  plan->col_base_stride = st.col_base_stride;
  plan->col_div = st.col_div ? st.col_div : 1;
  plan->n_transforms = st.n_transforms;
  if (st.pass1_log2n) {
    plan->epi[s - 1].tw_mode = 2;
    plan->epi[s - 1].tw_log2n = st.pass1_log2n;
  }
  if (info.kron_bits && st.kron_log2n) {
    plan->epi[s - 1].tw_mode = 3;
    plan->epi[s - 1].tw_log2n = st.kron_log2n;
    plan->epi[s - 1].tw_shift = static_cast<uint32_t>(info.rx[s - 1]);
  }
  if (!st.col_from_u)
    for (int i = 0; i < kMaxRowBits; ++i) plan->epi[s - 1].col[i] = 0;
}

}  // namespace tfft
