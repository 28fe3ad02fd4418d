// K1 / K2(dX): fused LoRA projection on the 5th-gen tensor cores (sm_100a), bf16 in, f32 accumulate.
//
//     Ts = scaling * (X  A^T)                 [M,R]    rank-R down projection, accumulator in TMEM
//     Y  = X W^T + bias + Ts B^T              [M,N]    frozen base GEMM + rank-R up projection
//
// One kernel serves the forward (X, W, lora_A, lora_B -> Y, Ts; modules/lora.py:12-14 + loralib
// Linear.forward) and the input-gradient half of the backward (dY, W^T, B^T, A^T -> dX, G), because
// dX = dY (W^T)^T + (s dY (B^T)^T) (A^T)^T has exactly the same shape of computation.
//
// Structure (one persistent CTA per SM, 14 warps):
//   warp 0        TMA producer: X / W / lora-down k-blocks into a kStages-deep 128B-swizzled smem ring,
//                 plus the [BN,R] lora-up tile once per output tile
//   warp 1        tcgen05.mma issuer (one elected lane).  Main accumulator [128,BN] f32, double buffered
//                 in TMEM; rank accumulator [128,R], double buffered.  The "tail" of a tile -- Ts B^T and the
//                 bias, both as UMMAs -- is issued a few k-blocks INTO the next tile, so the tensor pipe
//                 never waits for the rank-R round trip through the side warps.
//   warps 2..5    side warps: rank accumulator TMEM -> scale -> bf16 -> shared-memory A operand (+ t_save),
//                 and the per-tile bias operand (bias as hi+lo bf16 pair against a constant ones operand:
//                 the bias add costs one K=16 UMMA instead of per-element epilogue work)
//   warps 6..13   epilogue: TMEM -> registers -> bf16 -> 64B-swizzled staging -> TMA store (two warps per
//                 TMEM lane quarter, alternating 32-column chunks)
// Work items are (m-tile, n-group); the n-tiles of a group reuse the rank-R intermediate, so the
// down projection is computed once per group, not once per output tile.
#include "sdt_common.cuh"
#include "sm100_ptx.cuh"
#include "lora_gemm.cuh"

#include <mutex>
#include <unordered_map>

namespace sdt {

using namespace ptx;

// debug timeline of CTA 0: slot <- clock64 (only when a trace buffer was installed with sdt_debug_set(10, ptr))
#define SDT_TRACE(slot)                                                              \
  do {                                                                               \
    if (p.trace != nullptr && blockIdx.x == 0 && (slot) < 128) p.trace[(slot)] = clock64(); \
  } while (0)

uint64_t debug_get(int key);

template <int BN_, int R_>
struct LoraGemmCfg {
  static constexpr int BM = 128, BN = BN_, BK = 64, R = R_;
  static constexpr int X_BYTES = BM * BK * 2;              // 16 KiB
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int LA_BYTES = R * BK * 2;              // lora-down k-block [R,64]
  static constexpr int STAGE_BYTES = X_BYTES + W_BYTES + LA_BYTES;
  static constexpr int WS_STAGE_BYTES = X_BYTES + LA_BYTES;   // weight-stationary mode: the ring only carries X (+ lora-down)
  static constexpr int kMaxStages = 8;                        // mbarrier slots (weight-stationary mode can use more stages)
  static constexpr int LB_BYTES = ((BN * R * 2 + 1023) / 1024) * 1024;   // lora-up tile [BN,R]
  static constexpr int KEXT = R + 16;                      // rank-R intermediate + the "ones" k-step (bias)
  static constexpr int T_SBO = (KEXT / 8) * 128;           // bytes between 8-row groups of the A operand
  static constexpr int T_BYTES = (BM / 8) * T_SBO;
  static constexpr int BIAS_BYTES = ((BN * 32 + 1023) / 1024) * 1024;    // [BN,16] bf16, un-swizzled cores
  static constexpr int BAR_BYTES = 256;
  // 8 epilogue warps x ONE [32 rows x 128 B] transpose buffer (the CTA-pair kernel, which runs every large launch, stages a
  // warp's whole share of the tile and releases the accumulator before storing; here the ring needs the shared memory)
  static constexpr int STG_BLOCKS = 1;
  static constexpr int STG_BYTES = 8 * STG_BLOCKS * 4096;
  static constexpr int FIXED_BYTES = 1024 /*align slack*/ + LB_BYTES + STG_BYTES + T_BYTES + BIAS_BYTES + BAR_BYTES;
  static constexpr int RING_BYTES = ((232448 - FIXED_BYTES) / 1024) * 1024;   // everything else feeds the TMA ring
  static constexpr int kStages = RING_BYTES / STAGE_BYTES > 6 ? 6 : RING_BYTES / STAGE_BYTES;
  static constexpr int SMEM_BYTES = FIXED_BYTES + RING_BYTES;
  static constexpr int TMEM_COLS = 512;
  // TMEM: two accumulator buffers of BN + R columns: [main (BN) | rank (R)].  On the first tile of an item ONE UMMA of
  // N = BN + R computes both (the lora-down k-block sits right behind the W k-block in the stage, so [W ; A] is one B
  // operand): the X tile is fetched from shared memory once instead of twice -- the tensor core is bound by its operand
  // fetch (~64 B/clk), and a separate N = R UMMA costs almost half a main one for 1/10 of the FLOPs.
  static constexpr int ACC1_COL = BN + R, T_COL = BN;
  static_assert(2 * BN + 2 * R <= 512, "TMEM budget");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
  static_assert(R == 0 || R == 16 || R == 32 || R == 64, "rank must be padded to 16/32/64");
  static_assert(SMEM_BYTES <= 232448 && kStages >= 3, "shared memory budget");
};

constexpr int kGemmThreads = 14 * 32;

struct LoraGemmParams {
  float scaling;
  int M, N, K;
  int n_probs;            // problems of identical shape in this launch (<= G); items enumerate (m-tile, problem, n-group)
  int has_bias;           // every problem has a bias (all or none)
  int n_tiles, n_groups, group_size, n_items;
  int main;               // 0: only the rank-R projection is computed (t_out), no base GEMM
  int f16;                // operands / outputs are IEEE fp16 instead of bf16
  int ws;                 // weight-stationary: every CTA keeps the W k-blocks of ITS n-tile resident (K <= 320), see launch
  int ws_stages;
  long long* trace;       // debug: clock64 stamps of CTA 0 (null in production)
};

// item -> (problem, first row, n-group).  Consecutive items walk the problems and n-groups of ONE row tile, so CTAs that run
// side by side read the same rows of a shared X from L2.
struct ItemCoord { int prob, m0, g; };
template <int G>
__device__ __forceinline__ ItemCoord decode_item(int item, const LoraGemmParams& p, int BM) {
  ItemCoord c;
  if (G == 1) {
    c.prob = 0;
    c.m0 = (item / p.n_groups) * BM;
    c.g = item % p.n_groups;
  } else {
    const int per_m = p.n_probs * p.n_groups;
    const int mt = item / per_m, rem = item - mt * per_m;
    c.prob = rem / p.n_groups;
    c.m0 = mt * BM;
    c.g = rem - c.prob * p.n_groups;
  }
  return c;
}

template <int BN, int R, int G>
__global__ void __launch_bounds__(kGemmThreads, 1)
lora_gemm_kernel(const __grid_constant__ GemmGroup<G> gm, const LoraGemmParams p) {
  using C = LoraGemmCfg<BN, R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lb_smem = smem + C::RING_BYTES;
  uint8_t* stg_smem = lb_smem + C::LB_BYTES;
  uint8_t* t_smem = stg_smem + C::STG_BYTES;
  uint8_t* bias_smem = t_smem + C::T_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_smem + C::BIAS_BYTES);
  uint64_t* full = bars;                       // [kMaxStages]
  uint64_t* empty = bars + C::kMaxStages;      // [kMaxStages]
  uint64_t* acc_full = bars + 2 * C::kMaxStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* t_full = acc_empty + 2;            // rank accumulator complete (MMA -> side warps)
  uint64_t* t_ready = t_full + 1;              // A operand (+ bias operand) of the tail is in smem (side warps -> MMA)
  uint64_t* lb_full = t_ready + 1;             // lora-up tile landed (TMA -> MMA)
  uint64_t* lb_empty = lb_full + 1;            // tail MMAs of a tile completed (MMA -> producer, side warps)
  uint64_t* w_full = lb_empty + 1;             // weight-stationary: resident W k-blocks landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (p.K + C::BK - 1) / C::BK;
  const bool has_main = p.main != 0;
  const bool has_bias = has_main && p.has_bias != 0;
  const bool has_tail = has_main && (R > 0 || has_bias);   // UMMAs issued after the K loop of a tile
  // weight-stationary layout: [nk resident W k-blocks | ring of (X, lora-down) stages]; otherwise ring of (X, W, lora-down)
  const bool ws = p.ws != 0;
  const int n_stages = ws ? p.ws_stages : C::kStages;
  const int stage_bytes = ws ? C::WS_STAGE_BYTES : C::STAGE_BYTES;
  uint8_t* ring = ws ? smem + nk * C::W_BYTES : smem;

  if (threadIdx.x == 0) SDT_TRACE(0);
  if (warp == 0 && lane == 0) {
    for (int q = 0; q < (G == 1 ? 1 : p.n_probs); ++q) {
      prefetch_tmap(&gm.x[q]);
      if (has_main) prefetch_tmap(&gm.w[q]);
      if (R > 0) { prefetch_tmap(&gm.la[q]); if (has_main) prefetch_tmap(&gm.lb[q]); }
    }
    for (int s = 0; s < C::kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    mbar_init(t_full, 1);
    mbar_init(t_ready, 4);
    mbar_init(lb_full, 1);
    mbar_init(lb_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  if (warp >= 2 && warp < 6) {
    // constant part of the tail's A operand: k-step R/16 is [1, 1, 0, ..., 0] per row (pairs with bias hi/lo)
    const int row = (warp - 2) * 32 + lane;
    uint8_t* trow = t_smem + (row >> 3) * C::T_SBO + (row & 7) * 16;
    *reinterpret_cast<uint4*>(trow + (R / 8) * 128) = make_uint4(p.f16 ? 0x3C003C00u : 0x3F803F80u, 0u, 0u, 0u);   // 1.0, 1.0
    *reinterpret_cast<uint4*>(trow + (R / 8 + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
    // zero the second K chunk of the bias operand once (the first chunk is rewritten per tile)
    for (int n = row; n < BN; n += 128)
      *reinterpret_cast<uint4*>(bias_smem + (n >> 3) * 256 + 128 + (n & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  // The producer only needs the mbarriers it initialised itself: it arrives without waiting, so the first TMA loads
  // are in flight while the other warps finish the TMEM allocation and the constant operands.
  if (warp == 0) asm volatile("bar.arrive 1, %0;" ::"n"(kGemmThreads) : "memory");
  else           asm volatile("bar.sync 1, %0;" ::"n"(kGemmThreads) : "memory");
  tc_fence_after();
  const uint32_t tmem_base = warp == 0 ? 0u : *tmem_slot;
  if (threadIdx.x == 32) SDT_TRACE(1);
  // PDL: the prologue above is independent of the predecessor's results; global memory is touched only from here on
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;       // k-block counter across the whole CTA lifetime
      uint32_t tile_ctr = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord ic = decode_item<G>(item, p, C::BM);
        const int m0 = ic.m0, g = ic.g;
        const CUtensorMap* tm_x = &gm.x[ic.prob];
        const CUtensorMap* tm_w = &gm.w[ic.prob];
        const CUtensorMap* tm_la = &gm.la[ic.prob];
        const CUtensorMap* tm_lb = &gm.lb[ic.prob];
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, p.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          const int n0 = nt * C::BN;
          if (ws && tile_ctr == 0) {
            // this CTA's n-tile never changes (grid is a multiple of n_tiles): W and the lora-up tile are loaded once
            mbar_arrive_expect_tx(w_full, (uint32_t)nk * C::W_BYTES);
            for (int kb = 0; kb < nk; ++kb) tma_load_2d(smem + kb * C::W_BYTES, tm_w, kb * C::BK, n0, w_full);
            if (R > 0) {
              mbar_arrive_expect_tx(lb_full, BN * R * 2);
              tma_load_2d(lb_smem, tm_lb, 0, n0, lb_full);
            }
          }
          const uint32_t tx = C::X_BYTES + (has_main && !ws ? C::W_BYTES : 0) + (first ? C::LA_BYTES : 0);
          for (int kb = 0; kb < nk; ++kb, ++it) {
            const int s = it % n_stages;
            mbar_wait(&empty[s], ((it / n_stages) & 1) ^ 1);
            uint8_t* st = ring + s * stage_bytes;
            mbar_arrive_expect_tx(&full[s], tx);
            if (it == 0) SDT_TRACE(2);
            tma_load_2d(st, tm_x, kb * C::BK, m0, &full[s]);
            if (has_main && !ws) tma_load_2d(st + C::X_BYTES, tm_w, kb * C::BK, n0, &full[s]);
            if (first) tma_load_2d(st + C::X_BYTES + (ws ? 0 : C::W_BYTES), tm_la, kb * C::BK, 0, &full[s]);
          }
          if (R > 0 && has_main && !ws) {
            mbar_wait(lb_empty, (tile_ctr & 1) ^ 1);
            mbar_arrive_expect_tx(lb_full, BN * R * 2);
            tma_load_2d(lb_smem, tm_lb, 0, n0, lb_full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr int RR = R > 0 ? R : 16;
    const bool f16 = p.f16 != 0;
    const uint32_t idesc_main = idesc_operand_format(make_idesc_bf16(128, BN, 0, 0), f16);
    const uint32_t idesc_both = idesc_operand_format(make_idesc_bf16(128, BN + R, 0, 0), f16);      // [W ; lora-down] as one B operand
    const uint32_t idesc_t = idesc_operand_format(make_idesc_bf16(128, RR, 0, 0), f16);
    constexpr uint64_t d_sw128 = make_smem_desc_base(16, 1024, kLayoutSW128);
    // lora-up tile [BN,R], K-major, rows of R*2 bytes written by TMA with the matching swizzle
    constexpr uint32_t lb_layout = R == 64 ? kLayoutSW128 : (R == 32 ? kLayoutSW64 : kLayoutSW32);
    constexpr uint64_t d_lb = make_smem_desc_base(16, 8 * RR * 2, lb_layout);
    // tail A operand [128,R+16] and bias operand [BN,16]: K-major, un-swizzled core matrices (8 rows x 16 B = 128 B):
    // K-adjacent cores 128 B apart (LBO), 8-row groups SBO apart
    constexpr uint64_t d_t = make_smem_desc_base(128, C::T_SBO, kLayoutNone);
    constexpr uint64_t d_bias = make_smem_desc_base(128, 256, kLayoutNone);
    uint32_t it = 0, tile_ctr = 0, first_ctr = 0, ready_ctr = 0;
    // a finished K loop whose tail UMMAs (Ts B^T, bias) have not been issued yet
    bool pending = false, pend_needs_ready = false;
    uint32_t pend_tile = 0, pend_ready = 0;

    auto tail_ready = [&]() -> bool {
      // all 32 lanes probe the same barriers, so the result is warp-uniform
      if (R > 0 && !mbar_test(lb_full, ws ? 0u : (pend_tile & 1))) return false;   // ws: loaded once -> phase 0 stays complete
      if (pend_needs_ready && !mbar_test(t_ready, pend_ready & 1)) return false;
      return true;
    };
    auto issue_tail = [&]() {
      // tail of tile pend_tile: acc += Ts B^T (R/16 k-steps) + ones x bias (1 k-step); then hand the accumulator over
      if (R > 0) mbar_wait(lb_full, ws ? 0u : (pend_tile & 1));
      if (pend_needs_ready) mbar_wait(t_ready, pend_ready & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (pend_tile & 1) * C::ACC1_COL;
        const uint32_t ta = smem_u32(t_smem), ba = smem_u32(lb_smem);
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          umma_f16_ss(d, smem_desc(d_t, ta + k * 256), smem_desc(d_lb, ba + k * 32), idesc_main, 1u);
        if (has_bias) umma_f16_ss(d, smem_desc(d_t, ta + (R / 16) * 256), smem_desc(d_bias, smem_u32(bias_smem)), idesc_main, 1u);
        umma_commit(lb_empty);
        umma_commit(&acc_full[pend_tile & 1]);
      }
      __syncwarp();
      if (lane == 0 && pend_tile < 6) SDT_TRACE(11 + 4 * pend_tile);
      pending = false;
    };

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int g = decode_item<G>(item, p, C::BM).g;
      const int nt0 = g * p.group_size;
      const int nt1 = min(nt0 + p.group_size, p.n_tiles);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const bool first = (nt == nt0) && R > 0;
        const uint32_t buf = tile_ctr & 1;
        const uint32_t d_main = tmem_base + buf * C::ACC1_COL;
        const uint32_t d_tacc = d_main + C::T_COL;           // rank columns of the same buffer
        if (has_main) {
          // The accumulator buffer may still be with the epilogue.  A pending tail must not wait for that as well: its
          // epilogue would start only after this one has finished and the two could never overlap (timeline in profiles/:
          // store-bound shapes ran at epilogue + tail latency per tile instead of the epilogue alone).
          while (pending && !mbar_test(&acc_empty[buf], ((tile_ctr >> 1) & 1) ^ 1))
            if (tail_ready()) issue_tail();
          mbar_wait(&acc_empty[buf], ((tile_ctr >> 1) & 1) ^ 1);
          tc_fence_after();
        } else if (first_ctr >= 1) {
          // rank-only mode: the side warps must have drained the previous item's rank accumulator
          mbar_wait(t_ready, (first_ctr - 1) & 1);
          tc_fence_after();
        }
        if (lane == 0 && tile_ctr < 6) SDT_TRACE(8 + 4 * tile_ctr);
        if (ws && tile_ctr == 0) mbar_wait(w_full, 0);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % n_stages;
          mbar_wait(&full[s], (it / n_stages) & 1);
          if (lane == 0 && tile_ctr < 6 && kb == 0) SDT_TRACE(9 + 4 * tile_ctr);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t xa = smem_u32(ring + s * stage_bytes);
            const uint32_t wa = ws ? smem_u32(smem + kb * C::W_BYTES) : xa + C::X_BYTES;
            const uint32_t la = ws ? xa + C::X_BYTES : wa + C::W_BYTES;
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) {
              const uint64_t a_desc = smem_desc(d_sw128, xa + k * 32);
              if (has_main && first && !ws)
                umma_f16_ss(d_main, a_desc, smem_desc(d_sw128, wa + k * 32), idesc_both, (kb | k) != 0);
              else {
                if (has_main) umma_f16_ss(d_main, a_desc, smem_desc(d_sw128, wa + k * 32), idesc_main, (kb | k) != 0);
                if (first) umma_f16_ss(d_tacc, a_desc, smem_desc(d_sw128, la + k * 32), idesc_t, (kb | k) != 0);
              }
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          // the previous tile's tail goes out as soon as its operands are in place -- at the latest before this
          // tile's K loop ends (its accumulator hand-over must not wait for a whole tile)
          if (pending && (kb == nk - 1 || tail_ready())) issue_tail();
        }
        if (lane == 0 && tile_ctr < 6) SDT_TRACE(10 + 4 * tile_ctr);
        if (first) {
          if (elect_one()) umma_commit(t_full);
          __syncwarp();
          ++first_ctr;
        }
        if (has_main) {
          if (has_tail) {
            pending = true;
            pend_tile = tile_ctr;
            // the bias operand is rewritten per tile, except in weight-stationary mode where it is written once
            pend_needs_ready = first || (has_bias && !(ws && tile_ctr > 0));
            pend_ready = ready_ctr;
            if (pend_needs_ready) ++ready_ctr;
            if (tail_ready()) issue_tail();      // non-first tiles: operands are usually already there
          } else {
            if (elect_one()) umma_commit(&acc_full[buf]);
            __syncwarp();
          }
        }
      }
    }
    if (pending) issue_tail();
  } else if (warp < 6) {
    // ===================================== side warps ========================================
    // per tile: (a) bias operand of the tail, (b) on the first tile of an item: rank-R intermediate
    constexpr int RR = R > 0 ? R : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tid = (warp - 2) * 32 + lane;       // 0..127
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t tile_ctr = 0, first_ctr = 0;
    if (has_tail || (!has_main && R > 0)) {
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord ic = decode_item<G>(item, p, C::BM);
        const int m0 = ic.m0, g = ic.g;
        const float* bias = gm.bias[ic.prob];
        __nv_bfloat16* t_out = gm.t_out[ic.prob];
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, p.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          if (has_bias && !(ws && tile_ctr > 0)) {       // ws: the n-tile (hence the bias operand) never changes
            // the previous tile's tail must have finished reading the bias operand
            if (tile_ctr > 0) mbar_wait(lb_empty, (tile_ctr - 1) & 1);
            const int n0 = nt * C::BN;
            for (int n = tid; n < BN; n += 128) {
              const float b = (n0 + n < p.N) ? __ldg(bias + n0 + n) : 0.f;
              const float hi = round_act(b, p.f16 != 0);
              *reinterpret_cast<uint4*>(bias_smem + (n >> 3) * 256 + (n & 7) * 16) =
                  make_uint4(pack_act2(hi, b - hi, p.f16 != 0), 0u, 0u, 0u);
            }
          }
          if (first) {
            mbar_wait(t_full, first_ctr & 1);
            tc_fence_after();
            uint32_t packed[RR / 2];
#pragma unroll
            for (int c = 0; c < RR / 16; ++c) {
              uint32_t v[16];
              tmem_ld_x16(lane_addr + (tile_ctr & 1) * C::ACC1_COL + C::T_COL + c * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                packed[c * 8 + j] = pack_act2(__uint_as_float(v[2 * j]) * p.scaling, __uint_as_float(v[2 * j + 1]) * p.scaling, p.f16 != 0);
            }
            if (has_main) {
              uint8_t* trow = t_smem + (row >> 3) * C::T_SBO + (row & 7) * 16;
#pragma unroll
              for (int kc = 0; kc < RR / 8; ++kc)
                *reinterpret_cast<uint4*>(trow + kc * 128) =
                    make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
            }
            if (t_out != nullptr && g == 0 && m0 + row < p.M) {
              uint4* dst = reinterpret_cast<uint4*>(t_out + (size_t)(m0 + row) * RR);
#pragma unroll
              for (int kc = 0; kc < RR / 8; ++kc)
                dst[kc] = make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
            }
            ++first_ctr;
          }
          if (first || (has_bias && !(ws && tile_ctr > 0))) {
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_ready);
            if (warp == 2 && lane == 0 && tile_ctr < 6) SDT_TRACE(40 + tile_ctr);
          }
        }
      }
    }
  } else {
    // ===================================== epilogue warps ====================================
    // TMEM -> registers -> bf16 -> per-warp transpose buffer -> global, in blocks of [32 rows x 64 columns] (see
    // write_staged_block in sm100_ptx.cuh): full 128-byte lines per row and request.
    const int e = warp - 6;                       // 0..7
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = e >> 2;                      // which of the two warps of this quarter
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = smem_u32(stg_smem + e * C::STG_BLOCKS * 4096);
    uint32_t tile_ctr = 0;
    if (has_main) {
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord ic = decode_item<G>(item, p, C::BM);
        const int m0 = ic.m0, g = ic.g;
        uint8_t* yp = gm.y[ic.prob];
        const uint8_t* rp = gm.res[ic.prob];
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, p.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const uint32_t buf = tile_ctr & 1;
          const int n0 = nt * C::BN;
          mbar_wait(&acc_full[buf], (tile_ctr >> 1) & 1);
          if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE(48 + 2 * tile_ctr);
          tc_fence_after();
          // 32-column sub-chunks of this tile that hold live columns; 64-column blocks alternate between the two warps of a
          // quarter, and the warp that starts alternates with the tile (BN = 160 is 2.5 blocks)
          const int n_sub = min(C::BN / 32, (p.N - n0 + 31) / 32);
          uint32_t v[32];
          for (int cb = (tile_ctr + half) & 1; 2 * cb < n_sub; cb += 2) {
            const int subs = min(2, n_sub - 2 * cb);
            for (int h = 0; h < subs; ++h) {
              tmem_ld_x32(lane_addr + buf * C::ACC1_COL + (2 * cb + h) * 32, v);
              tmem_ld_wait();
              uint32_t pk[16];
              pack_acc32(v, pk, p.f16 != 0);
              stage_row_chunk(stg, lane, h, pk);
            }
            __syncwarp();
            write_staged_block(stg, lane, yp, m0 + q * 32, p.M, n0 + cb * 64, p.N, 4 * subs, rp, p.f16 != 0);
            __syncwarp();
          }
          // every tcgen05.ld of this buffer has completed (wait::ld above): release it to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE(49 + 2 * tile_ctr);
        }
      }
      if (warp == 6 && lane == 0) SDT_TRACE(62);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SDT_TRACE(63);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
    if (lane == 0) SDT_TRACE(64);
  }
}

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// Tensor maps are pure functions of (base, shape, pitch, box, swizzle); weights, packed operands and -- thanks to the
// caching allocator -- most activations keep their addresses from step to step, so the encode (a driver call) is cached.
struct TmapKey {
  uint64_t base, rows, cols, pitch;
  uint32_t box_rows, box_cols, swz;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols && swz == o.swz;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (((uint64_t)k.box_rows << 40) ^ ((uint64_t)k.box_cols << 8) ^ k.swz) + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmaps_mu;

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, TmapSwizzle swz) {
  const TmapKey key{reinterpret_cast<uint64_t>(base), rows, cols, pitch_bytes, box_rows, box_cols, (uint32_t)swz};
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return SDT_OK; }
  }
  PFN_encodeTiled enc = get_encode();
  SDT_REQUIRE(enc != nullptr, SDT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  SDT_REQUIRE(aligned16(base) && pitch_bytes % 16 == 0, SDT_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swz == TMAP_SW_128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swz == TMAP_SW_64  ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swz == TMAP_SW_32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                   : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDT_REQUIRE(r == CUDA_SUCCESS, SDT_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] pitch %llu box [%u x %u]", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes, box_rows, box_cols);
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    if (g_tmaps.size() > 16384) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return SDT_OK;
}

// A [rows, cols] matrix seen as its two row halves, [2][rows/2][cols] (128-byte swizzle): a box of box_rows >= rows/2 rows that
// starts at a NEGATIVE row coordinate brings one half surrounded by zero rows (out-of-bounds elements are zero-filled and still
// counted by the barrier).  The summed-source kernel uses it to place a source's lora-down rows inside a block of zeros.
int make_tmap_halves_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                          uint32_t box_rows, uint32_t box_cols) {
  const TmapKey key{reinterpret_cast<uint64_t>(base), rows, cols, pitch_bytes, box_rows, box_cols, 0x100u | (uint32_t)TMAP_SW_128};
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return SDT_OK; }
  }
  PFN_encodeTiled enc = get_encode();
  SDT_REQUIRE(enc != nullptr, SDT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  SDT_REQUIRE(aligned16(base) && pitch_bytes % 16 == 0 && rows % 2 == 0, SDT_ERR_ARG,
              "TMA operand must be 16-byte aligned with a 16-byte row pitch and an even number of rows");
  cuuint64_t dims[3] = {cols, rows / 2, 2};
  cuuint64_t strides[2] = {pitch_bytes, (rows / 2) * pitch_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDT_REQUIRE(r == CUDA_SUCCESS, SDT_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) for 2 x [%llu x %llu] pitch %llu box [%u x %u]", (int)r,
              (unsigned long long)rows / 2, (unsigned long long)cols, (unsigned long long)pitch_bytes, box_rows, box_cols);
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    if (g_tmaps.size() > 16384) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return SDT_OK;
}

// Pick the number of n-groups per m-tile: fewer groups = less recomputation of the rank-R projection, more groups =
// more work items to balance over the SMs.  Cost of a schedule ~ rounds * (columns per item + per-item overhead).
static void choose_groups(int m_tiles, int n_tiles, int BN, int R, int sms, int* group_size, int* n_groups) {
  if (R == 0) { *group_size = 1; *n_groups = n_tiles; return; }
  double best = 1e30;
  int best_gs = 1;
  for (int gs = 1; gs <= n_tiles; ++gs) {
    const int groups = (n_tiles + gs - 1) / gs;
    const long items = (long)m_tiles * groups;
    const long rounds = (items + sms - 1) / sms;
    const double cost = (double)rounds * (gs * (double)BN + R + 48.0);
    if (cost < best - 1e-9) { best = cost; best_gs = gs; }
  }
  *group_size = best_gs;
  *n_groups = (n_tiles + best_gs - 1) / best_gs;
}

template <int BN, int R, int G>
static int launch_lora_gemm(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, bool main,
                            bool f16, cudaStream_t st) {
  using C = LoraGemmCfg<BN, R>;
  static bool attr_set = false;
  if (!attr_set) {
    SDT_CUDA_OK(cudaFuncSetAttribute(lora_gemm_kernel<BN, R, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  GemmGroup<G> gm;
  for (int q = 0; q < G; ++q) {
    const LoraProblem& pr = probs[q < n_probs ? q : 0];     // unused slots repeat problem 0 (never dereferenced)
    int rc = make_tmap_2d_bf16(&gm.x[q], pr.x, M, K, K * 2, C::BM, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    if (main) {
      rc = make_tmap_2d_bf16(&gm.w[q], pr.w, N, K, K * 2, BN, C::BK, TMAP_SW_128);
      if (rc != SDT_OK) return rc;
    } else {
      gm.w[q] = gm.x[q];
    }
    gm.y[q] = reinterpret_cast<uint8_t*>(pr.y);
    gm.ym[q] = gm.x[q];                                     // this kernel stores from registers
    gm.ym32[q] = gm.x[q];
    if (R > 0) {
      rc = make_tmap_2d_bf16(&gm.la[q], pr.la, R, K, K * 2, R, C::BK, TMAP_SW_128);
      if (rc != SDT_OK) return rc;
      if (main) {
        rc = make_tmap_2d_bf16(&gm.lb[q], pr.lb, N, R, (uint64_t)R * 2, BN, R,
                               R == 64 ? TMAP_SW_128 : (R == 32 ? TMAP_SW_64 : TMAP_SW_32));
        if (rc != SDT_OK) return rc;
      } else {
        gm.lb[q] = gm.la[q];
      }
    } else {
      gm.la[q] = gm.x[q];
      gm.lb[q] = gm.x[q];
    }
    gm.bias[q] = pr.bias;
    gm.t_out[q] = reinterpret_cast<__nv_bfloat16*>(pr.t_out);
    gm.res[q] = reinterpret_cast<const uint8_t*>(pr.res);
  }
  LoraGemmParams p;
  p.scaling = scaling;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.n_probs = n_probs;
  p.has_bias = probs[0].bias != nullptr ? 1 : 0;
  p.main = main ? 1 : 0;
  p.f16 = f16 ? 1 : 0;
  p.trace = reinterpret_cast<long long*>(debug_get(10));
  const int m_tiles = (int)((M + C::BM - 1) / C::BM);
  p.n_tiles = main ? (int)((N + BN - 1) / BN) : 1;
  const int sms = num_sms();
  // the problems of a launch multiply the number of row tiles a scheduling round can draw from
  choose_groups(m_tiles * n_probs, p.n_tiles, BN, R, sms, &p.group_size, &p.n_groups);
  p.n_items = m_tiles * n_probs * p.n_groups;
  int grid = p.n_items < sms ? p.n_items : sms;
  // (Experimental, opt-in with sdt_debug_set(13, 2): measured no gain.  tools/umma_bench.cu shows why: with A read from
  // shared memory a UMMA costs N/2 + 38 cycles, so a schedule that pays a separate N = R rank UMMA on every tile is bound by
  // the tensor core's operand fetch even when the ring only carries X.)
  // Weight-stationary schedule for short K loops (K <= 320: the whole [BN, K] slab of W fits beside the ring): CTA c keeps
  // the W k-blocks, lora-up tile and bias of n-tile (c mod n_tiles) resident and streams only X.  Needs every CTA to see >= 2
  // row tiles.
  const int nk = (int)((K + C::BK - 1) / C::BK);
  const int ws_stages_fit = (C::RING_BYTES - nk * C::W_BYTES) / C::WS_STAGE_BYTES;
  p.ws = 0;
  p.ws_stages = 0;
  if (G == 1 && main && nk * C::W_BYTES < C::RING_BYTES && ws_stages_fit >= 3 && p.n_tiles <= sms && debug_get(13) == 2) {
    const int g = (sms / p.n_tiles) * p.n_tiles;
    if ((long)m_tiles * p.n_tiles >= 2L * g) {
      p.ws = 1;
      p.ws_stages = ws_stages_fit > C::kMaxStages ? C::kMaxStages : ws_stages_fit;
      p.group_size = 1;
      p.n_groups = p.n_tiles;
      p.n_items = m_tiles * p.n_tiles;
      grid = g;
    }
  }
  SDT_CUDA_OK(launch_kernel(lora_gemm_kernel<BN, R, G>, dim3(grid), dim3(kGemmThreads), C::SMEM_BYTES, st, true, gm, p));
  SDT_LAUNCH_OK("lora_gemm");
  return SDT_OK;
}

static int check_group(const LoraProblem* probs, int n_probs, int r, bool main) {
  SDT_REQUIRE(probs != nullptr && n_probs >= 1 && n_probs <= kMaxGroup, SDT_ERR_ARG,
              "lora_gemm: a launch takes 1..%d problems (got %d)", kMaxGroup, n_probs);
  for (int q = 0; q < n_probs; ++q) {
    const LoraProblem& pr = probs[q];
    SDT_REQUIRE(pr.x != nullptr && (!main || (pr.w != nullptr && pr.y != nullptr)), SDT_ERR_ARG, "lora_gemm: null operand in problem %d", q);
    SDT_REQUIRE((r == 0) == (pr.la == nullptr) && (r == 0 || !main || pr.lb != nullptr), SDT_ERR_ARG,
                "lora_gemm: lora operands must be given exactly when r > 0 (problem %d)", q);
    SDT_REQUIRE((pr.bias != nullptr) == (probs[0].bias != nullptr), SDT_ERR_ARG,
                "lora_gemm: the problems of one launch must all have a bias or all have none");
    SDT_REQUIRE(aligned16(pr.x) && aligned16(pr.w) && aligned16(pr.la) && aligned16(pr.lb) && aligned16(pr.y) && aligned16(pr.t_out) &&
                    aligned16(pr.res), SDT_ERR_ARG, "lora_gemm: pointers must be 16-byte aligned (problem %d)", q);
    SDT_REQUIRE(pr.res == nullptr || main, SDT_ERR_ARG, "lora_gemm: a residual needs the base projection (problem %d)", q);
  }
  return SDT_OK;
}

// bf16 entry used by sdt_lora_linear_fwd(_group) / sdt_lora_linear_bwd (lora_api.cu): n_probs problems of one shape
int lora_gemm_group_bf16(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, int r, bool main,
                         bool f16, cudaStream_t st) {
  SDT_REQUIRE(M > 0 && K > 0 && N > 0, SDT_ERR_ARG, "lora_gemm: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
  SDT_REQUIRE(M < (1ll << 31) && K < (1ll << 31) && N < (1ll << 31), SDT_ERR_UNSUPPORTED, "lora_gemm: dimension exceeds int32");
  SDT_REQUIRE(K % 8 == 0 && N % 8 == 0, SDT_ERR_UNSUPPORTED, "lora_gemm: K and N must be multiples of 8 (K=%lld N=%lld)", (long long)K, (long long)N);
  SDT_REQUIRE(r == 0 || r == 16 || r == 32 || r == 64, SDT_ERR_UNSUPPORTED,
              "lora_gemm: padded rank must be 0, 16, 32 or 64 (got %d)", r);
  SDT_REQUIRE(main || r > 0, SDT_ERR_ARG, "lora_gemm: nothing to compute");
  int rc = check_group(probs, n_probs, r, main);
  if (rc != SDT_OK) return rc;
  SDT_REQUIRE(n_probs == 1 || r > 0, SDT_ERR_UNSUPPORTED, "lora_gemm: grouped launches are built for r > 0 only");
  // CTA-pair (cta_group::2) kernel when there is a base GEMM and at least one full pair of row tiles: each SM loads half of
  // every B-type operand, and since the accumulator hand-over no longer carries a GPU-scope fence it wins at every K of the
  // step (A/B in profiles/r01_gemm_ab_*.txt: K = 320 was the last hold-out).  Very short K loops stay on the single-CTA kernel.
  // sdt_debug_set(11, 1) forces the single-CTA kernel (A/B measurements)
  const int64_t pair_min_k = debug_get(14) ? (int64_t)debug_get(14) : 256;
  if (main && M >= 256 && K >= pair_min_k && debug_get(11) == 0)
    return lora_gemm_pair_group_bf16(probs, n_probs, scaling, M, K, N, r, f16, st);
  const bool bn160 = !main || (N % 160 == 0) || (N % 128 != 0 && N > 128);
#define SDT_GEMM(BN, R, G) return launch_lora_gemm<BN, R, G>(probs, n_probs, scaling, M, K, N, main, f16, st)
#define SDT_GEMM_R(BN, G)                                                                     \
  switch (r) { case 16: SDT_GEMM(BN, 16, G); case 32: SDT_GEMM(BN, 32, G); default: SDT_GEMM(BN, 64, G); }
  if (n_probs == 1) {
    if (r == 0) { if (bn160) SDT_GEMM(160, 0, 1); else SDT_GEMM(128, 0, 1); }
    if (bn160) { SDT_GEMM_R(160, 1) } else { SDT_GEMM_R(128, 1) }
  }
  if (bn160) { SDT_GEMM_R(160, kMaxGroup) } else { SDT_GEMM_R(128, kMaxGroup) }
#undef SDT_GEMM_R
#undef SDT_GEMM
}

int lora_gemm_bf16(const void* x, const void* w, const float* bias, const void* la, const void* lb, float scaling, void* y,
                   void* t_out, int64_t M, int64_t K, int64_t N, int r, bool main, bool f16, cudaStream_t st, const void* res) {
  const LoraProblem pr{x, w, bias, la, lb, y, t_out, res};
  return lora_gemm_group_bf16(&pr, 1, scaling, M, K, N, r, main, f16, st);
}

}  // namespace sdt
