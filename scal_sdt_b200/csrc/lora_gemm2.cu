// K1 / K2(dX), CTA-pair variant: the same fused LoRA projection as lora_gemm.cu, but two SMs of one TPC work on
// one 256 x BN tile with tcgen05 `cta_group::2`.
//
// Why: the single-CTA kernel is bound by shared-memory FILL bandwidth (L2 -> SM), not by the tensor pipe: a 128 x 160
// tile ingests (128 + 160) x 64 x 2 B per k-block for 2.6 MFLOP, ~71 FLOP/B, and an SM ingests ~50 B/clk (timeline in
// profiles/).  With cta_group::2 each CTA still owns 128 rows of X but only HALF of the W tile (the tensor core reads
// the other half from the peer's shared memory), so the same math needs (128 + 80) x 64 x 2 B per SM: ~100 FLOP/B.
//
// Roles per CTA are those of lora_gemm.cu (producer / MMA issuer / 4 side warps / 8 epilogue warps) with these changes:
//   * both CTAs' TMA loads complete on the LEADER's (cluster rank 0) full barrier; only the leader issues UMMAs
//     (M = 256) and multicasts its tcgen05.commit arrivals to both CTAs;
//   * side and epilogue warps of both CTAs arrive remotely on the leader's t_ready / acc_empty barriers;
//   * every B-type operand (W, lora-down, lora-up, bias) is split by rows between the two CTAs.
#include "sdt_common.cuh"
#include "sm100_ptx.cuh"
#include "lora_gemm.cuh"

// the double-tile instantiations leave their work-item loop early (`continue` in front of the per-tile loop)
#pragma nv_diag_suppress 128

namespace sdt {

using namespace ptx;

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, TmapSwizzle swz);
int make_tmap_halves_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                          uint32_t box_rows, uint32_t box_cols);
uint64_t debug_get(int key);

// debug timeline of pair 0's leader CTA (same slots as lora_gemm.cu; enabled with sdt_debug_set(10, ptr))
#define SDT_TRACE2(slot)                                                                      \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0 && (slot) < 128) p.trace[(slot)] = clock64();  \
  } while (0)

namespace pair {

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared-memory object) inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
  return out;
}
// Arrive on a barrier of the peer CTA after writing operands into OWN shared memory.  Release at CTA scope (what CUTLASS'
// ClusterBarrier::arrive does): the writes were already pushed to the async proxy by fence.proxy.async, and the tensor core
// that reads them is this CTA's own.  `.release.cluster` would add a MEMBAR.ALL.GPU, which queues behind the epilogue's store
// traffic (~1000+ cycles per item, timeline in profiles/).
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Arrive WITHOUT release semantics.  `.release.cluster` compiles to MEMBAR.ALL.GPU + arrive: after the epilogue's burst of
// global stores that barrier waited ~1700 cycles per tile for the stores to be acknowledged by L2 (timeline in profiles/),
// although the hand-over only protects TMEM, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already order.
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tile load into OWN shared memory; the bytes complete on the barrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, int c_inner, int c_row, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_row), "r"(bar_cluster_addr)
      : "memory");
}
// the same for a box of a 3-D tensor map (make_tmap_halves_bf16: inner element, row inside the half -- may be negative --, half)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, int c_inner, int c_row, int c_half, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_row), "r"(c_half), "r"(bar_cluster_addr)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TS mode: the A operand (this CTA's 128 rows x 16 k) comes from tensor memory (lane = row, 32-bit column c = elements 2c, 2c+1)
__device__ __forceinline__ void umma2_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128 rows x 32 bytes (one K = 16 slice of a K-major tile, any swizzle the descriptor names) from shared memory into 8 TMEM columns,
// in both CTAs of the pair; ordered with the UMMAs of the issuing thread (tcgen05.cp and tcgen05.mma execute in issue order)
__device__ __forceinline__ void cp2_128x256b(uint32_t taddr, uint64_t s_desc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}
// arrive on the barrier at the same offset in BOTH CTAs when all prior UMMAs of this thread have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// the same, on the leader's barrier only
__device__ __forceinline__ void umma2_commit_leader(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)1)
      : "memory");
}

}  // namespace pair

using namespace pair;

// S = number of SOURCES whose products are summed into one output (S = 1: a plain projection; S = 3: the input gradient of
// q / k / v, dX = sum_s dY_s W_s + (s dY_s B_s) A_s, as ONE K loop over the concatenated contraction).  Every source has its own
// rank-R accumulator, lora-up tile and slice of the tail's A operand.
// DT ("double tile"): the two column tiles of a work item run ONE joint K loop -- every X k-block is brought into the SM once and
// feeds both accumulator buffers (stage = X | W of tile A | lora-down | W of tile B).  The K loop is bound by the bytes an SM can
// land from L2 (~55 B/clk, profiles/r02_tma_fill_rate.txt): 2 x 26.6 KB per k-block pair become 36.9 KB.  The price: both TMEM
// buffers belong to the item, so its epilogue no longer overlaps the next item's K loop -- for long K loops only.
template <int BN_, int R_, int S_ = 1, bool DT_ = false>
struct PairCfg {
  static constexpr int BM = 128, BN = BN_, BK = 64, R = R_, HN = BN_ / 2, HR = R_ / 2, S = S_;
  static constexpr int X_BYTES = BM * BK * 2;                       // own 128 rows of X
  static constexpr int W_BYTES = HN * BK * 2;                       // own half of the W tile
  // own half of the lora-down k-block; with S sources a block of S HR rows in which only the current source's rows are non-zero
  static constexpr int LA_ROWS = S * HR;
  static constexpr int LA_BYTES = ((LA_ROWS * BK * 2 + 1023) / 1024) * 1024;
  static constexpr bool DT = DT_;
  static constexpr int STAGE_BYTES = X_BYTES + W_BYTES + LA_BYTES + (DT ? W_BYTES : 0);
  static constexpr int LB_TILE = (((HN + LA_ROWS) * R * 2 + 1023) / 1024) * 1024;    // own half of one lora-up tile [BN/2, R] + S HR zero rows
  static constexpr int LB_BYTES = S * LB_TILE;
  static constexpr int KEXT = S * R + 16;
  static constexpr int T_SBO = (KEXT / 8) * 128;
  static constexpr int T_BYTES = (BM / 8) * T_SBO;
  static constexpr int BIAS_BYTES = (((HN + HR) * 32 + 255) / 256) * 256;      // own half of the bias operand [BN/2, 16] + HR zero rows (un-swizzled: no 1 KiB alignment; the 256 bytes saved are the fourth stage of the rank-16 double-tile kernel)
  static constexpr int BAR_BYTES = 256;
  // 64-column blocks of staging per epilogue warp.  2 (a warp's whole share of a <= 160-wide tile: the accumulator is handed
  // back before the stores, which pays for K = 320) or 1 (224-wide tiles: a pipeline stage is worth more there -- measured)
  static constexpr int STG_BLOCKS = BN <= 160 ? 2 : 1;
  // per epilogue warp: [32 rows x 128 B] transpose buffers.  A 160-wide tile gives a warp one full 64-column block and at most one odd
  // 32-column block, staged as [32 x 64 B]: 4 + 2 KiB instead of 2 x 4 -- the 16 KiB are a pipeline stage (5 -> 6 at rank 16; a
  // ring capped at 4 stages measured 7 % slower than 5, profiles/r02_pipeline_depth_ab.txt)
  static constexpr int STG_WARP = STG_BLOCKS == 2 && BN == 160 ? 6144 : STG_BLOCKS * 4096;
  static constexpr int STG_BYTES = 8 * STG_WARP;
  static constexpr int FIXED_BYTES = 1024 + LB_BYTES + STG_BYTES + T_BYTES + BIAS_BYTES + BAR_BYTES;
  static constexpr int kStagesMax = (232448 - FIXED_BYTES) / STAGE_BYTES;
#ifndef SDT_STAGE_CAP
#define SDT_STAGE_CAP 8              // experiments: -DSDT_STAGE_CAP=n builds the kernels with a shallower ring
#endif
  static constexpr int kStages = kStagesMax > SDT_STAGE_CAP ? SDT_STAGE_CAP : kStagesMax;
  static constexpr int SMEM_BYTES = FIXED_BYTES + kStages * STAGE_BYTES;
  static constexpr int TMEM_COLS = 512;
  // TMEM: two accumulator buffers of BN + R columns.  On the first tile of an item ONE UMMA of N = BN + R computes the base
  // GEMM and the rank projection together ([W half ; lora-down half] is one B operand per CTA), so X is fetched once.  With
  // cta_group::2 the N index runs over CTA 0's rows, then CTA 1's, so the columns of such a tile are
  //   [Y 0..HN) | T 0..HR) | Y HN..BN) | T HR..R)]      (acc_col(y) = y < HN ? y : y + HR ; acc_col(t) = t < HR ? HN + t : BN + t)
  // and the tail UMMAs use the same N with zero rows appended to the lora-up / bias operands.  Other tiles are plain [Y].
  // With S > 1 sources the rank block of a CTA has S HR rows, source s owning rows [s HR, (s+1) HR) and the TMA zero-filling the
  // others (a source adds 0 to the other sources' columns), so the first tile is
  //   [Y 0..HN) | T_0 0..HR) .. T_{S-1} 0..HR) | Y HN..BN) | T_0 HR..R) .. T_{S-1} HR..R)]
  // A separate N = R UMMA for the rank projection costs 39 cycles next to the N/2 = 80 of the BN-wide one: in a CTA pair a UMMA takes
  // max(39, N/2) cycles (profiles/r02_umma_pair_ss_vs_ts.txt).
  static constexpr int ACC1_COL = BN + S * R;
  static_assert(2 * ACC1_COL <= 512, "TMEM budget");
  // TS mode: k-blocks of X ([128 rows x 64 k] = 32 columns) staged in the TMEM columns behind the two accumulators
  static constexpr int A_TM_COL = 2 * ACC1_COL;
  static constexpr int A_SLOTS = (512 - A_TM_COL) / 32;
  static_assert(BN % 32 == 0 && HN % 8 == 0 && BN <= 256, "BN");
  static_assert(R == 0 || R == 16 || R == 32 || R == 64, "rank must be padded to 16/32/64");
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(W_BYTES % 1024 == 0, "operand tiles must keep 1024-byte alignment");
  static_assert(!DT || S == 1, "double tiles: one source");
};

constexpr int kPairThreads = 15 * 32;      // producer, UMMA issuer, 4 side warps, 8 epilogue warps, tail issuer

struct PairParams {
  float scaling;
  int M, N, K;
  int n_probs;            // problems of identical shape in this launch (<= G)
  int n_src;              // sources summed into each output (<= S; their operands are entries 0..n_src-1 of the group)
  int has_bias;
  int f16;                // operands / outputs are IEEE fp16 instead of bf16
  int n_tiles, n_groups, group_size, n_items;   // items are (256-row tile, problem, n-group)
  // GEGLU epilogue (ff.net.0.proj): N = 2 I, tile nt holds h columns [nt HN, (nt+1) HN) and the matching gate columns
  // I + [nt HN, (nt+1) HN); besides proj (y, for the backward) the epilogue writes act = h * gelu(gate) [M, I]
  int geglu_I;
  uint8_t* act_out;
  int mixed;              // problems have their own output width (GemmGroup::n / tile_begin); one column tile per work item
  int dt_items;           // double-tile kernel: every work item is exactly two column tiles (checked by the launcher)
  int ts;                 // A operand through tensor memory (tcgen05.cp of each k-block, then TS-mode UMMAs); needs C::A_SLOTS >= 1
  long long* trace;       // debug: clock64 stamps of CTA 0 (null in production)
};

// item -> (problem, 256-row tile index, n-group); consecutive items share the row tile (see lora_gemm.cu).  N / n_tiles are the
// problem's own in a mixed-width launch, the launch's otherwise.
struct PairItem { int prob, mt, g, N, n_tiles; };
template <int G>
__device__ __forceinline__ PairItem decode_pair_item(int item, const PairParams& p, const GemmGroup<G>& gm) {
  PairItem c;
  c.N = p.N;
  c.n_tiles = p.n_tiles;
  if (G == 1) {
    c.prob = 0;
    c.mt = item / p.n_groups;
    c.g = item % p.n_groups;
  } else if (p.mixed == 0) {
    const int per_m = p.n_probs * p.n_groups;
    c.mt = item / per_m;
    const int rem = item - c.mt * per_m;
    c.prob = rem / p.n_groups;
    c.g = rem - c.prob * p.n_groups;
  } else {
    const int per_m = gm.tile_begin[p.n_probs];
    c.mt = item / per_m;
    const int rem = item - c.mt * per_m;
    int q = 0;
    while (q + 1 < p.n_probs && rem >= gm.tile_begin[q + 1]) ++q;
    c.prob = q;
    c.g = rem - gm.tile_begin[q];
    c.N = gm.n[q];
    c.n_tiles = gm.tile_begin[q + 1] - gm.tile_begin[q];
  }
  return c;
}

template <int BN, int R, int G, int S, bool GEGLU = false, bool DT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
lora_gemm_pair_kernel(const __grid_constant__ GemmGroup<G> gm, const PairParams p) {
  using C = PairCfg<BN, R, S, DT>;
  static_assert(!DT || (G == 1 && S == 1 && !GEGLU), "double tiles: one plain problem");
  static_assert(S == 1 || (G >= S && R > 0), "summed sources live in the entries of the group");
  static_assert(!GEGLU || (G == 1 && S == 1 && R > 0 && C::STG_BLOCKS >= 2 && C::HN % 8 == 0), "GEGLU epilogue: one merged problem, <= 160-wide tiles");
  const int n_src = S == 1 ? 1 : p.n_src;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lb_smem = smem + C::kStages * C::STAGE_BYTES;
  uint8_t* stg_smem = lb_smem + C::LB_BYTES;
  uint8_t* t_smem = stg_smem + C::STG_BYTES;
  uint8_t* bias_smem = t_smem + C::T_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_smem + C::BIAS_BYTES);
  uint64_t* full = bars;                       // [kStages]  leader's is the live one
  uint64_t* empty = bars + C::kStages;         // [kStages]  both (multicast commit)
  uint64_t* acc_full = bars + 2 * C::kStages;  // [2]        both
  uint64_t* acc_empty = acc_full + 2;          // [2]        leader's (16 remote/local arrivals)
  uint64_t* t_full = acc_empty + 2;            //            both
  uint64_t* t_ready = t_full + 1;              //            leader's (8 arrivals)
  uint64_t* lb_full = t_ready + 1;             //            leader's
  uint64_t* lb_empty = lb_full + 1;            //            both
  uint64_t* kdone = lb_empty + 1;              // [2]        leader's: every UMMA of the tile's K loop has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kdone + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nk = (p.K + C::BK - 1) / C::BK;
  const bool has_bias = p.has_bias != 0;
  const bool has_tail = R > 0 || has_bias;
  const bool f16 = p.f16 != 0;

  if (threadIdx.x == 0) SDT_TRACE2(0);
  if (warp == 0) {
    if (lane == 0) {
      for (int q = 0; q < (S > 1 ? n_src : (G == 1 ? 1 : p.n_probs)); ++q) {
        prefetch_tmap(&gm.x[q]);
        prefetch_tmap(&gm.w[q]);
        if (R > 0) { prefetch_tmap(&gm.la[q]); prefetch_tmap(&gm.lb[q]); }
      }
      for (int q = 0; q < (S > 1 || G == 1 ? 1 : p.n_probs); ++q) { prefetch_tmap(&gm.ym[q]); prefetch_tmap(&gm.ym32[q]); }
      for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
      mbar_init(t_full, 1);
      mbar_init(t_ready, 8);
      mbar_init(lb_full, 1);
      mbar_init(lb_empty, 1);
      mbar_init(&kdone[0], 1);
      mbar_init(&kdone[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  // Cluster barrier, split: every thread ARRIVES now (warp 0 after the barrier init above, the only thing the peer CTA must
  // see) and waits later.  The producer waits at once and starts its TMA loads while the other warps are still allocating
  // TMEM and writing the constant operands: the first k-block lands ~1200 cycles earlier than with a full cluster sync here.
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (warp == 1) tmem_alloc2(tmem_slot, C::TMEM_COLS);
  if (warp >= 2 && warp < 6) {
    const int row = (warp - 2) * 32 + lane;
    uint8_t* trow = t_smem + (row >> 3) * C::T_SBO + (row & 7) * 16;
    *reinterpret_cast<uint4*>(trow + (S * R / 8) * 128) = make_uint4(f16 ? 0x3C003C00u : 0x3F803F80u, 0u, 0u, 0u);   // 1.0, 1.0
    *reinterpret_cast<uint4*>(trow + (S * R / 8 + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
    for (int n = row; n < C::HN + C::HR; n += 128) {
      *reinterpret_cast<uint4*>(bias_smem + (n >> 3) * 256 + 128 + (n & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
      if (n >= C::HN) *reinterpret_cast<uint4*>(bias_smem + (n >> 3) * 256 + (n & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    // zero rows behind the lora-up tile (rows of R*2 bytes; the TMA box only ever writes the first HN rows)
    for (int i = row; i < C::LA_ROWS * R * 2 / 16; i += 128)
      for (int q = 0; q < S; ++q)
        *reinterpret_cast<uint4*>(lb_smem + q * C::LB_TILE + C::HN * R * 2 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  // both CTAs' barriers are initialised before anything crosses the pair; the producer does not wait for the CTA-local set-up
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("bar.arrive 1, %0;" ::"n"(kPairThreads) : "memory");
  else           asm volatile("bar.sync 1, %0;" ::"n"(kPairThreads) : "memory");
  tc_fence_after();
  const uint32_t tmem_base = warp == 0 ? 0u : *tmem_slot;
  if (threadIdx.x == 32) SDT_TRACE2(1);
  // PDL: everything above touched only shared memory, TMEM and the kernel parameters.  From here on the kernel reads what its
  // predecessor in the stream wrote (and overwrites buffers it may still be reading), so wait for it -- and let our own
  // dependent start its prologue on every SM this grid has already left.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (lane == 0) {
      const uint32_t lb_full_leader = map_to_rank(lb_full, 0);
      uint32_t it = 0, tile_ctr = 0;
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const PairItem ic = decode_pair_item<G>(item, p, gm);
        const int m0 = ic.mt * 2 * C::BM + (int)rank * C::BM;
        const int g = ic.g;
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, ic.n_tiles);
        if (DT) {
          // tiles nt0 (A: carries the rank projection) and nt0 + 1 (B) in one joint K loop: X once per k-block
          const int nA = nt0 * C::BN + (int)rank * C::HN, nB = nA + C::BN;
          const uint32_t tx = 2u * (C::X_BYTES + 2 * C::W_BYTES + (R > 0 ? C::LA_ROWS * C::BK * 2 : 0));
          for (int kb = 0; kb < nk; ++kb, ++it) {
            const int s = it % C::kStages;
            mbar_wait(&empty[s], ((it / C::kStages) & 1) ^ 1);
            uint8_t* st = smem + s * C::STAGE_BYTES;
            const uint32_t full_leader = map_to_rank(&full[s], 0);
            if (leader) mbar_arrive_expect_tx(&full[s], tx);
            if (it == 0) SDT_TRACE2(2);
            tma_load_2d_pair(st, &gm.x[0], kb * C::BK, m0, full_leader);
            tma_load_2d_pair(st + C::X_BYTES, &gm.w[0], kb * C::BK, nA, full_leader);
            if (R > 0) tma_load_2d_pair(st + C::X_BYTES + C::W_BYTES, &gm.la[0], kb * C::BK, (int)rank * C::HR, full_leader);
            tma_load_2d_pair(st + C::X_BYTES + C::W_BYTES + C::LA_BYTES, &gm.w[0], kb * C::BK, nB, full_leader);
          }
          if (R > 0) {
            for (int h = 0; h < 2; ++h, ++tile_ctr) {
              mbar_wait(lb_empty, (tile_ctr & 1) ^ 1);
              if (leader) mbar_arrive_expect_tx(lb_full, 2u * C::HN * R * 2);
              tma_load_2d_pair(lb_smem, &gm.lb[0], 0, h == 0 ? nA : nB, lb_full_leader);
            }
          } else {
            tile_ctr += 2;
          }
          continue;
        }
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          // rows of W / lora-up this CTA contributes to the tile.  cta_group::2 runs the N index over CTA 0's rows, then CTA 1's:
          // for GEGLU CTA 0 brings the h rows and CTA 1 the gate rows of the same output columns -- no permuted weight copy
          const int n0 = GEGLU ? nt * C::HN + (int)rank * p.geglu_I : nt * C::BN + (int)rank * C::HN;
          const uint32_t tx = 2u * (C::X_BYTES + C::W_BYTES + (first ? C::LA_ROWS * C::BK * 2 : 0));     // zero-filled rows count too
          for (int src = 0; src < n_src; ++src) {
            const int q = S > 1 ? src : ic.prob;          // operand set: the source, or the problem of a grouped launch
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int s = it % C::kStages;
              mbar_wait(&empty[s], ((it / C::kStages) & 1) ^ 1);
              uint8_t* st = smem + s * C::STAGE_BYTES;
              const uint32_t full_leader = map_to_rank(&full[s], 0);
              if (leader) mbar_arrive_expect_tx(&full[s], tx);
              if (it == 0) SDT_TRACE2(2);
              tma_load_2d_pair(st, &gm.x[q], kb * C::BK, m0, full_leader);
              tma_load_2d_pair(st + C::X_BYTES, &gm.w[q], kb * C::BK, n0, full_leader);
              if (first) {
                // this CTA's half of the lora-down k-block; with S sources as rows [src HR, (src + 1) HR) of a zero-filled block
                if (S == 1) tma_load_2d_pair(st + C::X_BYTES + C::W_BYTES, &gm.la[q], kb * C::BK, (int)rank * C::HR, full_leader);
                else        tma_load_3d_pair(st + C::X_BYTES + C::W_BYTES, &gm.la[q], kb * C::BK, -src * C::HR, (int)rank, full_leader);
              }
            }
          }
          if (R > 0) {
            mbar_wait(lb_empty, (tile_ctr & 1) ^ 1);
            if (leader) mbar_arrive_expect_tx(lb_full, 2u * n_src * C::HN * R * 2);
            for (int src = 0; src < n_src; ++src)
              tma_load_2d_pair(lb_smem + src * C::LB_TILE, &gm.lb[S > 1 ? src : ic.prob], 0, n0, lb_full_leader);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== UMMA issuer: K loops (leader CTA only) ========================
    // Only the K loops: the tail of a tile (rank-R up-projection and bias, two or three UMMAs behind two or three barriers) is
    // issued by its own warp below.  A lone warp runs ~6 cycles per dependent instruction: with the tails in this loop the issuer
    // spent 1000-1300 cycles between the K loops of consecutive tiles (timeline in profiles/), half of a K = 320 K loop.
    if (leader) {
      const uint32_t idesc_main = idesc_operand_format(make_idesc_bf16(256, BN, 0, 0), f16);
      const uint32_t idesc_both = idesc_operand_format(make_idesc_bf16(256, BN + S * R, 0, 0), f16);    // first tile of an item: [W ; lora-down]
      constexpr uint64_t d_sw128 = make_smem_desc_base(16, 1024, kLayoutSW128);
      uint32_t it = 0, tile_ctr = 0;
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const PairItem ic = decode_pair_item<G>(item, p, gm);
        const int g = ic.g;
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, ic.n_tiles);
        if (DT) {
          // joint K loop of the item's two tiles: buffer 0 <- tile A ([W ; lora-down], N = BN + R), buffer 1 <- tile B; both buffers
          // must have been drained (tile_ctr is even at every item)
          const uint32_t ph = ((tile_ctr >> 1) & 1) ^ 1;
          mbar_wait(&acc_empty[0], ph);
          mbar_wait(&acc_empty[1], ph);
          tc_fence_after();
          if (lane == 0 && tile_ctr < 6) SDT_TRACE2(8 + 4 * tile_ctr);
          const uint32_t dA = tmem_base, dB = tmem_base + C::ACC1_COL;
          for (int kb = 0; kb < nk; ++kb, ++it) {
            const int s = it % C::kStages;
            mbar_wait(&full[s], (it / C::kStages) & 1);
            if (lane == 0 && tile_ctr < 6 && kb == 0) SDT_TRACE2(9 + 4 * tile_ctr);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t xa = smem_u32(smem + s * C::STAGE_BYTES);
              const uint32_t wa = xa + C::X_BYTES, wb = wa + C::W_BYTES + C::LA_BYTES;
#pragma unroll
              for (int k = 0; k < C::BK / 16; ++k) {
                const uint64_t a_desc = smem_desc(d_sw128, xa + k * 32);
                umma2_f16_ss(dA, a_desc, smem_desc(d_sw128, wa + k * 32), R > 0 ? idesc_both : idesc_main, (kb | k) != 0);
                umma2_f16_ss(dB, a_desc, smem_desc(d_sw128, wb + k * 32), idesc_main, (kb | k) != 0);
              }
              umma2_commit_both(&empty[s]);
            }
            __syncwarp();
          }
          if (lane == 0 && tile_ctr < 6) SDT_TRACE2(10 + 4 * tile_ctr);
          if (elect_one()) {
            if (R > 0) umma2_commit_both(t_full);
            if (has_tail) { umma2_commit_leader(&kdone[0]); umma2_commit_leader(&kdone[1]); }
            else          { umma2_commit_both(&acc_full[0]); umma2_commit_both(&acc_full[1]); }
          }
          __syncwarp();
          tile_ctr += 2;
          continue;
        }
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          const uint32_t buf = tile_ctr & 1;
          const uint32_t d_main = tmem_base + buf * C::ACC1_COL;
          mbar_wait(&acc_empty[buf], ((tile_ctr >> 1) & 1) ^ 1);
          tc_fence_after();
          if (lane == 0 && tile_ctr < 6) SDT_TRACE2(8 + 4 * tile_ctr);
          for (int src = 0; src < n_src; ++src) {
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int s = it % C::kStages;
              mbar_wait(&full[s], (it / C::kStages) & 1);
              if (lane == 0 && tile_ctr < 6 && kb == 0 && src == 0) SDT_TRACE2(9 + 4 * tile_ctr);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t xa = smem_u32(smem + s * C::STAGE_BYTES);
                const uint32_t wa = xa + C::X_BYTES;        // [W half ; lora-down block] is one B operand
                if (C::A_SLOTS >= 1 && p.ts) {
                  // TS mode (opt-in, for A/B only): the k-block of X goes to tensor memory first.  A single-CTA SS-mode UMMA pays ~38
                  // cycles for fetching A from shared memory, but a CTA-pair one does not (tools/umma_pair_bench.cu): measured slower
                  constexpr uint32_t kASlots = C::A_SLOTS >= 1 ? C::A_SLOTS : 1;
                  const uint32_t a_tm = tmem_base + C::A_TM_COL + (it % kASlots) * 32;
#pragma unroll
                  for (int k = 0; k < C::BK / 16; ++k) cp2_128x256b(a_tm + k * 8, smem_desc(d_sw128, xa + k * 32));
#pragma unroll
                  for (int k = 0; k < C::BK / 16; ++k)
                    umma2_f16_ts(d_main, a_tm + k * 8, smem_desc(d_sw128, wa + k * 32), first ? idesc_both : idesc_main, (src | kb | k) != 0);
                } else {
                  // every source accumulates into the same Y columns; the zero rows of its lora-down block leave the other
                  // sources' rank columns as they are
#pragma unroll
                  for (int k = 0; k < C::BK / 16; ++k)
                    umma2_f16_ss(d_main, smem_desc(d_sw128, xa + k * 32), smem_desc(d_sw128, wa + k * 32), first ? idesc_both : idesc_main,
                                 (src | kb | k) != 0);
                }
                umma2_commit_both(&empty[s]);
              }
              __syncwarp();
            }
          }
          if (lane == 0 && tile_ctr < 6) SDT_TRACE2(10 + 4 * tile_ctr);
          if (elect_one()) {
            if (first) umma2_commit_both(t_full);                    // side warps: the rank-R intermediate is complete
            if (has_tail) umma2_commit_leader(&kdone[buf]);          // tail issuer: the accumulator is ready for the tail
            else          umma2_commit_both(&acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 14) {
    // ===================================== tail issuer (leader CTA only) ==================================
    // Per tile: Y += T' lb^T (+ bias through the [1, 1, 0...] columns of T') once the K loop has completed (kdone: UMMAs of two
    // threads are ordered only through a commit), the lora-up tile has landed and the side warps have written T' / the bias
    // operand; then the accumulator goes to the epilogue and the lora-up / bias buffers back to their writers.
    if (leader && has_tail) {
      constexpr int RR = R > 0 ? R : 16;
      const uint32_t idesc_main = idesc_operand_format(make_idesc_bf16(256, BN, 0, 0), f16);
      const uint32_t idesc_both = idesc_operand_format(make_idesc_bf16(256, BN + S * R, 0, 0), f16);
      constexpr uint32_t lb_layout = R == 64 ? kLayoutSW128 : (R == 32 ? kLayoutSW64 : kLayoutSW32);
      constexpr uint64_t d_lb = make_smem_desc_base(16, 8 * RR * 2, lb_layout);
      constexpr uint64_t d_t = make_smem_desc_base(128, C::T_SBO, kLayoutNone);
      constexpr uint64_t d_bias = make_smem_desc_base(128, 256, kLayoutNone);
      const uint32_t ta = smem_u32(t_smem), ba = smem_u32(lb_smem);
      uint32_t tile_ctr = 0, ready_ctr = 0;
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const PairItem ic = decode_pair_item<G>(item, p, gm);
        const int nt0 = ic.g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, ic.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          const uint32_t buf = tile_ctr & 1;
          // the last of the three to complete is T' (it needs the K loop, then the side warps): waited for last
          mbar_wait(&kdone[buf], (tile_ctr >> 1) & 1);
          if (R > 0) mbar_wait(lb_full, tile_ctr & 1);
          if (first || has_bias) {
            mbar_wait(t_ready, ready_ctr & 1);
            ++ready_ctr;
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem_base + buf * C::ACC1_COL;
            const uint32_t idesc_tail = first ? idesc_both : idesc_main;   // same column layout as the tile's K loop
            for (int src = 0; src < n_src; ++src) {
#pragma unroll
              for (int k = 0; k < R / 16; ++k)
                umma2_f16_ss(d, smem_desc(d_t, ta + (src * (R / 16) + k) * 256), smem_desc(d_lb, ba + src * C::LB_TILE + k * 32),
                             idesc_tail, 1u);
            }
            if (has_bias) umma2_f16_ss(d, smem_desc(d_t, ta + (S * R / 16) * 256), smem_desc(d_bias, smem_u32(bias_smem)), idesc_tail, 1u);
            umma2_commit_both(lb_empty);
            umma2_commit_both(&acc_full[buf]);
          }
          __syncwarp();
          if (lane == 0 && tile_ctr < 6) SDT_TRACE2(11 + 4 * tile_ctr);
        }
      }
    }
  } else if (warp < 6) {
    // ===================================== side warps (both CTAs) ========================================
    constexpr int RR = R > 0 ? R : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tid = (warp - 2) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_ready_leader = map_to_rank(t_ready, 0);
    uint32_t tile_ctr = 0, first_ctr = 0;
    if (has_tail) {
      for (int item = pair_id; item < p.n_items; item += n_pairs) {
        const PairItem ic = decode_pair_item<G>(item, p, gm);
        const int m0 = ic.mt * 2 * C::BM + (int)rank * C::BM;
        const int g = ic.g;
        const float* bias = gm.bias[ic.prob];
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, ic.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          // the previous tile's tail has read T' and the bias operand (K loops and tails are issued by different threads: nothing
          // else orders the next K loop's completion behind that tail)
          if (tile_ctr > 0 && (first || has_bias)) mbar_wait(lb_empty, (tile_ctr - 1) & 1);
          if (has_bias) {
            const int n0 = GEGLU ? nt * C::HN + (int)rank * p.geglu_I : nt * C::BN + (int)rank * C::HN;
            for (int n = tid; n < C::HN; n += 128) {
              const float b = (n0 + n < ic.N) ? __ldg(bias + n0 + n) : 0.f;
              const float hi = round_act(b, f16);
              *reinterpret_cast<uint4*>(bias_smem + (n >> 3) * 256 + (n & 7) * 16) = make_uint4(pack_act2(hi, b - hi, f16), 0u, 0u, 0u);
            }
          }
          if (first) {
            mbar_wait(t_full, first_ctr & 1);
            tc_fence_after();
            for (int src = 0; src < n_src; ++src) {
              uint32_t packed[RR / 2];
              // two 8-column loads in flight per wait (this round trip is on the critical path of every first tile)
#pragma unroll
              for (int c = 0; c < RR / 8; c += 2) {
                // rank column t of source src lives at HN + src HR + t (t < HR: CTA 0's lora-down rows) or, for CTA 1's rows,
                // behind the Y columns and CTA 0's whole rank block: BN + S HR + src HR + (t - HR)
                uint32_t v[2][8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int t0 = (c + u) * 8;
                  const int col = t0 < C::HR ? C::HN + src * C::HR + t0 : C::BN + (S - 1) * C::HR + src * C::HR + t0;
                  tmem_ld_x8(lane_addr + (tile_ctr & 1) * C::ACC1_COL + col, v[u]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    packed[(c + u) * 4 + j] = pack_act2(__uint_as_float(v[u][2 * j]) * p.scaling, __uint_as_float(v[u][2 * j + 1]) * p.scaling, f16);
              }
              uint8_t* trow = t_smem + (row >> 3) * C::T_SBO + (row & 7) * 16 + src * (RR / 8) * 128;
#pragma unroll
              for (int kc = 0; kc < RR / 8; ++kc)
                *reinterpret_cast<uint4*>(trow + kc * 128) = make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
              __nv_bfloat16* t_out = gm.t_out[S > 1 ? src : ic.prob];
              if (t_out != nullptr && g == 0 && m0 + row < p.M) {
                uint4* dst = reinterpret_cast<uint4*>(t_out + (size_t)(m0 + row) * RR);
#pragma unroll
                for (int kc = 0; kc < RR / 8; ++kc)
                  dst[kc] = make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
              }
            }
            ++first_ctr;
          }
          if (first || has_bias) {
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) remote_arrive(t_ready_leader);
            if (warp == 2 && lane == 0 && tile_ctr < 6) SDT_TRACE2(40 + tile_ctr);
          }
        }
      }
    }
  } else if (warp < 14) {
    // ===================================== epilogue warps (both CTAs) ====================================
    // TMEM -> registers -> bf16 -> per-warp transpose buffer -> global in [32 rows x 64 columns] blocks (see lora_gemm.cu)
    const int e = warp - 6;
    const int q = warp & 3;
    const int half = e >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = smem_u32(stg_smem + e * C::STG_WARP);
    uint32_t acc_empty_leader[2] = {map_to_rank(&acc_empty[0], 0), map_to_rank(&acc_empty[1], 0)};
    uint32_t tile_ctr = 0;
    for (int item = pair_id; item < p.n_items; item += n_pairs) {
      const PairItem ic = decode_pair_item<G>(item, p, gm);
      const int m0 = ic.mt * 2 * C::BM + (int)rank * C::BM;
      const int g = ic.g;
      uint8_t* yp = gm.y[ic.prob];
      const uint8_t* rp = S > 1 ? nullptr : gm.res[ic.prob];
      const int nt0 = g * p.group_size;
      const int nt1 = min(nt0 + p.group_size, ic.n_tiles);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const uint32_t buf = tile_ctr & 1;
        const int n0 = nt * C::BN;
        const int gap = (nt == nt0 && R > 0) ? C::LA_ROWS : 0;      // first tile: Y columns >= HN sit behind CTA 0's rank block
        mbar_wait(&acc_full[buf], (tile_ctr >> 1) & 1);
        if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(48 + 2 * tile_ctr);
        tc_fence_after();
        const int n_sub = GEGLU ? C::BN / 32 : min(C::BN / 32, (ic.N - n0 + 31) / 32);     // GEGLU: I % HN == 0, tiles are full
        uint32_t v[32];
        // 32 output columns of this tile -> registers: one TMEM load, or two 16-column loads where the block straddles the gap of
        // a merged first tile (HN is a multiple of 16, so a half never straddles it)
        auto load_sub = [&](int sub) {
          const int y0 = sub * 32, y1 = y0 + 16;
          if (gap == 0 || y0 >= C::HN || y0 + 32 <= C::HN) {
            tmem_ld_x32(lane_addr + buf * C::ACC1_COL + y0 + (y0 >= C::HN ? gap : 0), v);
          } else {
            uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[0]);
            uint32_t(&hi)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[16]);
            tmem_ld_x16(lane_addr + buf * C::ACC1_COL + y0 + (y0 >= C::HN ? gap : 0), lo);
            tmem_ld_x16(lane_addr + buf * C::ACC1_COL + y1 + (y1 >= C::HN ? gap : 0), hi);
          }
          tmem_ld_wait();
        };
        if constexpr (C::STG_BLOCKS >= 2) {
          // Phase 1: drain this warp's share of the accumulator into its staging buffers and hand the TMEM buffer back at once.
          // Store-bound shapes (K = 320: 8x more bytes out than in per tile) spend ~3000 cycles per tile on the stores; holding
          // the accumulator that long left the MMA warp idle and the store stream with gaps (timeline in profiles/).
          // Plain outputs leave through the TMA: each 64-column block is staged as a [32 x 128 B] box in the 128-byte swizzle (full
          // lines per row: 64-byte rows run the memory system at about half its write bandwidth) and one lane issues its store.  The warp never waits on the TPC's store port (with st.global a warp sat ~2000 cycles per
          // tile in the issue of its stores, and the next accumulator waited for it); rows >= M and columns >= N are clipped by
          // the tensor map.  Residual / GEGLU epilogues keep the [32 x 128 B] staging and the register stores below.
          const bool tma_out = !GEGLU && rp == nullptr;
          if (lane == 0) tma_store_wait_read();      // the previous tile's boxes have left the staging buffers
          __syncwarp();
          int slot = 0;
          for (int cb = (tile_ctr + half) & 1; 2 * cb < n_sub; cb += 2, ++slot) {
            const int subs = min(2, n_sub - 2 * cb);
            for (int h = 0; h < subs; ++h) {
              load_sub(2 * cb + h);
              uint32_t pk[16];
              pack_acc32(v, pk, f16);
              // a full 64-column block is one [32 x 128 B] box in the 128-byte swizzle -- the layout the register path stages in;
              // the odd 32-column block of a tile is a [32 x 64 B] box in the 64-byte swizzle (for the register path too: the
              // second slot of a warp has room for exactly that)
              if (subs == 1) stage_row_sw64(stg + slot * 4096, lane, pk);
              else           stage_row_chunk(stg + slot * 4096, lane, h, pk);
            }
          }
          tc_fence_before();
          if (tma_out) fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) remote_arrive_relaxed(acc_empty_leader[buf]);
          if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(49 + 2 * tile_ctr);
          if (tma_out) {
            if (lane == 0) {
              slot = 0;
              for (int cb = (tile_ctr + half) & 1; 2 * cb < n_sub; cb += 2, ++slot)
                tma_store_2d(n_sub - 2 * cb >= 2 ? &gm.ym[ic.prob] : &gm.ym32[ic.prob], stg + slot * 4096, n0 + cb * 64, m0 + q * 32);
              tma_store_commit();
            }
            __syncwarp();
            if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(70 + tile_ctr);
            continue;
          }
          if constexpr (GEGLU) {
            // Phase 2 (GEGLU): act = h * gelu(gate) from the STAGED (already rounded) halves of the tile -- exactly what the unfused
            // sequence computes from proj in HBM.  h and gate of one output column sit in different 64-column blocks, i.e. in the
            // staging of both epilogue warps of this lane quarter: pair barrier, then each warp takes 16 of the quarter's 32 rows
            // and writes, per row, three full 128-byte lines: h and gate into proj (reference column order [h | gate], kept for the
            // backward) and act.  Lanes walk the 16-byte slots of a row: a store instruction covers four full lines.
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(80 + tile_ctr);
            {
              const uint32_t stg_q = smem_u32(stg_smem) + (uint32_t)(e & 3) * C::STG_WARP;     // half 0's staging of this quarter
              constexpr int kSlots = C::HN / 8;                                      // 16-byte slots per output row (64 columns: 8)
              auto staged = [&](int r, int tile_col) -> uint4 {
                const int cb = tile_col >> 6, sl = (tile_col & 63) >> 3;
                const uint32_t owner = (uint32_t)((cb + tile_ctr) & 1);               // which half staged block cb of this tile
                const uint32_t base = stg_q + owner * (4u * C::STG_WARP) + (uint32_t)(cb >> 1) * 4096u;
                return ld_shared_v4(base + r * 128 + ((sl ^ (r & 7)) << 4));
              };
              for (int task = lane; task < 16 * kSlots; task += 32) {
                const int r = half * 16 + task / kSlots, so = task % kSlots;
                const int grow = m0 + q * 32 + r;
                const uint4 hv = staged(r, 8 * so), gv = staged(r, C::HN + 8 * so);
                const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
                uint32_t ow[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float h0, h1, g0, g1;
                  if (f16) {
                    const float2 hh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
                    const float2 gg = __half22float2(*reinterpret_cast<const __half2*>(&gw[j]));
                    h0 = hh.x; h1 = hh.y; g0 = gg.x; g1 = gg.y;
                  } else {
                    h0 = bf16_bits_to_f32(hw[j] & 0xffffu); h1 = bf16_bits_to_f32(hw[j] >> 16);
                    g0 = bf16_bits_to_f32(gw[j] & 0xffffu); g1 = bf16_bits_to_f32(gw[j] >> 16);
                  }
                  ow[j] = pack_act2(h0 * gelu_erf(g0), h1 * gelu_erf(g1), f16);
                }
                if (grow < p.M) {
                  uint8_t* prow = yp + ((size_t)grow * p.N + nt * C::HN + 8 * so) * 2;
                  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(prow), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
                  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(prow + (size_t)p.geglu_I * 2), "r"(gv.x), "r"(gv.y), "r"(gv.z),
                               "r"(gv.w)
                               : "memory");
                  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p.act_out + ((size_t)grow * p.geglu_I + nt * C::HN + 8 * so) * 2),
                               "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3])
                               : "memory");
                }
              }
            }
            if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(90 + tile_ctr);
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");      // the partner has read my staging: it may be rewritten
          } else {
            // Phase 2: stream the staged blocks out (full 128-byte lines); the next tile's phase 1 follows in program order
            slot = 0;
            for (int cb = (tile_ctr + half) & 1; 2 * cb < n_sub; cb += 2, ++slot)
              write_staged_block(stg + slot * 4096, lane, yp, m0 + q * 32, p.M, n0 + cb * 64, ic.N, 4 * min(2, n_sub - 2 * cb), rp, f16,
                                 /*rows of 64 bytes*/ n_sub - 2 * cb < 2);
          }
          __syncwarp();
          if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(70 + tile_ctr);
        } else {
          // one staging block per warp: stage and store block by block, release the accumulator after the last load
          const bool tma_out = rp == nullptr;
          for (int cb = (tile_ctr + half) & 1; 2 * cb < n_sub; cb += 2) {
            const int subs = min(2, n_sub - 2 * cb);
            if (lane == 0) tma_store_wait_read();    // the previous block's boxes have left the staging buffer
            __syncwarp();
            for (int h = 0; h < subs; ++h) {
              load_sub(2 * cb + h);
              uint32_t pk[16];
              pack_acc32(v, pk, f16);
              if (tma_out && subs == 1) stage_row_sw64(stg, lane, pk);
              else                      stage_row_chunk(stg, lane, h, pk);
            }
            if (tma_out) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(subs == 2 ? &gm.ym[ic.prob] : &gm.ym32[ic.prob], stg, n0 + cb * 64, m0 + q * 32);
                tma_store_commit();
              }
              __syncwarp();
            } else {
              __syncwarp();
              write_staged_block(stg, lane, yp, m0 + q * 32, p.M, n0 + cb * 64, ic.N, 4 * subs, rp, f16);
              __syncwarp();
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) remote_arrive_relaxed(acc_empty_leader[buf]);
          if (warp == 6 && lane == 0 && tile_ctr < 6) SDT_TRACE2(49 + 2 * tile_ctr);
        }
      }
    }
    if (lane == 0) tma_store_wait_read();          // shared memory stays until the last boxes have been read out of it
    __syncwarp();
    if (warp == 6 && lane == 0) SDT_TRACE2(62);
  }

  // neither CTA may leave (or free TMEM) while the other can still touch its shared memory / barriers / TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) SDT_TRACE2(63);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
    if (lane == 0) SDT_TRACE2(64);
  }
}

static void choose_groups_pair(int m_tiles, int n_tiles, int BN, int R, int pairs, int* group_size, int* n_groups) {
  if (R == 0) { *group_size = 1; *n_groups = n_tiles; return; }
  double best = 1e30;
  int best_gs = 1;
  for (int gs = 1; gs <= n_tiles; ++gs) {
    const int groups = (n_tiles + gs - 1) / gs;
    const long items = (long)m_tiles * groups;
    const long rounds = (items + pairs - 1) / pairs;
    const double cost = (double)rounds * (gs * (double)BN + R + 48.0);
    if (cost < best - 1e-9) { best = cost; best_gs = gs; }
  }
  *group_size = best_gs;
  *n_groups = (n_tiles + best_gs - 1) / best_gs;
}

// n_probs problems as independent work items (S == 1), or n_probs SOURCES summed into probs[0].y (S > 1)
template <int BN, int R, int G, int S, bool GEGLU = false, bool DT = false>
static int launch_pair(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, bool f16, cudaStream_t st,
                       void* act_out = nullptr) {
  using C = PairCfg<BN, R, S, DT>;
  static bool attr_set = false;
  if (!attr_set) {
    SDT_CUDA_OK(cudaFuncSetAttribute(lora_gemm_pair_kernel<BN, R, G, S, GEGLU, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  GemmGroup<G> gm;
  for (int q = 0; q < G; ++q) {
    const LoraProblem& pr = probs[q < n_probs ? q : 0];
    int rc = make_tmap_2d_bf16(&gm.x[q], pr.x, M, K, K * 2, C::BM, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.w[q], pr.w, N, K, K * 2, C::HN, C::BK, TMAP_SW_128);     // GEGLU: same map, rows addressed per CTA
    if (rc != SDT_OK) return rc;
    gm.y[q] = reinterpret_cast<uint8_t*>(S > 1 ? probs[0].y : pr.y);
    rc = make_tmap_2d_bf16(&gm.ym[q], gm.y[q], M, N, N * 2, 32, 64, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.ym32[q], gm.y[q], M, N, N * 2, 32, 32, TMAP_SW_64);
    if (rc != SDT_OK) return rc;
    if (R > 0) {
      if (S == 1) rc = make_tmap_2d_bf16(&gm.la[q], pr.la, R, K, K * 2, C::HR, C::BK, TMAP_SW_128);
      else        rc = make_tmap_halves_bf16(&gm.la[q], pr.la, R, K, K * 2, C::LA_ROWS, C::BK);
      if (rc != SDT_OK) return rc;
      rc = make_tmap_2d_bf16(&gm.lb[q], pr.lb, N, R, (uint64_t)R * 2, C::HN, R, R == 64 ? TMAP_SW_128 : (R == 32 ? TMAP_SW_64 : TMAP_SW_32));
      if (rc != SDT_OK) return rc;
    } else {
      gm.la[q] = gm.x[q];
      gm.lb[q] = gm.x[q];
    }
    gm.bias[q] = pr.bias;
    gm.t_out[q] = reinterpret_cast<__nv_bfloat16*>(pr.t_out);
    gm.res[q] = reinterpret_cast<const uint8_t*>(pr.res);
    gm.n[q] = (int)N;
    gm.tile_begin[q] = 0;
  }
  gm.tile_begin[G] = 0;
  PairParams p;
  p.mixed = 0;
  p.scaling = scaling;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.n_probs = S > 1 ? 1 : n_probs;
  p.n_src = S > 1 ? n_probs : 1;
  p.has_bias = probs[0].bias != nullptr ? 1 : 0;
  p.f16 = f16 ? 1 : 0;
  p.geglu_I = GEGLU ? (int)(N / 2) : 0;
  p.act_out = reinterpret_cast<uint8_t*>(act_out);
  p.trace = reinterpret_cast<long long*>(debug_get(10));
  p.ts = (C::A_SLOTS >= 1 && debug_get(30)) ? 1 : 0;
  const int m_tiles = (int)((M + 2 * C::BM - 1) / (2 * C::BM));
  p.n_tiles = (int)((N + BN - 1) / BN);
  const int pairs_max = num_sms() / 2;
  choose_groups_pair(m_tiles * p.n_probs, p.n_tiles, BN, p.n_src * R, pairs_max, &p.group_size, &p.n_groups);
  // Many column tiles (the FF up-projections: 23 / 46 tiles of 224): every tile is its own work item.  Recomputing the rank
  // projection costs R / BN of a merged UMMA, and measured (tools/gemm_ab.py groups) it is 8-13 % faster than sharing it:
  // 8192x640x5120 57.0 -> 49.4 us, 2048x1280x10240 46.6 -> 42.9 us (1270 TF/s).  With few column tiles sharing still wins.
  // The same holds for long K loops (K >= 2048, the FF down-projections and their input gradients): 32768x2560x320 63.2 -> 57.6 us,
  // 8192x5120x640 54.7 -> 51.3 us.  Short K loops with few column tiles keep the shared intermediate.
  if (S == 1 && (p.n_tiles >= 16 || K >= 2048)) { p.group_size = 1; p.n_groups = p.n_tiles; }
  if (debug_get(20) && (int)debug_get(20) <= p.n_tiles) {          // A/B: force the number of n-tiles per work item
    p.group_size = (int)debug_get(20);
    p.n_groups = (p.n_tiles + p.group_size - 1) / p.group_size;
  }
  p.dt_items = DT ? 1 : 0;
  if (DT) {                                                         // every item = two FULL column tiles (the dispatcher checked N % (2 BN) == 0)
    if (N % (2 * BN) != 0) { set_error("lora_gemm_pair(double tile): N = %lld is not a multiple of %d", (long long)N, 2 * BN); return SDT_ERR_UNSUPPORTED; }
    p.group_size = 2;
    p.n_groups = p.n_tiles / 2;
  }
  p.n_items = m_tiles * p.n_probs * p.n_groups;
  const int pairs = p.n_items < pairs_max ? p.n_items : pairs_max;
  SDT_CUDA_OK(launch_kernel(lora_gemm_pair_kernel<BN, R, G, S, GEGLU, DT>, dim3(2 * pairs), dim3(kPairThreads), C::SMEM_BYTES, st, true, gm, p));
  SDT_LAUNCH_OK("lora_gemm_pair");
  return SDT_OK;
}

// CTA-pair entry; same contract as lora_gemm_group_bf16 with main == true (arguments already validated there)
int lora_gemm_pair_group_bf16(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, int r,
                              bool f16, cudaStream_t st) {
  const bool bn160 = (N % 160 == 0) || (N % 128 != 0 && N > 128);
#define SDT_PAIR(BN, R, G) return launch_pair<BN, R, G, 1>(probs, n_probs, scaling, M, K, N, f16, st)
#define SDT_PAIR_R(BN, G)                                                                     \
  switch (r) { case 16: SDT_PAIR(BN, 16, G); case 32: SDT_PAIR(BN, 32, G); default: SDT_PAIR(BN, 64, G); }
  // wide tiles (more FLOP per byte brought into the SM) when the rank accumulators still fit next to two main accumulators
  // (2 * (BW + R) <= 512 TMEM columns: 224-wide up to rank 32, 192-wide at rank 64), the ragged last tile wastes little and there
  // are enough tiles to keep every pair busy: N >= 2048, or N >= 1280 with at least four rounds of tiles (measured: 32768x320x1280
  // 41.3 -> 36.7 us, while 2048x*x1280 -- 48 tiles on 74 pairs -- is faster with 160-wide tiles)
  const int64_t BW = r <= 32 ? 224 : 192;
  const int64_t n_wide = (N + BW - 1) / BW * BW;
  const int64_t tiles_wide = ((M + 255) / 256) * (n_wide / BW);
  const int64_t wide_min_n = debug_get(22) ? (int64_t)debug_get(22) : ((N >= 1280 && tiles_wide >= 4 * (num_sms() / 2)) ? 1280 : 2048);
  // (rank 64 at K = 320: the 192-wide tiles measured 6 % slower than 160-wide ones -- 86.6 vs 81.6 us on 32768x320x2560 -- and 10 % faster from K = 640)
  const bool wide = N >= wide_min_n && n_wide * 100 <= N * 106 && (r <= 32 || K >= 640) && debug_get(12) == 0;
  // double tiles (one joint K loop for the two column tiles of an item, X landed once): long K loops whose items still fill the
  // machine.  sdt_debug_set(31, k): minimum K (default 1280); 1 = never.
  const int64_t dt_min_k = debug_get(31) ? (int64_t)debug_get(31) : 1280;
  const int64_t dt_items = ((M + 255) / 256) * (N / 320);
  if (n_probs == 1 && !wide && bn160 && r > 0 && N % 320 == 0 && dt_min_k > 1 && K >= dt_min_k && dt_items * 10 >= (num_sms() / 2) * 8) {
    switch (r) {
      case 16: return launch_pair<160, 16, 1, 1, false, true>(probs, n_probs, scaling, M, K, N, f16, st);
      case 32: return launch_pair<160, 32, 1, 1, false, true>(probs, n_probs, scaling, M, K, N, f16, st);
      default: return launch_pair<160, 64, 1, 1, false, true>(probs, n_probs, scaling, M, K, N, f16, st);
    }
  }
  if (n_probs == 1) {
    if (wide) { switch (r) { case 0: SDT_PAIR(224, 0, 1); case 16: SDT_PAIR(224, 16, 1); case 32: SDT_PAIR(224, 32, 1); default: SDT_PAIR(192, 64, 1); } }
    if (r == 0) { if (bn160) SDT_PAIR(160, 0, 1); else SDT_PAIR(128, 0, 1); }
    if (bn160) { SDT_PAIR_R(160, 1) } else { SDT_PAIR_R(128, 1) }
  }
  if (bn160) { SDT_PAIR_R(160, kMaxGroup) } else { SDT_PAIR_R(128, kMaxGroup) }
#undef SDT_PAIR_R
#undef SDT_PAIR
}

// Mixed output widths: n_probs projections of ONE (M, K, padded rank) with their own N_q -- to_k / to_v of every cross-attention of
// the UNet on the text context (SD1.5: 32 projections of the 616 x 768 context to 320 / 640 / 1280 columns).  Work items are
// (row tile, problem, column tile); eight launches of 9-15 us at 5-25 % of the tensor peak become one.
bool lora_gemm_pair_mixed_supported(int n_probs, int64_t M, int64_t K, const int64_t* Ns, int r) {
  if (n_probs < 1 || n_probs > kMaxMixed || !(r == 16 || r == 32 || r == 64) || M < 256 || K < 256 || K % 8 != 0) return false;
  for (int q = 0; q < n_probs; ++q)
    if (Ns[q] <= 0 || Ns[q] % 8 != 0 || Ns[q] >= (1ll << 31)) return false;
  return true;
}

template <int R>
static int launch_pair_mixed(const LoraProblem* probs, const int64_t* Ns, int n_probs, float scaling, int64_t M, int64_t K, bool f16,
                             cudaStream_t st) {
  constexpr int BN = 160, G = kMaxMixed;
  using C = PairCfg<BN, R, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    SDT_CUDA_OK(cudaFuncSetAttribute(lora_gemm_pair_kernel<BN, R, G, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  static thread_local GemmGroup<G> gm;           // ~18 KB of tensor maps: not on the caller's stack
  int tiles = 0;
  for (int q = 0; q < G; ++q) {
    const int qq = q < n_probs ? q : 0;
    const LoraProblem& pr = probs[qq];
    const int64_t N = Ns[qq];
    int rc = make_tmap_2d_bf16(&gm.x[q], pr.x, M, K, K * 2, C::BM, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.w[q], pr.w, N, K, K * 2, C::HN, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.la[q], pr.la, R, K, K * 2, C::HR, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.lb[q], pr.lb, N, R, (uint64_t)R * 2, C::HN, R, R == 64 ? TMAP_SW_128 : (R == 32 ? TMAP_SW_64 : TMAP_SW_32));
    if (rc != SDT_OK) return rc;
    gm.y[q] = reinterpret_cast<uint8_t*>(pr.y);
    rc = make_tmap_2d_bf16(&gm.ym[q], pr.y, M, N, N * 2, 32, 64, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&gm.ym32[q], pr.y, M, N, N * 2, 32, 32, TMAP_SW_64);
    if (rc != SDT_OK) return rc;
    gm.bias[q] = pr.bias;
    gm.t_out[q] = reinterpret_cast<__nv_bfloat16*>(pr.t_out);
    gm.res[q] = reinterpret_cast<const uint8_t*>(pr.res);
    gm.n[q] = (int)N;
    gm.tile_begin[q] = tiles;
    if (q < n_probs) tiles += (int)((N + BN - 1) / BN);
  }
  gm.tile_begin[G] = tiles;
  for (int q = n_probs; q <= G; ++q) gm.tile_begin[q] = tiles;
  PairParams p;
  p.mixed = 1;
  p.scaling = scaling;
  p.M = (int)M; p.N = 0; p.K = (int)K;
  p.n_probs = n_probs;
  p.n_src = 1;
  p.has_bias = probs[0].bias != nullptr ? 1 : 0;
  p.f16 = f16 ? 1 : 0;
  p.geglu_I = 0;
  p.act_out = nullptr;
  p.trace = reinterpret_cast<long long*>(debug_get(10));
  p.ts = (C::A_SLOTS >= 1 && debug_get(30)) ? 1 : 0;
  const int m_tiles = (int)((M + 2 * C::BM - 1) / (2 * C::BM));
  p.n_tiles = 0;
  p.group_size = 1;
  p.n_groups = 0;
  p.n_items = m_tiles * tiles;
  const int pairs_max = num_sms() / 2;
  const int pairs = p.n_items < pairs_max ? p.n_items : pairs_max;
  SDT_CUDA_OK(launch_kernel(lora_gemm_pair_kernel<BN, R, G, 1, false>, dim3(2 * pairs), dim3(kPairThreads), C::SMEM_BYTES, st, true, gm, p));
  SDT_LAUNCH_OK("lora_gemm_pair(mixed)");
  return SDT_OK;
}

int lora_gemm_pair_mixed_bf16(const LoraProblem* probs, const int64_t* Ns, int n_probs, float scaling, int64_t M, int64_t K, int r,
                              bool f16, cudaStream_t st) {
  SDT_REQUIRE(probs != nullptr && Ns != nullptr && lora_gemm_pair_mixed_supported(n_probs, M, K, Ns, r), SDT_ERR_UNSUPPORTED,
              "lora_gemm(mixed): needs 1..%d problems, padded rank 16/32/64, M >= 256, K >= 256 (got %d problems, r=%d, M=%lld, K=%lld)",
              kMaxMixed, n_probs, r, (long long)M, (long long)K);
  for (int q = 0; q < n_probs; ++q) {
    const LoraProblem& pr = probs[q];
    SDT_REQUIRE(pr.x && pr.w && pr.la && pr.lb && pr.y && pr.t_out, SDT_ERR_ARG, "lora_gemm(mixed): null pointer in problem %d", q);
    SDT_REQUIRE((pr.bias != nullptr) == (probs[0].bias != nullptr), SDT_ERR_ARG,
                "lora_gemm(mixed): the problems of one launch must all have a bias or all have none");
    SDT_REQUIRE(aligned16(pr.x) && aligned16(pr.w) && aligned16(pr.la) && aligned16(pr.lb) && aligned16(pr.y) && aligned16(pr.t_out) &&
                    aligned16(pr.res), SDT_ERR_ARG, "lora_gemm(mixed): pointers must be 16-byte aligned (problem %d)", q);
  }
  switch (r) {
    case 16: return launch_pair_mixed<16>(probs, Ns, n_probs, scaling, M, K, f16, st);
    case 32: return launch_pair_mixed<32>(probs, Ns, n_probs, scaling, M, K, f16, st);
    default: return launch_pair_mixed<64>(probs, Ns, n_probs, scaling, M, K, f16, st);
  }
}

// GEGLU epilogue (ff.net.0.proj): proj [M, 2I] = X W^T + b + s (X A^T) B^T  AND  act [M, I] = proj[:, :I] * gelu(proj[:, I:]) from ONE
// launch.  Tile nt (256 x 128) holds the h columns [64 nt, 64 nt + 64) and the matching gate columns (CTA 0 of the pair brings the
// h rows of W / lora-up / bias, CTA 1 the gate rows), so the activation is formed from the staged tile and the separate GEGLU
// pass over proj -- read 2I, write I per token -- disappears; proj is still written (the backward needs it), in its reference
// layout.  64 columns = 128 bytes: h, gate and act of a row are each ONE full line (with 80-column halves -- 160-byte runs -- the
// stores ran at half rate: 148 us instead of 60 + 45 for the K = 320 projection, measured).
bool lora_gemm_pair_geglu_supported(int64_t M, int64_t K, int64_t I, int r) {
  return (r == 16 || r == 32 || r == 64) && M >= 256 && K >= 64 && K % 8 == 0 && I % 64 == 0 && I >= 64;
}
int lora_gemm_pair_geglu_bf16(const LoraProblem& pr, void* act_out, float scaling, int64_t M, int64_t K, int64_t I, int r, bool f16,
                              cudaStream_t st) {
  SDT_REQUIRE(lora_gemm_pair_geglu_supported(M, K, I, r), SDT_ERR_UNSUPPORTED,
              "lora_gemm(geglu): needs padded rank 16/32/64, M >= 256, I %% 64 == 0 (got r=%d M=%lld I=%lld)", r, (long long)M, (long long)I);
  SDT_REQUIRE(pr.x && pr.w && pr.la && pr.lb && pr.y && pr.t_out && act_out, SDT_ERR_ARG, "lora_gemm(geglu): null pointer");
  SDT_REQUIRE(aligned16(pr.x) && aligned16(pr.w) && aligned16(pr.la) && aligned16(pr.lb) && aligned16(pr.y) && aligned16(pr.t_out) &&
                  aligned16(act_out), SDT_ERR_ARG, "lora_gemm(geglu): pointers must be 16-byte aligned");
  switch (r) {
    case 16: return launch_pair<128, 16, 1, 1, true>(&pr, 1, scaling, M, K, 2 * I, f16, st, act_out);
    case 32: return launch_pair<128, 32, 1, 1, true>(&pr, 1, scaling, M, K, 2 * I, f16, st, act_out);
    default: return launch_pair<128, 64, 1, 1, true>(&pr, 1, scaling, M, K, 2 * I, f16, st, act_out);
  }
}

// Summed sources: Y = sum_s X_s W_s^T + (scaling X_s la_s^T) lb_s^T  with T_s = scaling X_s la_s^T written to probs[s].t_out.
// The input gradient of projections that read ONE tensor (q / k / v): X_s = dY_s, W_s = W_s^T [K,N], la_s = B_s^T, lb_s = A_s^T.
// 2 or 3 sources, padded rank 16 or 32 (three rank-64 accumulators do not fit in TMEM next to two main accumulators).
bool lora_gemm_pair_sum_supported(int n_src, int64_t M, int64_t K, int64_t N, int r) {
  return n_src >= 2 && n_src <= 3 && (r == 16 || r == 32) && M >= 256 && K >= 64 && K % 8 == 0 && N % 8 == 0;
}
int lora_gemm_pair_sum_bf16(const LoraProblem* probs, int n_src, float scaling, int64_t M, int64_t K, int64_t N, int r, bool f16,
                            cudaStream_t st) {
  SDT_REQUIRE(lora_gemm_pair_sum_supported(n_src, M, K, N, r), SDT_ERR_UNSUPPORTED,
              "lora_gemm(sum): needs 2..3 sources, padded rank 16/32, M >= 256 (got %d sources, r=%d, M=%lld)", n_src, r, (long long)M);
  for (int q = 0; q < n_src; ++q) {
    const LoraProblem& pr = probs[q];
    SDT_REQUIRE(pr.x && pr.w && pr.la && pr.lb && probs[0].y && pr.bias == nullptr, SDT_ERR_ARG, "lora_gemm(sum): bad operands in source %d", q);
    SDT_REQUIRE(aligned16(pr.x) && aligned16(pr.w) && aligned16(pr.la) && aligned16(pr.lb) && aligned16(probs[0].y) && aligned16(pr.t_out),
                SDT_ERR_ARG, "lora_gemm(sum): pointers must be 16-byte aligned (source %d)", q);
  }
  if (r == 16) return launch_pair<160, 16, kMaxGroup, 3>(probs, n_src, scaling, M, K, N, f16, st);
  return launch_pair<160, 32, kMaxGroup, 3>(probs, n_src, scaling, M, K, N, f16, st);
}

}  // namespace sdt
