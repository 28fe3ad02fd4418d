// K1 / K2(dX): fused LoRA projection on the 5th-gen tensor cores (sm_100a), bf16 in, f32 accumulate.
//
//     Ts = scaling * (X  A^T)                 [M,R]    rank-R down projection, accumulator in TMEM
//     Y  = X W^T + bias + Ts B^T              [M,N]    frozen base GEMM + rank-R up projection
//
// One kernel serves the forward (X, W, lora_A, lora_B -> Y, Ts; modules/lora.py:12-14 + loralib
// Linear.forward) and the input-gradient half of the backward (dY, W^T, B^T, A^T -> dX, G), because
// dX = dY (W^T)^T + (s dY (B^T)^T) (A^T)^T has exactly the same shape of computation.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: X / W / lora-down k-blocks into a kStages-deep 128B-swizzled smem ring,
//               plus the [BN,R] lora-up tile once per output tile
//   warp 1      tcgen05.mma issuer (one elected lane): main accumulator [128,BN] f32 (double buffered
//               in TMEM) and the rank-R accumulator [128,R]; after the K loop the rank-R intermediate
//               comes back as a bf16 A-operand in shared memory and one more UMMA adds Ts B^T
//   warps 2..5  epilogue: TMEM -> registers -> (+bias, bf16) -> smem transpose -> coalesced 16 B stores;
//               also scale + convert the rank-R intermediate (never leaves the SM except as t_save)
// Work items are (m-tile, n-group); the n-tiles of a group reuse the rank-R intermediate, so the
// down projection is computed once per group, not once per output tile.
#include "sdt_common.cuh"
#include "sm100_ptx.cuh"

#include <mutex>
#include <unordered_map>

namespace sdt {

using namespace ptx;

template <int BN_, int R_>
struct LoraGemmCfg {
  static constexpr int BM = 128, BN = BN_, BK = 64, R = R_;
  static constexpr int kStages = (R_ >= 64) ? 3 : 4;
  static constexpr int X_BYTES = BM * BK * 2;              // 16 KiB
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int LA_BYTES = R * BK * 2;              // lora-down k-block [R,64]
  static constexpr int STAGE_BYTES = X_BYTES + W_BYTES + LA_BYTES;
  static constexpr int LB_BYTES = ((BN * R * 2 + 1023) / 1024) * 1024;   // lora-up tile [BN,R]
  static constexpr int T_BYTES = BM * R * 2;               // rank-R intermediate as UMMA A operand
  static constexpr int STG_ROW = 80;                       // 64 B of payload + 16 B pad: conflict-free transposes
  static constexpr int STG_BYTES = 4 * 32 * STG_ROW;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * STAGE_BYTES + LB_BYTES + T_BYTES + STG_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ACC1_COL = BN, T_COL = 2 * BN;
  static_assert(2 * BN + R <= 512, "TMEM budget");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
  static_assert(R == 0 || R == 16 || R == 32 || R == 64, "rank must be padded to 16/32/64");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct LoraGemmParams {
  const float* bias;      // [N] or null
  __nv_bfloat16* y;       // [M,N] or null when !main
  __nv_bfloat16* t_out;   // [M,R] or null
  float scaling;
  int M, N, K;
  int n_tiles, n_groups, group_size, n_items;
  int main;               // 0: only the rank-R projection is computed (t_out), no base GEMM
};

template <int BN, int R>
__global__ void __launch_bounds__(192, 1)
lora_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_la, const __grid_constant__ CUtensorMap tm_lb,
                 const LoraGemmParams p) {
  using C = LoraGemmCfg<BN, R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lb_smem = smem + C::kStages * C::STAGE_BYTES;
  uint8_t* t_smem = lb_smem + C::LB_BYTES;
  uint8_t* stg_smem = t_smem + C::T_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_smem + C::STG_BYTES);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + C::kStages;         // [kStages]
  uint64_t* acc_full = bars + 2 * C::kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* t_full = acc_empty + 2;
  uint64_t* t_ready = t_full + 1;
  uint64_t* lb_full = t_ready + 1;
  uint64_t* lb_empty = lb_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lb_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (p.K + C::BK - 1) / C::BK;
  const bool has_main = p.main != 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x);
    if (has_main) prefetch_tmap(&tm_w);
    if (R > 0) { prefetch_tmap(&tm_la); if (has_main) prefetch_tmap(&tm_lb); }
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_init(t_full, 1);
    mbar_init(t_ready, 4);
    mbar_init(lb_full, 1);
    mbar_init(lb_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;       // k-block counter across the whole CTA lifetime
      uint32_t tile_ctr = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int m0 = (item / p.n_groups) * C::BM;
        const int g = item % p.n_groups;
        const int nt0 = g * p.group_size;
        const int nt1 = min(nt0 + p.group_size, p.n_tiles);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const bool first = (nt == nt0) && R > 0;
          const int n0 = nt * C::BN;
          const uint32_t tx = C::X_BYTES + (has_main ? C::W_BYTES : 0) + (first ? C::LA_BYTES : 0);
          for (int kb = 0; kb < nk; ++kb, ++it) {
            const int s = it % C::kStages;
            mbar_wait(&empty[s], ((it / C::kStages) & 1) ^ 1);
            uint8_t* st = smem + s * C::STAGE_BYTES;
            mbar_arrive_expect_tx(&full[s], tx);
            tma_load_2d(st, &tm_x, kb * C::BK, m0, &full[s]);
            if (has_main) tma_load_2d(st + C::X_BYTES, &tm_w, kb * C::BK, n0, &full[s]);
            if (first) tma_load_2d(st + C::X_BYTES + C::W_BYTES, &tm_la, kb * C::BK, 0, &full[s]);
          }
          if (R > 0 && has_main) {
            mbar_wait(lb_empty, (tile_ctr & 1) ^ 1);
            mbar_arrive_expect_tx(lb_full, BN * R * 2);
            tma_load_2d(lb_smem, &tm_lb, 0, n0, lb_full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc_main = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t idesc_t = make_idesc_bf16(128, R > 0 ? R : 16, 0, 0);
    constexpr uint64_t d_sw128 = make_smem_desc_base(16, 1024, kLayoutSW128);
    // lora-up tile [BN,R], K-major, rows of R*2 bytes written by TMA with the matching swizzle
    constexpr uint32_t lb_layout = R == 64 ? kLayoutSW128 : (R == 32 ? kLayoutSW64 : kLayoutSW32);
    constexpr uint64_t d_lb = make_smem_desc_base(16, 8 * (R > 0 ? R : 16) * 2, lb_layout);
    // rank-R intermediate [128,R], K-major, un-swizzled core matrices (8 rows x 16 B, 128 B each):
    // K-adjacent cores 128 B apart (LBO), 8-row groups (R/8)*128 B apart (SBO)
    constexpr uint64_t d_t = make_smem_desc_base(128, ((R > 0 ? R : 16) / 8) * 128, kLayoutNone);
    uint32_t it = 0, tile_ctr = 0, item_ctr = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_ctr) {
      const int g = item % p.n_groups;
      const int nt0 = g * p.group_size;
      const int nt1 = min(nt0 + p.group_size, p.n_tiles);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const bool first = (nt == nt0) && R > 0;
        const uint32_t buf = tile_ctr & 1;
        const uint32_t d_main = tmem_base + buf * C::ACC1_COL;
        const uint32_t d_tacc = tmem_base + C::T_COL;
        if (has_main) {
          mbar_wait(&acc_empty[buf], ((tile_ctr >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % C::kStages;
          mbar_wait(&full[s], (it / C::kStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t xa = smem_u32(smem + s * C::STAGE_BYTES);
            const uint32_t wa = xa + C::X_BYTES;
            const uint32_t la = wa + C::W_BYTES;
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) {
              const uint64_t a_desc = smem_desc(d_sw128, xa + k * 32);
              if (has_main) umma_f16_ss(d_main, a_desc, smem_desc(d_sw128, wa + k * 32), idesc_main, (kb | k) != 0);
              if (first) umma_f16_ss(d_tacc, a_desc, smem_desc(d_sw128, la + k * 32), idesc_t, (kb | k) != 0);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (first) {
          if (elect_one()) umma_commit(t_full);
          __syncwarp();
          // without the up-projection below nothing else orders the next item's rank-R MMAs after the
          // epilogue's read of this item's rank-R accumulator
          if (!has_main) mbar_wait(t_ready, item_ctr & 1);
        }
        if (R > 0 && has_main) {
          mbar_wait(lb_full, tile_ctr & 1);
          if (first) mbar_wait(t_ready, item_ctr & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t ta = smem_u32(t_smem), ba = smem_u32(lb_smem);
#pragma unroll
            for (int k = 0; k < R / 16; ++k)
              umma_f16_ss(d_main, smem_desc(d_t, ta + k * 256), smem_desc(d_lb, ba + k * 32), idesc_main, 1u);
            umma_commit(lb_empty);
          }
          __syncwarp();
        }
        if (has_main) {
          if (elect_one()) umma_commit(&acc_full[buf]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================== epilogue warps ====================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* stg = stg_smem + q * 32 * C::STG_ROW;
    uint32_t tile_ctr = 0, item_ctr = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_ctr) {
      const int m0 = (item / p.n_groups) * C::BM;
      const int g = item % p.n_groups;
      const int nt0 = g * p.group_size;
      const int nt1 = min(nt0 + p.group_size, p.n_tiles);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const bool first = (nt == nt0) && R > 0;
        if (first) {
          // ---- rank-R intermediate: TMEM f32 -> scale -> bf16 -> smem A operand (+ t_save) ----
          mbar_wait(t_full, item_ctr & 1);
          tc_fence_after();
          constexpr int RR = R > 0 ? R : 16;
          uint32_t packed[RR / 2];
#pragma unroll
          for (int c = 0; c < RR / 16; ++c) {
            uint32_t v[16];
            tmem_ld_x16(lane_addr + C::T_COL + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              packed[c * 8 + j] = pack_bf16x2(__uint_as_float(v[2 * j]) * p.scaling, __uint_as_float(v[2 * j + 1]) * p.scaling);
          }
          uint8_t* trow = t_smem + (row >> 3) * ((RR / 8) * 128) + (row & 7) * 16;
#pragma unroll
          for (int kc = 0; kc < RR / 8; ++kc)
            *reinterpret_cast<uint4*>(trow + kc * 128) =
                make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
          if (p.t_out != nullptr && g == 0 && m0 + row < p.M) {
            uint4* dst = reinterpret_cast<uint4*>(p.t_out + (size_t)(m0 + row) * RR);
#pragma unroll
            for (int kc = 0; kc < RR / 8; ++kc)
              dst[kc] = make_uint4(packed[kc * 4], packed[kc * 4 + 1], packed[kc * 4 + 2], packed[kc * 4 + 3]);
          }
          fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_ready);
        }
        if (!has_main) continue;
        // ---- main accumulator: TMEM -> (+bias) -> bf16 -> transpose through smem -> global ----
        const uint32_t buf = tile_ctr & 1;
        mbar_wait(&acc_full[buf], (tile_ctr >> 1) & 1);
        tc_fence_after();
        const int n0 = nt * C::BN;
#pragma unroll 1
        for (int c = 0; c < C::BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;                // warp-uniform
          uint32_t v[32];
          tmem_ld_x32(lane_addr + buf * C::ACC1_COL + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int cc = col0 + 2 * j;
              const float b0 = cc < p.N ? __ldg(p.bias + cc) : 0.f, b1 = cc + 1 < p.N ? __ldg(p.bias + cc + 1) : 0.f;
              pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + b0, __uint_as_float(v[2 * j + 1]) + b1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          }
          uint8_t* srow = stg + lane * C::STG_ROW;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(srow + j * 16) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          __syncwarp();
          // lane l stores the 16 B chunk (l & 3) of rows (l >> 2) + 8 j: each row segment is 64 contiguous bytes
          const int ch = lane & 3;
          const int colc = col0 + ch * 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = (lane >> 2) + 8 * j;
            const int grow = m0 + q * 32 + r;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * C::STG_ROW + ch * 16);
            if (grow < p.M && colc < p.N)
              *reinterpret_cast<uint4*>(p.y + (size_t)grow * p.N + colc) = val;
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
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

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, TmapSwizzle swz) {
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
  return SDT_OK;
}

template <int BN, int R>
static int launch_lora_gemm(const void* x, const void* w, const float* bias, const void* la, const void* lb, float scaling,
                            void* y, void* t_out, int64_t M, int64_t K, int64_t N, bool main, cudaStream_t st) {
  using C = LoraGemmCfg<BN, R>;
  static bool attr_set = false;
  if (!attr_set) {
    SDT_CUDA_OK(cudaFuncSetAttribute(lora_gemm_kernel<BN, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tm_x, tm_w, tm_la, tm_lb;
  int rc = make_tmap_2d_bf16(&tm_x, x, M, K, K * 2, C::BM, C::BK, TMAP_SW_128);
  if (rc != SDT_OK) return rc;
  if (main) {
    rc = make_tmap_2d_bf16(&tm_w, w, N, K, K * 2, BN, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
  } else {
    tm_w = tm_x;
  }
  if (R > 0) {
    rc = make_tmap_2d_bf16(&tm_la, la, R, K, K * 2, R, C::BK, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    if (main) {
      rc = make_tmap_2d_bf16(&tm_lb, lb, N, R, (uint64_t)R * 2, BN, R,
                             R == 64 ? TMAP_SW_128 : (R == 32 ? TMAP_SW_64 : TMAP_SW_32));
      if (rc != SDT_OK) return rc;
    } else {
      tm_lb = tm_la;
    }
  } else {
    tm_la = tm_x;
    tm_lb = tm_x;
  }
  LoraGemmParams p;
  p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.t_out = reinterpret_cast<__nv_bfloat16*>(t_out);
  p.scaling = scaling;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.main = main ? 1 : 0;
  const int m_tiles = (int)((M + C::BM - 1) / C::BM);
  p.n_tiles = main ? (int)((N + BN - 1) / BN) : 1;
  // n-groups: as few as possible (the rank-R projection is recomputed once per group) while still
  // giving every SM a couple of work items
  const int sms = num_sms();
  int n_groups = 1;
  if (R > 0) {
    while (n_groups < p.n_tiles && m_tiles * n_groups < 2 * sms) ++n_groups;
  } else {
    n_groups = p.n_tiles;
  }
  p.group_size = (p.n_tiles + n_groups - 1) / n_groups;
  p.n_groups = (p.n_tiles + p.group_size - 1) / p.group_size;
  p.n_items = m_tiles * p.n_groups;
  const int grid = p.n_items < sms ? p.n_items : sms;
  lora_gemm_kernel<BN, R><<<grid, 192, C::SMEM_BYTES, st>>>(tm_x, tm_w, tm_la, tm_lb, p);
  SDT_LAUNCH_OK("lora_gemm");
  return SDT_OK;
}

// bf16 entry used by sdt_lora_linear_fwd / sdt_lora_linear_bwd (api in lora_api.cu)
int lora_gemm_bf16(const void* x, const void* w, const float* bias, const void* la, const void* lb, float scaling, void* y,
                   void* t_out, int64_t M, int64_t K, int64_t N, int r, bool main, cudaStream_t st) {
  SDT_REQUIRE(M > 0 && K > 0 && N > 0, SDT_ERR_ARG, "lora_gemm: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
  SDT_REQUIRE(M < (1ll << 31) && K < (1ll << 31) && N < (1ll << 31), SDT_ERR_UNSUPPORTED, "lora_gemm: dimension exceeds int32");
  SDT_REQUIRE(K % 8 == 0 && N % 8 == 0, SDT_ERR_UNSUPPORTED, "lora_gemm: K and N must be multiples of 8 (K=%lld N=%lld)", (long long)K, (long long)N);
  SDT_REQUIRE(r == 0 || r == 16 || r == 32 || r == 64, SDT_ERR_UNSUPPORTED,
              "lora_gemm: padded rank must be 0, 16, 32 or 64 (got %d)", r);
  SDT_REQUIRE(main || r > 0, SDT_ERR_ARG, "lora_gemm: nothing to compute");
  const bool bn160 = !main || (N % 160 == 0) || (N % 128 != 0 && N > 128);
#define SDT_GEMM(BN, R) return launch_lora_gemm<BN, R>(x, w, bias, la, lb, scaling, y, t_out, M, K, N, main, st)
  if (bn160) {
    switch (r) { case 0: SDT_GEMM(160, 0); case 16: SDT_GEMM(160, 16); case 32: SDT_GEMM(160, 32); default: SDT_GEMM(160, 64); }
  } else {
    switch (r) { case 0: SDT_GEMM(128, 0); case 16: SDT_GEMM(128, 16); case 32: SDT_GEMM(128, 32); default: SDT_GEMM(128, 64); }
  }
#undef SDT_GEMM
}

}  // namespace sdt
