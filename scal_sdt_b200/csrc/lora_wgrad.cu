// K2 (dA, dB): LoRA weight gradients = reductions over the token dimension on the tensor cores.
//
//     dA[j,k] += sum_m G[m,j]  X[m,k]         (G = scaling * dY B, written by the dX kernel)
//     dB[n,j] += sum_m dY[m,n] Ts[m,j]        (Ts = scaling * X A^T, saved by the forward)
//
// Both are C[f,j] = sum_m U[m,f] V[m,j] with U = X or dY ([M,F], token-major as it lies in HBM) and
// V = G or Ts ([M,R]).  The contraction runs over rows, so both operands are "MN-major" for UMMA: the
// TMA boxes are taken from U and V exactly as stored (no transposes anywhere) and the MMA reads them
// through MN-major shared-memory descriptors.  A CTA owns a 128-feature block and one slice of the
// token range.  The partial sums of the token slices are combined DETERMINISTICALLY: every CTA writes its [128 x R] partial to
// a caller-provided workspace, and the CTA that arrives last on its feature block (a counter per block) adds all slices in
// slice order and accumulates the total into dA / dB -- one launch, run-to-run bit-identical gradients, like the
// reference's torch path.  Without a workspace the partials leave through f32 red.global.add (order not fixed).
//
// HBM-bound by construction: every element of U is read once for R multiply-adds.
#include "sdt_common.cuh"
#include "sm100_ptx.cuh"

namespace sdt {

using namespace ptx;

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, TmapSwizzle swz);

template <int R_>
struct WgradCfg {
  static constexpr int R = R_, BF = 128, BMK = 64;          // 128 features x 64 tokens per stage
  static constexpr int kStages = 6;
  static constexpr int U_BYTES = BMK * BF * 2;              // two [64 x 64] boxes, 8 KiB each
  static constexpr int V_BYTES = ((BMK * R * 2 + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = U_BYTES + V_BYTES;
  static constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + 256;
  static constexpr int TMEM_COLS = R < 32 ? 32 : R;
};

// One launch covers the dA and dB reductions of up to kWgradSites sites OF ANY SHAPES (a whole transformer block: q / k / v /
// out of both attentions, the two feed-forward projections, proj_in / proj_out -- or just one site): 2 sub-problems per
// site, each with its own token count, feature width, tensor maps and split of the token range.  The grid is 1-D; a CTA finds
// its sub-problem by scanning the prefix table, then (feature block, token slice) = (local / n_splits, local % n_splits).  The
// descriptors sit in one __grid_constant__ block (about 9 KB; kernel parameters may be 32 KB since CUDA 12.1).
constexpr int kWgradSites = 16, kWgradSubs = 2 * kWgradSites;
struct alignas(64) WgradSub {
  CUtensorMap u, v;
  float* out;              // dA [r_true,F] (transposed = 1) or dB [F,r_true] (transposed = 0)
  int F, transposed;
  int M;                   // token rows of this sub-problem
  int rows_per_split;      // multiple of 64
  int n_splits;
  int cta_begin;           // first blockIdx.x of this sub-problem
  int blk_begin;           // index of its first feature block in the per-launch counter array
  int pad;
};
struct WgradBatch {
  WgradSub sub[kWgradSubs];
  int n_subs;
};
// deterministic mode workspace: [kWgradMaxBlocks counters][partials: one [R x 128] f32 slab per CTA]
constexpr int kWgradMaxBlocks = 4096, kWgradMaxCtas = 1024;
constexpr size_t kWgradWorkspaceBytes = (size_t)kWgradMaxBlocks * 4 + (size_t)kWgradMaxCtas * 64 * 128 * 4;

struct WgradParams {
  int r_true;
  unsigned int* counters; // deterministic mode: arrivals per feature block (left at zero); null = atomics
  float* partial;         // deterministic mode: [CTA][R][128]
  // descriptors are host-built so that the layout constants live in one place (and can be probed)
  uint64_t a_desc_base, b_desc_base;
  uint32_t a_step, b_step, idesc;
};

template <int R>
__global__ void __launch_bounds__(128, 1)
lora_wgrad_kernel(const __grid_constant__ WgradBatch gw, const WgradParams p) {
  using C = WgradCfg<R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* done = bars + 2 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int which = 0;
  while (which + 1 < gw.n_subs && (int)blockIdx.x >= gw.sub[which + 1].cta_begin) ++which;
  const WgradSub& sb = gw.sub[which];
  float* const out = sb.out;
  const int F = sb.F, transposed = sb.transposed;
  const CUtensorMap* tm_u = &sb.u;
  const CUtensorMap* tm_v = &sb.v;
  const int local = (int)blockIdx.x - sb.cta_begin;
  const int n_splits = sb.n_splits;
  const int fb = local / n_splits, split = local - fb * n_splits;     // neighbouring CTAs: the slices of one feature block
  const int f0 = fb * C::BF;
  const int m_begin = split * sb.rows_per_split;
  const int m_end = min(m_begin + sb.rows_per_split, sb.M);
  const int nk = m_end > m_begin ? (m_end - m_begin + C::BMK - 1) / C::BMK : 0;
  const int first_cta = sb.cta_begin + fb * n_splits;                  // slab of slice 0 of this feature block
  unsigned int* const counter = p.counters != nullptr ? p.counters + sb.blk_begin + fb : nullptr;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(tm_u);
    prefetch_tmap(tm_v);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // PDL: U / V come from the kernels in front of us; the prologue above did not need them
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % C::kStages;
        mbar_wait(&empty[s], ((kb / C::kStages) & 1) ^ 1);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        const int m = m_begin + kb * C::BMK;
        mbar_arrive_expect_tx(&full[s], C::U_BYTES + C::BMK * R * 2);
        tma_load_2d(st, tm_u, f0, m, &full[s]);                    // features f0 .. f0+63
        tma_load_2d(st + C::U_BYTES / 2, tm_u, f0 + 64, m, &full[s]);   // features f0+64 .. f0+127
        tma_load_2d(st + C::U_BYTES, tm_v, 0, m, &full[s]);
      }
    }
  } else if (warp == 1) {
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % C::kStages;
      mbar_wait(&full[s], (kb / C::kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ua = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t va = ua + C::U_BYTES;
#pragma unroll
        for (int k = 0; k < C::BMK / 16; ++k)
          umma_f16_ss(tmem_base, smem_desc(p.a_desc_base, ua + k * p.a_step), smem_desc(p.b_desc_base, va + k * p.b_step),
                      p.idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  }

  // ---- epilogue: all four warps, thread = feature row ----
  const int tid = warp * 32 + lane;
  const int f = f0 + tid;
  if (p.partial == nullptr) {
    if (nk > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c = 0; c < R / 16; ++c) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c * 16, v);
        tmem_ld_wait();
        if (f < F) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int jj = c * 16 + j;
            if (jj < p.r_true) {
              float* dst = transposed ? out + (size_t)jj * F + f : out + (size_t)f * p.r_true + jj;
              atomicAdd(dst, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  } else {
    // stage 1: this CTA's partial [R][128 features] (features contiguous: coalesced both ways)
    float* mine = p.partial + (size_t)blockIdx.x * (R * 128);
    if (nk > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < R / 16; ++c) {
      uint32_t v[16];
      if (nk > 0) {
        tmem_ld_x16(taddr + c * 16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < p.r_true) mine[(c * 16 + j) * 128 + tid] = __uint_as_float(v[j]);
    }
    // stage 2: the last CTA to arrive on this feature block sums the slices in slice order and accumulates into dA / dB
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == (unsigned int)n_splits - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last != 0u) {
      __threadfence();
      // The slab of one slice is [R][128] f32 = R * 32 float4; thread t owns float4 number t + 128 i (rank row j = i * 4 + t / 32,
      // features 4 (t % 32) .. + 3).  All of a thread's loads of 4 consecutive slices are issued before the first add (4 / 2 / 1 slices at rank 16 / 32 / 64), so the
      // ~200 KB of partials stream in at L2 bandwidth instead of one dependent load at a time; the slices are still added in
      // slice order, which is what makes the result reproducible.
      constexpr int V = R / 4;                       // float4 per thread per slice
      constexpr int U = R <= 16 ? 4 : (R == 32 ? 2 : 1);   // slices in flight (register budget: U * V float4)
      const float4* base = reinterpret_cast<const float4*>(p.partial + (size_t)first_cta * (R * 128)) + tid;
      float4 acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const unsigned int n_sl = (unsigned int)n_splits;
      unsigned int sl = 0;
      for (; sl + U <= n_sl; sl += U) {
        float4 v[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < V; ++i)
            if (i * 4 < p.r_true) v[u][i] = __ldcg(base + (size_t)(sl + u) * (R * 32) + i * 128);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < V; ++i)
            if (i * 4 < p.r_true) { acc[i].x += v[u][i].x; acc[i].y += v[u][i].y; acc[i].z += v[u][i].z; acc[i].w += v[u][i].w; }
      }
      for (; sl < n_sl; ++sl) {
#pragma unroll
        for (int i = 0; i < V; ++i)
          if (i * 4 < p.r_true) {
            const float4 v = __ldcg(base + (size_t)sl * (R * 32) + i * 128);
            acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
          }
      }
      const int fq = f0 + 4 * (tid & 31);            // first of this thread's four features
      // accumulate into dA / dB: ALL old values are loaded before the first store (written as `+=` per element the compiler
      // must keep every load behind the previous store -- 16 dependent L2 round trips, ~15 us per launch, measured)
      if (fq < F) {                                  // F % 8 == 0 and fq % 4 == 0: the four features are in range together
        if (transposed) {
          float4 o[V];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int j = i * 4 + (tid >> 5);
            if (j < p.r_true) o[i] = __ldcg(reinterpret_cast<const float4*>(out + (size_t)j * F + fq));
          }
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int j = i * 4 + (tid >> 5);
            if (j < p.r_true)
              *reinterpret_cast<float4*>(out + (size_t)j * F + fq) =
                  make_float4(o[i].x + acc[i].x, o[i].y + acc[i].y, o[i].z + acc[i].z, o[i].w + acc[i].w);
          }
        } else {
          float o[V][4];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int j = i * 4 + (tid >> 5);
            if (j < p.r_true) {
#pragma unroll
              for (int q = 0; q < 4; ++q) o[i][q] = __ldcg(out + (size_t)(fq + q) * p.r_true + j);
            }
          }
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int j = i * 4 + (tid >> 5);
            if (j < p.r_true) {
              out[(size_t)(fq + 0) * p.r_true + j] = o[i][0] + acc[i].x;
              out[(size_t)(fq + 1) * p.r_true + j] = o[i][1] + acc[i].y;
              out[(size_t)(fq + 2) * p.r_true + j] = o[i][2] + acc[i].z;
              out[(size_t)(fq + 3) * p.r_true + j] = o[i][3] + acc[i].w;
            }
          }
        }
      }
      if (threadIdx.x == 0) *counter = 0u;           // ready for the next launch on this workspace
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// debug overrides (sdt_debug_set): lets the GPU tests probe descriptor hypotheses without a rebuild
static uint64_t g_dbg[32] = {0};
void debug_set(int key, uint64_t value) { if (key >= 0 && key < 32) g_dbg[key] = value; }
uint64_t debug_get(int key) { return (key >= 0 && key < 32) ? g_dbg[key] : 0; }

struct WgradSite {       // one LoRA site: dA[j,k] += sum_m G[m,j] X[m,k]  and  dB[n,j] += sum_m dY[m,n] Ts[m,j]
  const void* x; const void* g; float* dA;
  const void* dy; const void* ts; float* dB;
  int64_t M, K, N;
};

template <int R>
static int launch_wgrad(const WgradSite* sites, int n_sites, int r_true, void* ws, bool f16, cudaStream_t st) {
  using C = WgradCfg<R>;
  static bool attr_set = false;
  if (!attr_set) {
    SDT_CUDA_OK(cudaFuncSetAttribute(lora_wgrad_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const TmapSwizzle vsw = R == 64 ? TMAP_SW_128 : (R == 32 ? TMAP_SW_64 : TMAP_SW_32);
  static thread_local WgradBatch gw;            // ~9 KB: keep it off the stack of the caller's thread
  gw.n_subs = 2 * n_sites;
  // work units (128 features x 64 tokens) of every sub-problem
  long total_units = 0;
  for (int q = 0; q < gw.n_subs; ++q) {
    const WgradSite& site = sites[q / 2];
    const int64_t F = (q & 1) == 0 ? site.K : site.N;
    total_units += ((F + C::BF - 1) / C::BF) * ((site.M + C::BMK - 1) / C::BMK);
  }
  // A CTA streams at most `cpc` 64-token chunks of its feature block: enough CTAs for `per_sm` per SM over the whole launch, at
  // least `min_chunks` stages of work each (tools/wgrad_ab.py: one CTA per SM beats two on the 22-190 MB shapes -- fewer
  // partials -- and ties on the small ones).  sdt_debug_set(15, min | per_sm << 8) overrides.
  const int min_chunks = (g_dbg[15] & 0xff) ? (int)(g_dbg[15] & 0xff) : 8;
  const int per_sm = (g_dbg[15] >> 8) ? (int)(g_dbg[15] >> 8) : 1;
  long cpc = (total_units + (long)per_sm * num_sms() - 1) / ((long)per_sm * num_sms());
  if (cpc < min_chunks) cpc = min_chunks;
  int ctas = 0, blocks = 0;
  for (int q = 0; q < kWgradSubs; ++q) {
    WgradSub& sb = gw.sub[q];
    const WgradSite& site = sites[q < gw.n_subs ? q / 2 : 0];          // unused slots repeat site 0 (never selected)
    const bool is_dA = (q & 1) == 0;
    const void* u = is_dA ? site.x : site.dy;
    const void* v = is_dA ? site.g : site.ts;
    const int64_t F = is_dA ? site.K : site.N;
    int rc = make_tmap_2d_bf16(&sb.u, u, site.M, F, F * 2, C::BMK, 64, TMAP_SW_128);
    if (rc != SDT_OK) return rc;
    rc = make_tmap_2d_bf16(&sb.v, v, site.M, R, (uint64_t)R * 2, C::BMK, R, vsw);
    if (rc != SDT_OK) return rc;
    sb.out = is_dA ? site.dA : site.dB;
    sb.F = (int)F;
    sb.transposed = is_dA ? 1 : 0;
    sb.M = (int)site.M;
    const int m_chunks = (int)((site.M + C::BMK - 1) / C::BMK);
    int splits = (int)((m_chunks + cpc - 1) / cpc);
    const int chunks_per_split = (m_chunks + splits - 1) / splits;
    splits = (m_chunks + chunks_per_split - 1) / chunks_per_split;      // no empty slices
    sb.rows_per_split = chunks_per_split * C::BMK;
    sb.n_splits = splits;
    sb.cta_begin = ctas;
    sb.blk_begin = blocks;
    sb.pad = 0;
    if (q < gw.n_subs) {
      const int f_blocks = (int)((F + C::BF - 1) / C::BF);
      ctas += f_blocks * splits;
      blocks += f_blocks;
    }
  }
  WgradParams p;
  p.r_true = r_true;
  // A = U tile, MN-major, 128B swizzle: 64-feature blocks 8 KiB apart (LBO), 8-token groups 1 KiB apart (SBO),
  //     16 tokens per MMA -> 2 KiB per k-step
  // B = V tile, MN-major, rows of R*2 bytes with the matching swizzle: 8-token groups 8*R*2 B apart (SBO)
  const uint32_t vlayout = R == 64 ? kLayoutSW128 : (R == 32 ? kLayoutSW64 : kLayoutSW32);
  p.a_desc_base = make_smem_desc_base(C::U_BYTES / 2, 1024, kLayoutSW128);
  p.b_desc_base = make_smem_desc_base(8 * R * 2, 8 * R * 2, vlayout);
  p.a_step = 16 * 128;
  p.b_step = 16 * R * 2;
  p.idesc = idesc_operand_format(make_idesc_bf16(128, R, 1, 1), f16);
  if (g_dbg[0]) p.a_desc_base = g_dbg[1];
  if (g_dbg[2]) p.b_desc_base = g_dbg[3];
  if (g_dbg[4]) p.a_step = (uint32_t)g_dbg[5];
  if (g_dbg[6]) p.b_step = (uint32_t)g_dbg[7];
  if (g_dbg[8]) p.idesc = (uint32_t)g_dbg[9];
  p.counters = nullptr;
  p.partial = nullptr;
  if (ws != nullptr) {
    SDT_REQUIRE(blocks <= kWgradMaxBlocks && ctas <= kWgradMaxCtas, SDT_ERR_UNSUPPORTED,
                "lora_wgrad: %d feature blocks / %d CTAs exceed the deterministic workspace (%d / %d): launch fewer sites at once",
                blocks, ctas, kWgradMaxBlocks, kWgradMaxCtas);
    p.counters = reinterpret_cast<unsigned int*>(ws);
    p.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + (size_t)kWgradMaxBlocks * 4);
  }
  SDT_CUDA_OK(launch_kernel(lora_wgrad_kernel<R>, dim3(ctas), dim3(128), C::SMEM_BYTES, st, true, gw, p));
  SDT_LAUNCH_OK("lora_wgrad");
  return SDT_OK;
}

size_t lora_wgrad_workspace_bytes() { return kWgradWorkspaceBytes; }
int lora_wgrad_max_sites() { return kWgradSites; }

// dA / dB of n_sites sites (1..kWgradSites) of ANY shapes, one padded rank, in ONE launch
int lora_wgrad_batch_bf16(const WgradSite* sites, int n_sites, int r, int r_true, void* ws, bool f16, cudaStream_t st) {
  SDT_REQUIRE(sites != nullptr && n_sites >= 1 && n_sites <= kWgradSites, SDT_ERR_ARG, "lora_wgrad: 1..%d sites per launch (got %d)",
              kWgradSites, n_sites);
  for (int q = 0; q < n_sites; ++q) {
    const WgradSite& t = sites[q];
    SDT_REQUIRE(t.x && t.g && t.dA && t.dy && t.ts && t.dB, SDT_ERR_ARG, "lora_wgrad: null pointer (site %d)", q);
    SDT_REQUIRE(t.M > 0 && t.K > 0 && t.N > 0 && t.K % 8 == 0 && t.N % 8 == 0 && t.M < (1ll << 31), SDT_ERR_ARG,
                "lora_wgrad: bad sizes M=%lld K=%lld N=%lld (site %d)", (long long)t.M, (long long)t.K, (long long)t.N, q);
    SDT_REQUIRE(aligned16(t.x) && aligned16(t.g) && aligned16(t.dy) && aligned16(t.ts), SDT_ERR_ARG,
                "lora_wgrad: operands must be 16-byte aligned (site %d)", q);
  }
  SDT_REQUIRE(r_true >= 1 && r_true <= r, SDT_ERR_ARG, "lora_wgrad: r_true=%d outside [1,%d]", r_true, r);
  SDT_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15u) == 0, SDT_ERR_ARG, "lora_wgrad: workspace must be 16-byte aligned");
  switch (r) {
    case 16: return launch_wgrad<16>(sites, n_sites, r_true, ws, f16, st);
    case 32: return launch_wgrad<32>(sites, n_sites, r_true, ws, f16, st);
    case 64: return launch_wgrad<64>(sites, n_sites, r_true, ws, f16, st);
  }
  set_error("lora_wgrad: padded rank must be 16, 32 or 64 (got %d)", r);
  return SDT_ERR_UNSUPPORTED;
}

}  // namespace sdt
