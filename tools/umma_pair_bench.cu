// Microbenchmark (CTA pair, tcgen05 cta_group::2, M = 256): cost of one UMMA k-step (K = 16) as a function of N and of where the
// A operand comes from:
//   SS     A read from shared memory by the UMMA itself (what lora_gemm_pair_kernel does)
//   TS     A already in tensor memory (lower bound of a TS-mode main loop)
//   CP+TS  A copied shared -> tensor memory by tcgen05.cp.128x256b right before the UMMA that reads it (what a TS-mode main loop
//          fed by TMA would have to do), per k-step or per 64-wide k-block
//   CP     the copies alone
// and a bit-exact check that tcgen05.cp of a 128-byte-swizzled K-major tile produces the TS-mode A layout (D_ss == D_cpts).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scal_sdt_b200/csrc tools/umma_pair_bench.cu -o tools/_build/umma_pair_bench
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>

using namespace sdt::ptx;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// 128 rows x 32 bytes (one K = 16 slice of a K-major bf16 tile) from shared memory to 8 TMEM columns, in both CTAs of the pair
__device__ __forceinline__ void cp2_128x256b(uint32_t taddr, uint64_t s_desc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}
__device__ __forceinline__ void commit2_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

struct RateOut { long long cycles; };
enum { SS = 0, TS = 1, CPTS_STEP = 2, CPTS_BLOCK = 3, CP_ONLY = 4, SS_ALT = 5, N_MODES = 6 };

// smem per CTA: A [128 x 64] bf16 SW128 (16 KiB) | B [128 x 64] bf16 SW128 (16 KiB, the first N/2 rows are this CTA's half).
// MODE is a template argument and every descriptor is formed before the timed loop: the loop body is the UMMAs / copies and nothing
// else.  (A first version chose the mode with run-time branches inside the loop; a lone issuing thread then needed ~113 cycles per
// UMMA for the branches alone and every N up to 224 "cost" 113 cycles -- an artefact of the benchmark, not of the tensor core.)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_rate_kernel(int N, int iters, RateOut* out, float* dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 16384;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  // small integers as bf16: every product and partial sum is exact, so SS and CP+TS must agree bit for bit
  for (int i = threadIdx.x; i < 32768 / 2; i += 128) {
    const uint32_t h = (uint32_t)(i + 17 * rank) * 2654435761u;
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16((float)((int)((h >> 20) % 7u) - 3));
  }
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc2(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_tm = tmem_base + 384;            // A staging: 2 x 32 columns
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    constexpr uint64_t d_sw128 = make_smem_desc_base(16, 1024, kLayoutSW128);
    long long t0 = 0;
    if (elect_one()) {
      uint64_t ad[4], bd[4];
      for (int k = 0; k < 4; ++k) {
        ad[k] = smem_desc(d_sw128, smem_u32(a_s) + k * 32);
        bd[k] = smem_desc(d_sw128, smem_u32(b_s) + k * 32);
      }
      // defined contents for the TS-only mode
      for (int k = 0; k < 4; ++k) cp2_128x256b(a_tm + k * 8, ad[k]);
      for (int k = 0; k < 4; ++k) cp2_128x256b(a_tm + 32 + k * 8, ad[k]);
      if (dump != nullptr) {
        // layout check: ONE k-block, accumulate off on the first UMMA
        if (MODE == CPTS_BLOCK)
          for (int k = 0; k < 4; ++k) cp2_128x256b(a_tm + k * 8, ad[k]);
        for (int k = 0; k < 4; ++k) {
          if (MODE == SS || MODE == SS_ALT) umma2_ss(tmem_base, ad[k], bd[k], idesc, k != 0);
          if (MODE == CPTS_STEP) cp2_128x256b(a_tm + k * 8, ad[k]);
          if (MODE == TS || MODE == CPTS_STEP || MODE == CPTS_BLOCK) umma2_ts(tmem_base, a_tm + k * 8, bd[k], idesc, k != 0);
        }
      } else {
        t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; it += 2) {
          // two k-blocks per trip, A staging slots 0 / 1 in turn
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t slot = a_tm + h * 32;
            if (MODE == CPTS_BLOCK) {
#pragma unroll
              for (int k = 0; k < 4; ++k) cp2_128x256b(slot + k * 8, ad[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (MODE == SS) umma2_ss(tmem_base, ad[k], bd[k], idesc, 1u);
              // two independent accumulators in turn (N <= 176): is the floor a dependent-accumulate latency or an issue interval?
              if (MODE == SS_ALT) umma2_ss(tmem_base + (k & 1) * 176, ad[k], bd[k], idesc, 1u);
              if (MODE == CPTS_STEP || MODE == CP_ONLY) cp2_128x256b(slot + k * 8, ad[k]);
              if (MODE == TS || MODE == CPTS_STEP || MODE == CPTS_BLOCK) umma2_ts(tmem_base, slot + k * 8, bd[k], idesc, 1u);
            }
          }
        }
      }
      commit2_both(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    for (int off = 16; off > 0; off >>= 1) t0 = max(t0, __shfl_xor_sync(0xffffffffu, t0, off));
    if (lane == 0) out[blockIdx.x >> 1].cycles = t1 - t0;
  } else {
    mbar_wait(&done_bar, 0);
  }
  tc_fence_after();
  if (dump != nullptr) {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 8) {
      uint32_t v[8];
      tmem_ld_x8(lane_addr + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 8; ++j) dump[((size_t)rank * 128 + threadIdx.x) * 256 + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

typedef void (*KernelFn)(int, int, RateOut*, float*);
static KernelFn kernel_of(int mode) {
  switch (mode) {
    case SS: return pair_rate_kernel<SS>;
    case TS: return pair_rate_kernel<TS>;
    case CPTS_STEP: return pair_rate_kernel<CPTS_STEP>;
    case CPTS_BLOCK: return pair_rate_kernel<CPTS_BLOCK>;
    case CP_ONLY: return pair_rate_kernel<CP_ONLY>;
    default: return pair_rate_kernel<SS_ALT>;
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  const int smem = 1024 + 32768;
  for (int m = 0; m < N_MODES; ++m) CK(cudaFuncSetAttribute(kernel_of(m), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RateOut* out;
  CK(cudaMalloc(&out, 74 * sizeof(RateOut)));
  float* dump;
  CK(cudaMalloc(&dump, 256 * 256 * 4));
  const char* names[] = {"SS", "TS (A resident)", "CP+TS per k-step", "CP+TS per k-block", "CP only", "SS, 2 accumulators"};
  // ---- layout check: one k-block, accumulate off on the first UMMA
  {
    std::vector<float> ref(256 * 256), got(256 * 256);
    const int N = 160;
    CK(cudaMemset(dump, 0, 256 * 256 * 4));
    kernel_of(SS)<<<2, 128, smem>>>(N, 1, out, dump);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), dump, ref.size() * 4, cudaMemcpyDeviceToHost));
    for (int mode : {CPTS_STEP, CPTS_BLOCK}) {
      CK(cudaMemset(dump, 0, 256 * 256 * 4));
      kernel_of(mode)<<<2, 128, smem>>>(N, 1, out, dump);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(got.data(), dump, got.size() * 4, cudaMemcpyDeviceToHost));
      double maxd = 0, maxref = 0;
      for (int r = 0; r < 256; ++r)
        for (int c = 0; c < N; ++c) {
          maxd = fmax(maxd, fabs((double)ref[r * 256 + c] - got[r * 256 + c]));
          maxref = fmax(maxref, fabs((double)ref[r * 256 + c]));
        }
      printf("layout check %-18s: max|D_ss - D| = %g (max |D_ss| %g) -> %s\n", names[mode], maxd, maxref, maxd == 0 && maxref > 0 ? "OK" : "MISMATCH");
    }
  }
  // ---- rates
  const int iters = 2000;
  for (int grid_pairs : {1, 74})
    for (int N : {16, 32, 64, 96, 128, 160, 176, 208, 224, 240, 256})
      for (int mode = 0; mode < N_MODES; ++mode) {
        if (mode == SS_ALT && N > 176) continue;
        kernel_of(mode)<<<2 * grid_pairs, 128, smem>>>(N, iters, out, nullptr);
        CK(cudaDeviceSynchronize());
        std::vector<RateOut> h(grid_pairs);
        CK(cudaMemcpy(h.data(), out, grid_pairs * sizeof(RateOut), cudaMemcpyDeviceToHost));
        double cyc = 0;
        for (auto& r : h) cyc += r.cycles;
        cyc /= grid_pairs;
        printf("pairs %2d N=%3d %-18s: %7.1f cyc/k-step (floor %5.1f)\n", grid_pairs, N, names[mode], cyc / (iters * 4.0), N / 2.0);
      }
  return 0;
}
