// Microbenchmark: tcgen05.mma issue rate on sm_100a as a function of N, operand source (A from shared memory "SS" vs A
// from tensor memory "TS") and concurrent shared-memory fill traffic (bulk copies, as the TMA producer generates).
// Also checks the TMEM layout of a TS-mode A operand numerically against the SS result.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scal_sdt_b200/csrc tools/umma_bench.cu -o tools/_build/umma_bench
//   tools/_build/umma_bench            (on the GPU box)
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>

using namespace sdt::ptx;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// rate benchmark.  smem: A [128 x 64] bf16 SW128 (16 KiB) | B [256 x 64] bf16 SW128 (32 KiB) | fill scratch 2 x 32 KiB
// mode 0: SS   mode 1: TS (A operand at TMEM column 384 + 8 k)   fill: bytes of bulk copy per "k-block" of 4 UMMAs (0 = none)
struct RateOut { long long cycles; long long ns; };

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n2, int mode, int iters, int fill_bytes, const uint8_t* gsrc, RateOut* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 16384;
  uint8_t* fill_s = smem + 16384 + 32768;
  __shared__ uint64_t done_bar, fill_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i * 2654435761u & 0x007F007Fu);
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&fill_bar[0], 1);
    mbar_init(&fill_bar[1], 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t idesc2 = n2 > 0 ? make_idesc_bf16(128, n2, 0, 0) : 0;
    constexpr uint64_t d_sw128 = make_smem_desc_base(16, 1024, kLayoutSW128);
    long long t0 = 0, t1 = 0;
    unsigned long long g0 = 0, g1 = 0;
    if (elect_one()) {
      t0 = clock64();
      g0 = gtime();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bd = smem_desc(d_sw128, smem_u32(b_s) + k * 32);
          if (mode == 0) {
            const uint64_t ad = smem_desc(d_sw128, smem_u32(a_s) + k * 32);
            umma_f16_ss(tmem_base, ad, bd, idesc, 1u);
            if (n2 > 0) umma_f16_ss(tmem_base + 256, ad, bd, idesc2, 1u);
          } else {
            umma_f16_ts(tmem_base, tmem_base + 384 + k * 8, bd, idesc, 1u);
            if (n2 > 0) umma_f16_ts(tmem_base + 256, tmem_base + 384 + k * 8, bd, idesc2, 1u);
          }
        }
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    if (lane == 0) {
      // elect_one picks one lane; its timestamps were taken in that lane -- recompute here conservatively
    }
    t1 = clock64();
    g1 = gtime();
    // reduce: the elected lane holds t0/g0
    for (int off = 16; off > 0; off >>= 1) {
      t0 = max(t0, __shfl_xor_sync(0xffffffffu, t0, off));
      g0 = max(g0, __shfl_xor_sync(0xffffffffu, g0, off));
    }
    if (lane == 0) {
      out[blockIdx.x].cycles = t1 - t0;
      out[blockIdx.x].ns = (long long)(g1 - g0);
      stop_flag = 1;
    }
  } else if (warp == 1 && fill_bytes > 0) {
    // emulate the TMA producer: bulk copies global -> shared, two in flight
    if (lane == 0) {
      uint32_t n = 0;
      const size_t span = 64u << 20;
      size_t off = (size_t)blockIdx.x * 262144;
      while (!stop_flag) {
        const int b = n & 1;
        if (n >= 2) mbar_wait(&fill_bar[b], ((n >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&fill_bar[b], fill_bytes);
        bulk_g2s(fill_s + b * 32768, gsrc + (off % span), fill_bytes, &fill_bar[b]);
        off += fill_bytes;
        ++n;
      }
      // drain
      if (n >= 1) mbar_wait(&fill_bar[(n - 1) & 1], ((n - 1) >> 1) & 1);
      if (n >= 2) mbar_wait(&fill_bar[(n - 2) & 1], ((n - 2) >> 1) & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// layout check: D_ss = A B^T with A in shared memory vs D_ts with A written to TMEM by tcgen05.st (lane = row, 32-bit
// column c = (A[m][2c], A[m][2c+1])).  A [128 x 16], B [16 x 16], canonical un-swizzled K-major core matrices.
__global__ void __launch_bounds__(128, 1) ts_check_kernel(float* d_ss, float* d_ts) {
  __shared__ __align__(1024) uint8_t a_s[128 * 32];
  __shared__ __align__(1024) uint8_t b_s[16 * 32];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = threadIdx.x;
  auto aval = [](int r, int k) { return (float)(((r * 7 + k * 3) % 13) - 6); };
  auto bval = [](int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); };
  for (int k = 0; k < 16; ++k) {
    *reinterpret_cast<__nv_bfloat16*>(a_s + (m >> 3) * 256 + (k >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2) = __float2bfloat16(aval(m, k));
    if (m < 16) *reinterpret_cast<__nv_bfloat16*>(b_s + (m >> 3) * 256 + (k >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2) = __float2bfloat16(bval(m, k));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  // A -> TMEM columns [128, 136)
  uint32_t regs[8];
  for (int c = 0; c < 8; ++c) {
    __nv_bfloat162 v = __floats2bfloat162_rn(aval(m, 2 * c), aval(m, 2 * c + 1));
    regs[c] = *reinterpret_cast<uint32_t*>(&v);
  }
  tmem_st_x8(lane_addr + 128, regs);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint64_t d_none = make_smem_desc_base(128, 256, kLayoutNone);
      const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
      umma_f16_ss(tmem_base + 0, smem_desc(d_none, smem_u32(a_s)), smem_desc(d_none, smem_u32(b_s)), idesc, 0u);
      umma_f16_ts(tmem_base + 32, tmem_base + 128, smem_desc(d_none, smem_u32(b_s)), idesc, 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v0[16], v1[16];
  tmem_ld_x16(lane_addr + 0, v0);
  tmem_ld_x16(lane_addr + 32, v1);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) {
    d_ss[m * 16 + j] = __uint_as_float(v0[j]);
    d_ts[m * 16 + j] = __uint_as_float(v1[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  // ---- TS layout check
  {
    float *d_ss, *d_ts;
    CK(cudaMalloc(&d_ss, 128 * 16 * 4));
    CK(cudaMalloc(&d_ts, 128 * 16 * 4));
    ts_check_kernel<<<1, 128>>>(d_ss, d_ts);
    CK(cudaDeviceSynchronize());
    std::vector<float> hs(2048), ht(2048);
    CK(cudaMemcpy(hs.data(), d_ss, 8192, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ht.data(), d_ts, 8192, cudaMemcpyDeviceToHost));
    double maxd = 0, maxref = 0, refd = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)(((m * 7 + k * 3) % 13) - 6) * (double)(((n * 5 + k * 11) % 9) - 4);
        maxd = fmax(maxd, fabs(hs[m * 16 + n] - ht[m * 16 + n]));
        refd = fmax(refd, fabs(hs[m * 16 + n] - ref));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("TS layout check: max|D_ss - D_ts| = %g   max|D_ss - ref| = %g  (max |ref| %g)  -> %s\n", maxd, refd, maxref,
           (maxd == 0 && refd == 0) ? "TS A layout = lane row, column c = (k=2c, 2c+1): OK" : "MISMATCH");
  }
  // ---- rates
  uint8_t* gsrc;
  CK(cudaMalloc(&gsrc, (64u << 20) + (1u << 20) + 148u * 262144));
  CK(cudaMemset(gsrc, 1, (64u << 20) + (1u << 20)));
  RateOut* out;
  CK(cudaMalloc(&out, 148 * sizeof(RateOut)));
  const int smem = 1024 + 16384 + 32768 + 65536;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  struct Case { int N, n2; };
  const Case cases[] = {{16, 0}, {64, 0}, {128, 0}, {160, 0}, {176, 0}, {160, 16}, {224, 0}, {256, 0}};
  for (int grid : {1, 148})
    for (int fill : {0, 16384, 32768})
      for (int mode : {0, 1})
        for (const Case& c : cases) {
          rate_kernel<<<grid, 128, smem>>>(c.N, c.n2, mode, iters, fill, gsrc, out);
          CK(cudaDeviceSynchronize());
          std::vector<RateOut> h(grid);
          CK(cudaMemcpy(h.data(), out, grid * sizeof(RateOut), cudaMemcpyDeviceToHost));
          double cyc = 0, ns = 0;
          for (auto& r : h) { cyc += r.cycles; ns += r.ns; }
          cyc /= grid; ns /= grid;
          const double per = cyc / (iters * 4.0);
          const double flops = 2.0 * 128 * (c.N + c.n2) * 16 * iters * 4.0;
          printf("grid %3d fill %5d %s N=%3d%s: %7.1f cyc/k-step (floor %5.1f)  %6.1f ns total clk %.0f MHz  %.1f TF/s/chip-equivalent\n", grid, fill,
                 mode ? "TS" : "SS", c.N, c.n2 ? "+16" : "   ", per, (c.N + c.n2) / 2.0, ns / 1000.0, cyc / ns * 1000.0,
                 flops / ns * 1e-3 * 148 / 1.0);
        }
  return 0;
}
