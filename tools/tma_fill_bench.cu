// Microbenchmark: how many bytes per clock an SM can pull out of L2 into shared memory with TMA tile loads, and whether several
// SMs that need the SAME tile are cheaper when it is multicast through a cluster than when every SM fetches it itself.
//
// Every CTA (one per SM, 148 of them) keeps a ring of 8 x 16 KiB stages full of [128 rows x 64 bf16] tiles (128-byte swizzle, the X
// k-block of the LoRA GEMM) out of an L2-resident matrix; nothing consumes the data, a stage is refilled as soon as it has landed.
//   unicast / distinct   every CTA reads its own tiles
//   unicast / shared CL  the CL CTAs of a cluster read the same tiles at the same time, each with its own loads (L2 de-duplication?)
//   multicast CL         the CL CTAs of a cluster each load 1/CL of the tile's rows and multicast them to all CL CTAs
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scal_sdt_b200/csrc tools/tma_fill_bench.cu -o tools/_build/tma_fill_bench -lcuda
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <vector>

using namespace sdt::ptx;

constexpr int kStages = 8, kTileRows = 128, kTileBytes = kTileRows * 128, kTilesPerGroup = 16;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, int c_inner, int c_row, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_row), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t rank) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(bar)), "r"(rank));
  // relaxed: no data is handed over through this barrier, and a release at cluster scope is a MEMBAR.ALL.GPU per arrive
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}

struct Out { long long cycles; };

// CL = cluster size (1, 2, 4); MC = multicast; shared = the CTAs of a cluster read the same tiles
template <int CL, bool MC>
__global__ void __launch_bounds__(64, 1) fill_kernel(const __grid_constant__ CUtensorMap full_map, const __grid_constant__ CUtensorMap part_map,
                                                    int shared, int iters, Out* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kStages], empty[kStages];
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  const int cluster_id = blockIdx.x / CL;
  const int group = (shared || MC) ? cluster_id : (int)blockIdx.x;       // which set of tiles this CTA reads
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    fence_mbar_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (!MC) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int it = 0; it < iters + kStages; ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&full[s], ((it / kStages) - 1) & 1);               // the tile of the previous round has landed here
        if (it < iters) {
          const int row0 = (group * kTilesPerGroup + it % kTilesPerGroup) * kTileRows;
          mbar_arrive_expect_tx(&full[s], kTileBytes);
          tma_load_2d(smem + s * kTileBytes, &full_map, 0, row0, &full[s]);
        }
      }
      out[blockIdx.x].cycles = clock64() - t0;
    }
  } else if (threadIdx.x == 0) {
    // producer: my 1 / CL of the tile goes to every CTA of the cluster once all of them have released the stage
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      if (it >= kStages) mbar_wait(&empty[s], ((it / kStages) - 1) & 1);
      const int row0 = (group * kTilesPerGroup + it % kTilesPerGroup) * kTileRows;
      mbar_arrive_expect_tx(&full[s], kTileBytes);
      tma_load_2d_mc(smem + s * kTileBytes + rank * (kTileBytes / CL), &part_map, 0, row0 + (int)rank * (kTileRows / CL), &full[s],
                     (uint16_t)((1u << CL) - 1));
    }
    for (int it = iters - kStages; it < iters; ++it) mbar_wait(&empty[it % kStages], (it / kStages) & 1);     // everything has landed everywhere
    out[blockIdx.x].cycles = clock64() - t0;
  } else if (threadIdx.x == 32) {
    // consumer: a landed stage is handed straight back to every producer of the cluster
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      mbar_wait(&full[s], (it / kStages) & 1);
      for (uint32_t r = 0; r < CL; ++r) remote_arrive(&empty[s], r);
    }
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static CUtensorMap make_map(void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
  return m;
}

template <int CL, bool MC>
static void run(const char* name, void* base, uint64_t rows, int shared, Out* out) {
  const int grid = 148 / CL * CL, iters = 4000, smem = 1024 + kStages * kTileBytes;
  CK(cudaFuncSetAttribute(fill_kernel<CL, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const CUtensorMap full_map = make_map(base, rows, kTileRows), part_map = make_map(base, rows, kTileRows / CL);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {                 // the first pass brings the matrix into L2
    CK(cudaLaunchKernelEx(&cfg, fill_kernel<CL, MC>, full_map, part_map, shared, iters, out));
    CK(cudaDeviceSynchronize());
  }
  std::vector<Out> h(grid);
  CK(cudaMemcpy(h.data(), out, grid * sizeof(Out), cudaMemcpyDeviceToHost));
  double cyc = 0;
  for (auto& o : h) cyc += o.cycles;
  cyc /= grid;
  const double per_sm = (double)iters * kTileBytes / cyc;
  const double l2_factor = (shared || MC) ? 1.0 / CL : 1.0;       // distinct bytes leaving L2 per byte landing in an SM (if fully de-duplicated)
  printf("%-34s: %6.1f B/clk landing per SM, %7.0f B/clk chip-wide; distinct bytes requested from L2: %6.1f B/clk per SM\n", name, per_sm,
         per_sm * grid, per_sm * l2_factor);
}

int main() {
  CK(cudaFree(0));
  const uint64_t rows = (uint64_t)148 * kTilesPerGroup * kTileRows;     // 148 groups x 16 tiles x 16 KiB = 37 MiB: L2-resident
  void* base;
  CK(cudaMalloc(&base, rows * 128));
  CK(cudaMemset(base, 1, rows * 128));
  Out* out;
  CK(cudaMalloc(&out, 148 * sizeof(Out)));
  run<1, false>("unicast, distinct tiles", base, rows, 0, out);
  run<2, false>("unicast, cluster of 2 shares tiles", base, rows, 1, out);
  run<4, false>("unicast, cluster of 4 shares tiles", base, rows, 1, out);
  run<2, true>("multicast, cluster of 2", base, rows, 1, out);
  run<4, true>("multicast, cluster of 4", base, rows, 1, out);
  run<1, false>("unicast, distinct tiles (again)", base, rows, 0, out);
  return 0;
}
