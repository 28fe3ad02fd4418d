// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences) and the UMMA shared-memory + instruction descriptors.  No CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>      // CUtensorMap (types only; the encode function is fetched at run time)
#include <stdint.h>

namespace sdt {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// non-blocking probe: true when the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---- proxy / tcgen05 fences ---------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates are (inner element index, row index); completes `bytes` on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c_inner, int c_row, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_row), "r"(smem_u32(bar))
      : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
// whole warp; writes the base address (lane<<16 | column) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- UMMA ---------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// instruction descriptor, kind::f16, bf16 x bf16 -> f32
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt  [15] A major (0 = K)  [16] B major
//   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same descriptor with IEEE fp16 operands (A fmt = B fmt = 0): clear the two format fields
constexpr uint32_t kIdescFmtBf16 = (1u << 7) | (1u << 10);
__host__ __device__ constexpr uint32_t idesc_operand_format(uint32_t idesc_bf16, bool f16) {
  return f16 ? (idesc_bf16 & ~kIdescFmtBf16) : idesc_bf16;
}

// shared-memory matrix descriptor
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version (1 on sm_100)   [61,64) layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B
enum : uint32_t { kLayoutNone = 0, kLayoutSW128 = 2, kLayoutSW64 = 4, kLayoutSW32 = 6 };
__host__ __device__ constexpr uint64_t make_smem_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t base, uint32_t smem_addr) {
  return base | (uint64_t)((smem_addr >> 4) & 0x3FFF);
}

// ---- TMEM -> registers: this warp's 32 lanes (thread i = lane i of the warp's quarter), N consecutive columns
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- epilogue stores ------------------------------------------------------------------------------------------------
// A thread of an epilogue warp holds 32 consecutive output columns of ONE row (TMEM lane = row); written as they lie that
// is 64 bytes per row and request, and the memory system then runs at about half its write bandwidth (measured: the
// K = 320 projections were bound by exactly this, with TMA stores of 32-column boxes as much as with register stores).
// So each warp transposes a [32 rows x 64 columns] block through 4 KiB of its own shared memory (XOR-swizzled 16-byte
// slots, conflict-free both ways, __syncwarp only) and writes it out with 8 lanes per row: every store instruction covers
// four full 128-byte lines.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint32_t* r) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// stage this lane's 32 packed columns (sub-chunk h = 0 / 1 of the 64-column block) of row `lane`
__device__ __forceinline__ void stage_row_chunk(uint32_t stg, int lane, int h, const uint32_t (&pk)[16]) {
  const uint32_t row_base = stg + lane * 128;
#pragma unroll
  for (int j = 0; j < 4; ++j) st_shared_v4(row_base + (((4 * h + j) ^ (lane & 7)) << 4), pk + 4 * j);
}
// The same 32 values as one row of a [32 rows x 64 bytes] block in the 64-byte TMA swizzle (16-byte chunk j of row r at chunk
// j ^ ((r >> 1) & 3)): what a tensor map with a 32-column box and SWIZZLE_64B reads.  Eight consecutive lanes cover eight distinct
// 16-byte bank groups: conflict-free.
__device__ __forceinline__ void stage_row_sw64(uint32_t stg, int lane, const uint32_t (&pk)[16]) {
  const uint32_t row_base = stg + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) st_shared_v4(row_base + ((j ^ ((lane >> 1) & 3)) << 4), pk + 4 * j);
}
// TMA store of one box from (own) shared memory; completion is tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c_inner, int c_row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c_inner), "r"(c_row), "r"(src_smem)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every bulk group committed by this thread has finished READING shared memory (the buffers may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// a + b on four packed pairs of 16-bit values (bf16 or fp16), each sum rounded once: what torch's elementwise add of two
// bf16 / fp16 tensors produces
__device__ __forceinline__ uint32_t add_packed_pair(uint32_t a, uint32_t b, bool f16) {
  if (f16) {
    const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&a)), y = __half22float2(*reinterpret_cast<const __half2*>(&b));
    const __half2 r = __floats2half2_rn(x.x + y.x, x.y + y.y);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const float lo = __uint_as_float(a << 16) + __uint_as_float(b << 16);
  const float hi = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
  const __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&r);
}

// write the staged block out: rows [row0, row0 + 32) x 16-byte slots [0, n_slots) starting at column col0 of a row-major
// bf16 matrix with N columns (N % 8 == 0); rows >= M and columns >= N are clipped (sw64: n_slots <= 4).  res (optional, same shape as y): the
// residual stream the projection's output is added to (x + ff(h), y + res of a transformer block) -- read here, added to the
// rounded output and rounded again, exactly torch's separate add, without that add's pass over both tensors.
__device__ __forceinline__ void write_staged_block(uint32_t stg, int lane, uint8_t* y, int row0, int M, int col0, int N, int n_slots,
                                                   const uint8_t* res = nullptr, bool f16 = false, bool sw64 = false) {
  const int c = lane & 7;
  // staged as [32 x 128 B] (16-byte slot c of row r at c ^ (r & 7)) or, the odd 32-column block, as [32 x 64 B] (slot c at c ^ ((r >> 1) & 3))
  auto slot_addr = [&](int rl) -> uint32_t { return sw64 ? stg + rl * 64 + ((c ^ ((rl >> 1) & 3)) << 4) : stg + rl * 128 + ((c ^ (rl & 7)) << 4); };
  const bool col_ok = c < n_slots && col0 + 8 * c < N;
  if (res == nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + (lane >> 3);
      if (col_ok && row0 + rl < M) {
        const uint4 v = ld_shared_v4(slot_addr(rl));
        asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(y + ((size_t)(row0 + rl) * N + col0 + 8 * c) * 2), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      }
    }
    return;
  }
  uint4 r[8];                                        // all residual loads in flight before the first store
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + (lane >> 3);
    if (col_ok && row0 + rl < M)
      asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(r[i].x), "=r"(r[i].y), "=r"(r[i].z), "=r"(r[i].w)
                   : "l"(res + ((size_t)(row0 + rl) * N + col0 + 8 * c) * 2));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + (lane >> 3);
    if (col_ok && row0 + rl < M) {
      const uint4 v = ld_shared_v4(slot_addr(rl));
      asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(y + ((size_t)(row0 + rl) * N + col0 + 8 * c) * 2),
                   "r"(add_packed_pair(v.x, r[i].x, f16)), "r"(add_packed_pair(v.y, r[i].y, f16)), "r"(add_packed_pair(v.z, r[i].z, f16)),
                   "r"(add_packed_pair(v.w, r[i].w, f16))
                   : "memory");
    }
  }
}

}  // namespace ptx

// ---- host side: tensor-map encode through the runtime's driver entry point (no -lcuda) -----------
enum TmapSwizzle { TMAP_SW_NONE = 0, TMAP_SW_32 = 1, TMAP_SW_64 = 2, TMAP_SW_128 = 3 };
// 2-D row-major bf16 tensor [rows, cols] with `pitch_bytes` between rows; box = box_rows x box_cols.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols, TmapSwizzle swz);

}  // namespace sdt
