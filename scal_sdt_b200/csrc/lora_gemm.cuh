// Types shared by the fused LoRA GEMM kernels (lora_gemm.cu: single CTA; lora_gemm2.cu: CTA pair) and their callers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdt {

// One projection Y = X W^T + bias + (s X la^T) lb^T of a launch.  Every problem of a launch has the same (M, K, N, R,
// scaling); pointers may repeat (q / k / v read the same X).
struct LoraProblem {
  const void* x;       // [M,K] bf16
  const void* w;       // [N,K] bf16 (null when only the rank-R projection is wanted)
  const float* bias;   // [N] f32 or null
  const void* la;      // lora-down [R,K] bf16
  const void* lb;      // lora-up   [N,R] bf16
  void* y;             // [M,N] bf16
  void* t_out;         // [M,R] bf16 or null
  const void* res;     // [M,N] bf16 residual added to the output in the epilogue (y = round(round(proj) + res)), or null
};

// Problems per launch: q/k/v of a self-attention (3), k/v of cross-attentions that read the same text context (2-4).
constexpr int kMaxGroup = 4;

// Kernel-parameter block: the tensor maps of up to G problems (G = 1: 536 bytes; G = 4: 2.1 KB, indexed dynamically in
// the parameter space -- no copy to local memory, `prefetch.tensormap` / TMA take the generic address of the entry).
template <int G>
struct GemmGroup {
  CUtensorMap x[G], w[G], la[G], lb[G];
  CUtensorMap ym[G];           // outputs as [32 rows x 64 columns] boxes, 128-byte swizzle (CTA-pair kernel: TMA stores)
  CUtensorMap ym32[G];         // ... and as [32 x 32] boxes, 64-byte swizzle, for the odd 32-column block of a tile
  uint8_t* y[G];               // outputs again, for the paths that store from registers (residual / GEGLU epilogues, single-CTA kernel)
  const float* bias[G];
  __nv_bfloat16* t_out[G];
  const uint8_t* res[G];       // residual stream added in the epilogue, or null
  // mixed-width launches (PairParams::mixed): output width of every problem and the prefix of its column tiles
  int n[G];
  int tile_begin[G + 1];
};

// Problems of ONE (M, K, rank) but different output widths in one launch: to_k / to_v of every cross-attention of the UNet read
// the same text context (616 x 768) and project it to 320 / 640 / 1280 columns
constexpr int kMaxMixed = 32;

// entry points (lora_gemm.cu / lora_gemm2.cu)
// (the names say bf16 for history: `f16` selects IEEE fp16 operands and outputs on the same kernels)
int lora_gemm_group_bf16(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, int r, bool main,
                         bool f16, cudaStream_t st);
int lora_gemm_pair_group_bf16(const LoraProblem* probs, int n_probs, float scaling, int64_t M, int64_t K, int64_t N, int r,
                              bool f16, cudaStream_t st);

// summed sources (input gradient of q / k / v): see lora_gemm2.cu
bool lora_gemm_pair_sum_supported(int n_src, int64_t M, int64_t K, int64_t N, int r);
int lora_gemm_pair_sum_bf16(const LoraProblem* probs, int n_src, float scaling, int64_t M, int64_t K, int64_t N, int r, bool f16,
                            cudaStream_t st);

// mixed output widths (one X, many projections): see lora_gemm2.cu
bool lora_gemm_pair_mixed_supported(int n_probs, int64_t M, int64_t K, const int64_t* Ns, int r);
int lora_gemm_pair_mixed_bf16(const LoraProblem* probs, const int64_t* Ns, int n_probs, float scaling, int64_t M, int64_t K, int r,
                              bool f16, cudaStream_t st);

// GEGLU epilogue (ff.net.0.proj): see lora_gemm2.cu
bool lora_gemm_pair_geglu_supported(int64_t M, int64_t K, int64_t I, int r);
int lora_gemm_pair_geglu_bf16(const LoraProblem& pr, void* act_out, float scaling, int64_t M, int64_t K, int64_t I, int r, bool f16,
                              cudaStream_t st);

}  // namespace sdt
