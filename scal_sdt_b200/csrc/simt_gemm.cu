// FFMA (CUDA-core) strided GEMM.  Two uses:
//   1. the SDT_F32 variant of the LoRA projection (sdt_lora_linear_fwd/bwd with dtype f32) -- the
//      parity path for the reference's fp32 CPU configuration (BASELINE cfg1), where 1e-5 relative
//      is out of reach for single-pass tensor-core math;
//   2. sdt_simt_gemm_f32, the slow on-device cross-check the GPU tests run against the tcgen05 kernels.
//
//   C[m,n] = alpha * sum_k A[m*lda_m + k*lda_k] * B[n*ldb_n + k*ldb_k] + beta * C[m,n] + bias[n]
#include "sdt_common.cuh"

namespace sdt {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
simt_gemm_kernel(const float* __restrict__ A, int64_t lda_m, int64_t lda_k, const float* __restrict__ B, int64_t ldb_n,
                 int64_t ldb_k, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, float alpha, float beta,
                 int64_t M, int64_t N, int64_t K) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  float acc[4][4] = {};

  for (int64_t k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + u * 256;                 // 1024 elements per operand tile
      int i, kk;
      if (lda_k == 1) { kk = e & 15; i = e >> 4; } else { i = e & 63; kk = e >> 6; }   // unit stride along threads
      const int64_t m = m0 + i, k = k0 + kk;
      As[kk][i] = (m < M && k < K) ? __ldg(A + m * lda_m + k * lda_k) : 0.f;
      if (ldb_k == 1) { kk = e & 15; i = e >> 4; } else { i = e & 63; kk = e >> 6; }
      const int64_t n = n0 + i;
      const int64_t kb = k0 + kk;
      Bs[kk][i] = (n < N && kb < K) ? __ldg(B + n * ldb_n + kb * ldb_k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias != nullptr) v += __ldg(bias + n);
      if (beta != 0.f) v += beta * C[m * ldc + n];
      C[m * ldc + n] = v;
    }
  }
}

int simt_gemm(const float* A, int64_t lda_m, int64_t lda_k, const float* B, int64_t ldb_n, int64_t ldb_k, float* C,
              int64_t ldc, const float* bias, float alpha, float beta, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  if (M <= 0 || N <= 0) return SDT_OK;
  dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM));
  SDT_REQUIRE(grid.y <= 65535, SDT_ERR_UNSUPPORTED, "simt_gemm: M=%lld too large for the f32 path", (long long)M);
  simt_gemm_kernel<<<grid, 256, 0, st>>>(A, lda_m, lda_k, B, ldb_n, ldb_k, C, ldc, bias, alpha, beta, M, N, K);
  SDT_LAUNCH_OK("simt_gemm");
  return SDT_OK;
}

// ---- SDT_F32 LoRA projection ------------------------------------------------------------------
int lora_fwd_f32(const float* x, const float* w, const float* bias, const float* A, const float* B, float scaling,
                 float* y, float* t_save, int64_t M, int64_t K, int64_t N, int r, cudaStream_t st) {
  int rc = simt_gemm(x, K, 1, w, K, 1, y, N, bias, 1.f, 0.f, M, N, K, st);                 // Y = X W^T + b
  if (rc != SDT_OK || r == 0) return rc;
  SDT_REQUIRE(t_save != nullptr, SDT_ERR_ARG, "sdt_lora_linear_fwd(f32): t_save is required when r > 0");
  rc = simt_gemm(x, K, 1, A, K, 1, t_save, r, nullptr, scaling, 0.f, M, r, K, st);         // Ts = s X A^T
  if (rc != SDT_OK) return rc;
  return simt_gemm(t_save, r, 1, B, r, 1, y, N, nullptr, 1.f, 1.f, M, N, r, st);           // Y += Ts B^T
}

int lora_bwd_f32(const float* dy, const float* x, const float* w, const float* A, const float* B, const float* t_save,
                 float scaling, float* dx, float* g_ws, float* dA, float* dB, int64_t M, int64_t K, int64_t N, int r,
                 cudaStream_t st) {
  int rc = SDT_OK;
  if (r > 0) {
    SDT_REQUIRE(g_ws && t_save && dA && dB, SDT_ERR_ARG, "sdt_lora_linear_bwd(f32): g_ws, t_save, dA, dB are required");
    rc = simt_gemm(dy, N, 1, B, 1, r, g_ws, r, nullptr, scaling, 0.f, M, r, N, st);        // G = s dY B
    if (rc != SDT_OK) return rc;
  }
  if (dx != nullptr) {
    rc = simt_gemm(dy, N, 1, w, 1, K, dx, K, nullptr, 1.f, 0.f, M, K, N, st);              // dX = dY W
    if (rc != SDT_OK) return rc;
    if (r > 0) {
      rc = simt_gemm(g_ws, r, 1, A, 1, K, dx, K, nullptr, 1.f, 1.f, M, K, r, st);          // dX += G A
      if (rc != SDT_OK) return rc;
    }
  }
  if (r > 0) {
    rc = simt_gemm(g_ws, 1, r, x, 1, K, dA, K, nullptr, 1.f, 1.f, r, K, M, st);            // dA += G^T X
    if (rc != SDT_OK) return rc;
    rc = simt_gemm(dy, 1, N, t_save, 1, r, dB, r, nullptr, 1.f, 1.f, N, r, M, st);         // dB += dY^T Ts
  }
  return rc;
}

}  // namespace sdt

extern "C" int sdt_simt_gemm_f32(const float* A, int64_t lda_m, int64_t lda_k, const float* B, int64_t ldb_n, int64_t ldb_k,
                                 float* C, int64_t ldc, const float* bias, float alpha, float beta,
                                 int64_t M, int64_t N, int64_t K, void* stream) {
  SDT_REQUIRE(A && B && C, SDT_ERR_ARG, "sdt_simt_gemm_f32: null pointer");
  return sdt::simt_gemm(A, lda_m, lda_k, B, ldb_n, ldb_k, C, ldc, bias, alpha, beta, M, N, K, (cudaStream_t)stream);
}
