// C ABI of the LoRA projection (K1 forward, K2 backward): argument checks + dtype dispatch.
//   SDT_BF16 / SDT_F16 -> tcgen05 / TMEM / TMA kernels (lora_gemm.cu, lora_gemm2.cu, lora_wgrad.cu); the 16-bit format is an
//                         instruction-descriptor field and a conversion in the epilogues, the kernels are the same
//   SDT_F32  -> FFMA kernels (simt_gemm.cu): parity path of the reference's fp32 configuration
#include "sdt_common.cuh"
#include "lora_gemm.cuh"

namespace sdt {
int lora_gemm_bf16(const void* x, const void* w, const float* bias, const void* la, const void* lb, float scaling, void* y,
                   void* t_out, int64_t M, int64_t K, int64_t N, int r, bool main, bool f16, cudaStream_t st, const void* res = nullptr);
size_t lora_wgrad_workspace_bytes();
int lora_wgrad_max_sites();
int lora_fwd_f32(const float* x, const float* w, const float* bias, const float* A, const float* B, float scaling,
                 float* y, float* t_save, int64_t M, int64_t K, int64_t N, int r, cudaStream_t st);
int lora_bwd_f32(const float* dy, const float* x, const float* w, const float* A, const float* B, const float* t_save,
                 float scaling, float* dx, float* g_ws, float* dA, float* dB, int64_t M, int64_t K, int64_t N, int r,
                 cudaStream_t st);
struct WgradSite { const void* x; const void* g; float* dA; const void* dy; const void* ts; float* dB; int64_t M, K, N; };
int lora_wgrad_batch_bf16(const WgradSite* sites, int n_sites, int r, int r_true, void* ws, bool f16, cudaStream_t st);
void debug_set(int key, uint64_t value);
}  // namespace sdt

using namespace sdt;

static int lora_linear_fwd_impl(const void* x, const void* w, const float* bias, const void* A, const void* B, float scaling, void* y,
                                void* t_save, const void* residual, int64_t M, int64_t K, int64_t N, int r, int dtype, void* stream);

extern "C" int sdt_lora_linear_fwd(const void* x, const void* w, const float* bias, const void* A, const void* B,
                                   float scaling, void* y, void* t_save, int64_t M, int64_t K, int64_t N, int r,
                                   int dtype, void* stream) {
  return lora_linear_fwd_impl(x, w, bias, A, B, scaling, y, t_save, nullptr, M, K, N, r, dtype, stream);
}

extern "C" int sdt_lora_linear_fwd_res(const void* x, const void* w, const float* bias, const void* A, const void* B,
                                       float scaling, const void* residual, void* y, void* t_save, int64_t M, int64_t K, int64_t N,
                                       int r, int dtype, void* stream) {
  SDT_REQUIRE(residual != nullptr, SDT_ERR_ARG, "sdt_lora_linear_fwd_res: null residual");
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED,
              "sdt_lora_linear_fwd_res: the residual epilogue is bf16 / fp16 (fp32: add after sdt_lora_linear_fwd)");
  return lora_linear_fwd_impl(x, w, bias, A, B, scaling, y, t_save, residual, M, K, N, r, dtype, stream);
}

static int lora_linear_fwd_impl(const void* x, const void* w, const float* bias, const void* A, const void* B, float scaling, void* y,
                                void* t_save, const void* residual, int64_t M, int64_t K, int64_t N, int r, int dtype, void* stream) {
  SDT_REQUIRE(x && w && y, SDT_ERR_ARG, "sdt_lora_linear_fwd: null pointer");
  SDT_REQUIRE(M > 0 && K > 0 && N > 0 && r >= 0, SDT_ERR_ARG, "sdt_lora_linear_fwd: bad sizes M=%lld K=%lld N=%lld r=%d",
              (long long)M, (long long)K, (long long)N, r);
  SDT_REQUIRE((r == 0) == (A == nullptr) && (r == 0) == (B == nullptr), SDT_ERR_ARG,
              "sdt_lora_linear_fwd: A and B must be given exactly when r > 0");
  SDT_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(A) && aligned16(B) && aligned16(t_save),
              SDT_ERR_ARG, "sdt_lora_linear_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SDT_BF16 || dtype == SDT_F16) {
    SDT_REQUIRE(r == 0 || t_save != nullptr, SDT_ERR_ARG, "sdt_lora_linear_fwd: t_save is required when r > 0");
    return lora_gemm_bf16(x, w, bias, A, B, scaling, y, t_save, M, K, N, r, true, dtype == SDT_F16, st, residual);
  }
  if (dtype == SDT_F32)
    return lora_fwd_f32((const float*)x, (const float*)w, bias, (const float*)A, (const float*)B, scaling, (float*)y,
                        (float*)t_save, M, K, N, r, st);
  set_error("sdt_lora_linear_fwd: unsupported dtype %d (there is no fallback path)", dtype);
  return SDT_ERR_UNSUPPORTED;
}

extern "C" int sdt_lora_linear_fwd_multi_supported(int n_problems, int64_t M, int64_t K, const int64_t* Ns, int r) {
  return (Ns != nullptr && lora_gemm_pair_mixed_supported(n_problems, M, K, Ns, r)) ? 1 : 0;
}

extern "C" int sdt_lora_linear_fwd_multi(const sdt_lora_problem* problems, const int64_t* Ns, int n_problems, float scaling, int64_t M,
                                         int64_t K, int r, int dtype, void* stream) {
  SDT_REQUIRE(problems != nullptr && Ns != nullptr, SDT_ERR_ARG, "sdt_lora_linear_fwd_multi: null pointer");
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_linear_fwd_multi: bf16 / fp16 only (there is no fallback)");
  return lora_gemm_pair_mixed_bf16(reinterpret_cast<const LoraProblem*>(problems), Ns, n_problems, scaling, M, K, r, dtype == SDT_F16,
                                   (cudaStream_t)stream);
}

extern "C" int sdt_lora_linear_geglu_supported(int64_t M, int64_t K, int64_t I, int r) {
  return lora_gemm_pair_geglu_supported(M, K, I, r) ? 1 : 0;
}

extern "C" int sdt_lora_linear_geglu_fwd(const void* x, const void* w, const float* bias, const void* A, const void* B, float scaling,
                                         void* proj, void* act, void* t_save, int64_t M, int64_t K, int64_t I, int r, int dtype,
                                         void* stream) {
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_linear_geglu_fwd: bf16 / fp16 only (there is no fallback)");
  const LoraProblem pr{x, w, bias, A, B, proj, t_save};
  return lora_gemm_pair_geglu_bf16(pr, act, scaling, M, K, I, r, dtype == SDT_F16, (cudaStream_t)stream);
}

static_assert(sizeof(sdt_lora_problem) == sizeof(LoraProblem), "sdt_lora_problem mirrors sdt::LoraProblem");

extern "C" int sdt_lora_linear_fwd_group(const sdt_lora_problem* problems, int n_problems, float scaling, int64_t M, int64_t K,
                                         int64_t N, int r, int dtype, void* stream) {
  SDT_REQUIRE(problems != nullptr && n_problems >= 1 && n_problems <= SDT_MAX_GROUP, SDT_ERR_ARG,
              "sdt_lora_linear_fwd_group: 1..%d problems per launch (got %d)", SDT_MAX_GROUP, n_problems);
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED,
              "sdt_lora_linear_fwd_group: bf16 / fp16 only (fp32 sites go through sdt_lora_linear_fwd one by one; there is no fallback)");
  SDT_REQUIRE(r == 16 || r == 32 || r == 64, SDT_ERR_UNSUPPORTED, "sdt_lora_linear_fwd_group: padded rank must be 16, 32 or 64 (got %d)", r);
  for (int q = 0; q < n_problems; ++q)
    SDT_REQUIRE(problems[q].x && problems[q].w && problems[q].A && problems[q].B && problems[q].y && problems[q].t_save, SDT_ERR_ARG,
                "sdt_lora_linear_fwd_group: null pointer in problem %d", q);
  return lora_gemm_group_bf16(reinterpret_cast<const LoraProblem*>(problems), n_problems, scaling, M, K, N, r, true,
                              dtype == SDT_F16, (cudaStream_t)stream);
}

extern "C" int sdt_lora_linear_bwd(const void* dy, const void* x, const void* wt, const void* At, const void* Bt,
                                   const void* t_save, float scaling, void* dx, void* g_ws, float* dA, float* dB,
                                   int64_t M, int64_t K, int64_t N, int r, int r_true, int dtype, void* ws, void* stream) {
  SDT_REQUIRE(dy, SDT_ERR_ARG, "sdt_lora_linear_bwd: null dy");
  SDT_REQUIRE(M > 0 && K > 0 && N > 0 && r >= 0, SDT_ERR_ARG, "sdt_lora_linear_bwd: bad sizes");
  SDT_REQUIRE(dx != nullptr || r > 0, SDT_ERR_ARG, "sdt_lora_linear_bwd: nothing to compute");
  SDT_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(wt) && aligned16(At) && aligned16(Bt) && aligned16(t_save) &&
                  aligned16(dx) && aligned16(g_ws), SDT_ERR_ARG, "sdt_lora_linear_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SDT_BF16 || dtype == SDT_F16) {
    const bool f16 = dtype == SDT_F16;
    // dA == NULL and dB == NULL: only dX and G are computed; the caller batches the two reductions of several sites into one
    // sdt_lora_wgrad_batch launch (a transformer block's worth) instead of one small launch per site
    const bool defer = (dA == nullptr && dB == nullptr);
    if (r > 0)
      SDT_REQUIRE(Bt && g_ws && (dx == nullptr || At) && (defer || (x && t_save && dA && dB)), SDT_ERR_ARG,
                  "sdt_lora_linear_bwd: Bt, g_ws (and At for dX; x, t_save, dA, dB unless deferred) are required when r > 0");
    SDT_REQUIRE(dx == nullptr || wt != nullptr, SDT_ERR_ARG, "sdt_lora_linear_bwd: wt is required for dX");
    // G = s dY B ; dX = dY W + G A   -- the forward kernel with (dY, W^T, B^T, A^T)
    int rc = lora_gemm_bf16(dy, wt, nullptr, Bt, At, scaling, dx, g_ws, M, /*contraction*/ N, /*outputs*/ K, r,
                            dx != nullptr, f16, st);
    if (rc != SDT_OK || r == 0 || defer) return rc;
    // dA[j,k] += sum_m G[m,j] X[m,k]  and  dB[n,j] += sum_m dY[m,n] Ts[m,j]: one launch for both reductions
    const WgradSite site{x, g_ws, dA, dy, t_save, dB, M, K, N};
    return lora_wgrad_batch_bf16(&site, 1, r, r_true, ws, f16, st);
  }
  if (dtype == SDT_F32) {
    SDT_REQUIRE(r_true == r, SDT_ERR_ARG, "sdt_lora_linear_bwd(f32): r_true must equal r");
    return lora_bwd_f32((const float*)dy, (const float*)x, (const float*)wt, (const float*)At, (const float*)Bt,
                        (const float*)t_save, scaling, (float*)dx, (float*)g_ws, dA, dB, M, K, N, r, st);
  }
  set_error("sdt_lora_linear_bwd: unsupported dtype %d (there is no fallback path)", dtype);
  return SDT_ERR_UNSUPPORTED;
}

extern "C" int sdt_lora_linear_bwd_group_supported(int n_problems, int need_dx, int64_t M, int64_t K, int64_t N, int r) {
  if (!need_dx)      // G_q = s dY_q B_q of up to SDT_MAX_GROUP projections as the work items of one launch
    return (n_problems >= 2 && n_problems <= SDT_MAX_GROUP && (r == 16 || r == 32 || r == 64)) ? 1 : 0;
  // the summed GEMM contracts over N (features of dY) and produces K columns of dX
  return lora_gemm_pair_sum_supported(n_problems, M, N, K, r) ? 1 : 0;
}

extern "C" int sdt_lora_linear_bwd_group(const sdt_lora_bwd_problem* problems, int n_problems, float scaling, void* dx, int64_t M,
                                         int64_t K, int64_t N, int r, int r_true, int dtype, void* ws, void* stream) {
  SDT_REQUIRE(problems != nullptr, SDT_ERR_ARG, "sdt_lora_linear_bwd_group: null pointer");
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_linear_bwd_group: bf16 / fp16 only (there is no fallback)");
  const bool f16 = dtype == SDT_F16;
  SDT_REQUIRE(M > 0 && K > 0 && N > 0, SDT_ERR_ARG, "sdt_lora_linear_bwd_group: bad sizes");
  SDT_REQUIRE(sdt_lora_linear_bwd_group_supported(n_problems, dx != nullptr, M, K, N, r), SDT_ERR_UNSUPPORTED,
              "sdt_lora_linear_bwd_group: unsupported group (%d projections, r=%d, M=%lld, dx %s): use sdt_lora_linear_bwd per site",
              n_problems, r, (long long)M, dx ? "wanted" : "not wanted");
  cudaStream_t st = (cudaStream_t)stream;
  LoraProblem pr[SDT_MAX_GROUP];
  for (int q = 0; q < n_problems; ++q) {
    const sdt_lora_bwd_problem& b = problems[q];
    SDT_REQUIRE(b.dy && b.x && b.At && b.Bt && b.t_save && b.g_ws && (dx == nullptr || b.wt), SDT_ERR_ARG,
                "sdt_lora_linear_bwd_group: null pointer in problem %d", q);
    SDT_REQUIRE((b.dA == nullptr) == (b.dB == nullptr) && (b.dA == nullptr) == (problems[0].dA == nullptr), SDT_ERR_ARG,
                "sdt_lora_linear_bwd_group: dA / dB must be given for every problem or for none (deferred reductions)");
    // G_q = s dY_q B_q ; dX += dY_q W_q + G_q A_q   -- the forward kernel's roles with (dY, W^T, B^T, A^T)
    pr[q] = LoraProblem{b.dy, dx ? b.wt : nullptr, nullptr, b.Bt, dx ? b.At : nullptr, dx, b.g_ws};
  }
  int rc = dx != nullptr ? lora_gemm_pair_sum_bf16(pr, n_problems, scaling, M, /*contraction*/ N, /*outputs*/ K, r, f16, st)
                         : lora_gemm_group_bf16(pr, n_problems, scaling, M, N, K, r, /*main=*/false, f16, st);
  if (rc != SDT_OK || problems[0].dA == nullptr) return rc;
  // the dA / dB reductions of the whole group: one launch
  WgradSite sites[SDT_MAX_GROUP];
  for (int q = 0; q < n_problems; ++q) {
    const sdt_lora_bwd_problem& b = problems[q];
    sites[q] = WgradSite{b.x, b.g_ws, b.dA, b.dy, b.t_save, b.dB, M, K, N};
  }
  return lora_wgrad_batch_bf16(sites, n_problems, r, r_true, ws, f16, st);
}

extern "C" size_t sdt_lora_wgrad_workspace_bytes(void) { return lora_wgrad_workspace_bytes(); }

static_assert(sizeof(sdt_wgrad_site) == sizeof(WgradSite), "sdt_wgrad_site mirrors sdt::WgradSite");

extern "C" int sdt_lora_wgrad_max_sites(void) { return lora_wgrad_max_sites(); }

extern "C" int sdt_lora_wgrad_batch(const sdt_wgrad_site* sites, int n_sites, int r, int r_true, int dtype, void* ws, void* stream) {
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_wgrad_batch: bf16 / fp16 only (there is no fallback)");
  return lora_wgrad_batch_bf16(reinterpret_cast<const WgradSite*>(sites), n_sites, r, r_true, ws, dtype == SDT_F16,
                               (cudaStream_t)stream);
}

extern "C" int sdt_debug_set(int key, uint64_t value) {
  debug_set(key, value);
  return SDT_OK;
}
