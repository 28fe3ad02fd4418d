// f2 neighbour of the LoRA GEMMs: LayerNorm over the channel dimension of token-major bf16 activations [M, C] with FROZEN
// affine parameters, optionally fused with the residual add that precedes it in the transformer block
// (diffusers BasicTransformerBlock: x = x + attn(...); h = norm(x) -- the reference reaches it through the UNet it loads,
// modules/model.py:82-91,304).
//
//   forward :  xs = bf16(x + res)        (only when res != NULL; written to xs_out: the new residual stream)
//              y  = (xs - mean) * rstd * gamma + beta ;   stats[row] = (mean, rstd)
//   backward:  g = dy * gamma ;  dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat))  (+ dres: the gradient that arrives on
//              the residual stream itself)
//
// HBM-bound: one warp per row, the whole row (C <= 2048) lives in registers between the two reductions, 128-bit loads and
// stores, no shared memory.  Algorithmic bytes per row: forward 2 C s (+ 2 C s with the fused add), backward 3 C s (+ C s).
#include "sdt_common.cuh"

namespace sdt {

constexpr int kLnThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_bits_to_f32(u.x & 0xffffu); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = bf16_bits_to_f32(u.y & 0xffffu); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = bf16_bits_to_f32(u.z & 0xffffu); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = bf16_bits_to_f32(u.w & 0xffffu); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ROWS consecutive rows per warp and iteration: all their loads are issued before the first reduction (a 640-byte row alone
// does not keep enough bytes in flight per SM to reach HBM bandwidth).  Rows stay PACKED (bf16 pairs) in registers between
// the passes and are unpacked again where needed -- a shift per element, against half the register footprint and twice the
// resident warps.
template <int NCH, int ROWS, bool kRes>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, const float4* __restrict__ gamma,
              const float4* __restrict__ beta, uint4* __restrict__ xs_out, uint4* __restrict__ y, float2* __restrict__ stats,
              int64_t M, int C8, float eps, float inv_C) {
  pdl_wait();                 // PDL (sdt_common.cuh): inputs come from the kernel in front of us
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kLnThreads / 32);
  for (int64_t row0 = warp * ROWS; row0 < M; row0 += n_warps * ROWS) {
    uint4 xv[ROWS][NCH], rv[ROWS][NCH];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        xv[r][i] = make_uint4(0, 0, 0, 0);
        if (row0 + r < M && c < C8) {
          xv[r][i] = ld_stream(x + (row0 + r) * C8 + c);
          if (kRes) rv[r][i] = ld_stream(res + (row0 + r) * C8 + c);
        }
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t row = row0 + r;
      if (row >= M) break;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (c < C8) {
          float v[8];
          unpack8(xv[r][i], v);
          if (kRes) {
            float rr[8];
            unpack8(rv[r][i], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rr[j];
            xv[r][i] = pack8(v);                          // torch's bf16 add: the stream is rounded before it is normalised
            st_stream(xs_out + row * C8 + c, xv[r][i]);
            unpack8(xv[r][i], v);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) sum += v[j];
        }
      }
      const float mean = warp_sum(sum) * inv_C;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane + 32 * i < C8) {
          float v[8];
          unpack8(xv[r][i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; sq += d * d; }
        }
      const float rstd = rsqrtf(warp_sum(sq) * inv_C + eps);
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (c < C8) {
          const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
          const float4 b0 = __ldg(beta + 2 * c), b1 = __ldg(beta + 2 * c + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float v[8], o[8];
          unpack8(xv[r][i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * gg[j] + bb[j];
          st_stream(y + row * C8 + c, pack8(o));
        }
      }
      if (lane == 0) stats[row] = make_float2(mean, rstd);
    }
  }
}

template <int NCH, int ROWS, bool kRes>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const uint4* __restrict__ xs, const uint4* __restrict__ dy, const uint4* __restrict__ dres,
              const float4* __restrict__ gamma, const float2* __restrict__ stats, uint4* __restrict__ dx, int64_t M, int C8,
              float inv_C) {
  pdl_wait();                 // PDL (sdt_common.cuh): inputs come from the kernel in front of us
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kLnThreads / 32);
  for (int64_t row0 = warp * ROWS; row0 < M; row0 += n_warps * ROWS) {
    uint4 xv[ROWS][NCH], gv[ROWS][NCH], rv[ROWS][NCH];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (row0 + r < M && c < C8) {
          xv[r][i] = ld_stream(xs + (row0 + r) * C8 + c);
          gv[r][i] = ld_stream(dy + (row0 + r) * C8 + c);
          if (kRes) rv[r][i] = ld_stream(dres + (row0 + r) * C8 + c);
        }
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t row = row0 + r;
      if (row >= M) break;
      const float2 ms = stats[row];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (c < C8) {
          float xh[8], g[8];
          unpack8(xv[r][i], xh);
          unpack8(gv[r][i], g);
          const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float gj = g[j] * gg[j];
            s1 += gj;
            s2 += gj * ((xh[j] - ms.x) * ms.y);
          }
        }
      }
      s1 = warp_sum(s1) * inv_C;
      s2 = warp_sum(s2) * inv_C;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane + 32 * i;
        if (c < C8) {
          float xh[8], g[8], o[8];
          unpack8(xv[r][i], xh);
          unpack8(gv[r][i], g);
          const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = ms.y * (g[j] * gg[j] - s1 - ((xh[j] - ms.x) * ms.y) * s2);
          if (kRes) {
            float rr[8];
            unpack8(rv[r][i], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += rr[j];
          }
          st_stream(dx + row * C8 + c, pack8(o));
        }
      }
    }
  }
}

// Short rows (C <= 320: the top UNet level, 3/4 of all tokens): LPR = 8 lanes per row and 32 / LPR rows per warp and
// iteration, up to five vectors per lane.  With a whole warp per 640-byte row 24 of 64 lane slots were idle (40 vectors on 2 x 32
// lanes) and every row paid two 5-step shuffle reductions; here all lanes carry data and a reduction is 3 or 4 steps.  A load
// instruction of the warp covers 32 / LPR full 128-byte segments (rows of 320 k elements are 128-byte aligned).
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, bool kRes>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_sub_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, const float4* __restrict__ gamma,
                  const float4* __restrict__ beta, uint4* __restrict__ xs_out, uint4* __restrict__ y, float2* __restrict__ stats,
                  int64_t M, int C8, float eps, float inv_C) {
  constexpr int NCH = 5, RPW = 32 / LPR;
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, sl = lane % LPR, rsub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kLnThreads / 32);
  for (int64_t row0 = warp * RPW; row0 < M; row0 += n_warps * RPW) {
    const int64_t row = row0 + rsub;
    const bool row_ok = row < M;                       // the shuffles below are executed by every lane of the warp
    uint4 xv[NCH], rv[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      xv[i] = make_uint4(0, 0, 0, 0);
      if (row_ok && c < C8) {
        xv[i] = ld_stream(x + row * C8 + c);
        if (kRes) rv[i] = ld_stream(res + row * C8 + c);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      if (row_ok && c < C8) {
        float v[8];
        unpack8(xv[i], v);
        if (kRes) {
          float rr[8];
          unpack8(rv[i], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rr[j];
          xv[i] = pack8(v);                              // torch's bf16 add: the stream is rounded before it is normalised
          st_stream(xs_out + row * C8 + c, xv[i]);
          unpack8(xv[i], v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[j];
      }
    }
    const float mean = group_sum<LPR>(sum) * inv_C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (row_ok && sl + LPR * i < C8) {
        float v[8];
        unpack8(xv[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; sq += d * d; }
      }
    const float rstd = rsqrtf(group_sum<LPR>(sq) * inv_C + eps);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      if (row_ok && c < C8) {
        const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
        const float4 b0 = __ldg(beta + 2 * c), b1 = __ldg(beta + 2 * c + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float v[8], o[8];
        unpack8(xv[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * gg[j] + bb[j];
        st_stream(y + row * C8 + c, pack8(o));
      }
    }
    if (row_ok && sl == 0) stats[row] = make_float2(mean, rstd);
  }
}

template <int LPR, bool kRes>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_sub_kernel(const uint4* __restrict__ xs, const uint4* __restrict__ dy, const uint4* __restrict__ dres,
                  const float4* __restrict__ gamma, const float2* __restrict__ stats, uint4* __restrict__ dx, int64_t M, int C8,
                  float inv_C) {
  constexpr int NCH = 5, RPW = 32 / LPR;
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, sl = lane % LPR, rsub = lane / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kLnThreads / 32);
  for (int64_t row0 = warp * RPW; row0 < M; row0 += n_warps * RPW) {
    const int64_t row = row0 + rsub;
    const bool row_ok = row < M;
    uint4 xv[NCH], gv[NCH], rv[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      if (row_ok && c < C8) {
        xv[i] = ld_stream(xs + row * C8 + c);
        gv[i] = ld_stream(dy + row * C8 + c);
        if (kRes) rv[i] = ld_stream(dres + row * C8 + c);
      }
    }
    const float2 ms = row_ok ? stats[row] : make_float2(0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      if (row_ok && c < C8) {
        float xh[8], g[8];
        unpack8(xv[i], xh);
        unpack8(gv[i], g);
        const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gj = g[j] * gg[j];
          s1 += gj;
          s2 += gj * ((xh[j] - ms.x) * ms.y);
        }
      }
    }
    s1 = group_sum<LPR>(s1) * inv_C;
    s2 = group_sum<LPR>(s2) * inv_C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + LPR * i;
      if (row_ok && c < C8) {
        float xh[8], g[8], o[8];
        unpack8(xv[i], xh);
        unpack8(gv[i], g);
        const float4 g0 = __ldg(gamma + 2 * c), g1 = __ldg(gamma + 2 * c + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ms.y * (g[j] * gg[j] - s1 - ((xh[j] - ms.x) * ms.y) * s2);
        if (kRes) {
          float rr[8];
          unpack8(rv[i], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rr[j];
        }
        st_stream(dx + row * C8 + c, pack8(o));
      }
    }
  }
}

static int ln_grid(int64_t M, int rows) {
  const int64_t per_block = (int64_t)(kLnThreads / 32) * rows;
  const int64_t blocks = (M + per_block - 1) / per_block;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace sdt

using namespace sdt;

// by row length: rows of up to 40 vectors (C <= 320) go to the sub-warp kernel with 8 lanes per row (measured: 22.9 -> 18.3 us forward,
// 22.5 -> 17.8 us backward at 32768 x 320 with the residual; 16 lanes per row at C = 640 was 7 % SLOWER than a warp per row and
// is not used), longer ones to a warp per row
#define SDT_LN_DISPATCH(KERNEL, RES, ...)                                                                     \
  do {                                                                                                        \
    const int nch = (C8 + 31) / 32;                                                                           \
    if (C8 <= 40) SDT_CUDA_OK(launch_kernel(KERNEL##_sub_kernel<8, RES>, dim3(ln_grid(M, 4)), dim3(kLnThreads), 0, st, true, __VA_ARGS__));       \
    else if (nch <= 2) SDT_CUDA_OK(launch_kernel(KERNEL##_kernel<2, 2, RES>, dim3(ln_grid(M, 2)), dim3(kLnThreads), 0, st, true, __VA_ARGS__));      \
    else if (nch <= 3) SDT_CUDA_OK(launch_kernel(KERNEL##_kernel<3, 1, RES>, dim3(ln_grid(M, 1)), dim3(kLnThreads), 0, st, true, __VA_ARGS__)); \
    else if (nch <= 5) SDT_CUDA_OK(launch_kernel(KERNEL##_kernel<5, 1, RES>, dim3(ln_grid(M, 1)), dim3(kLnThreads), 0, st, true, __VA_ARGS__)); \
    else SDT_CUDA_OK(launch_kernel(KERNEL##_kernel<8, 1, RES>, dim3(ln_grid(M, 1)), dim3(kLnThreads), 0, st, true, __VA_ARGS__));               \
  } while (0)

extern "C" int sdt_layer_norm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* xs_out, void* y,
                                  float* stats, int64_t M, int C, float eps, void* stream) {
  SDT_REQUIRE(x && gamma && beta && y && stats, SDT_ERR_ARG, "sdt_layer_norm_fwd: null pointer");
  SDT_REQUIRE((res == nullptr) == (xs_out == nullptr), SDT_ERR_ARG, "sdt_layer_norm_fwd: res and xs_out go together");
  SDT_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, SDT_ERR_UNSUPPORTED,
              "sdt_layer_norm_fwd: needs C %% 8 == 0 and C <= 2048 (got M=%lld C=%d)", (long long)M, C);
  SDT_REQUIRE(aligned16(x) && aligned16(res) && aligned16(gamma) && aligned16(beta) && aligned16(xs_out) && aligned16(y) &&
                  (reinterpret_cast<uintptr_t>(stats) & 7u) == 0, SDT_ERR_ARG, "sdt_layer_norm_fwd: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int C8 = C / 8;
  const float inv_C = 1.0f / (float)C;
  if (res != nullptr)
    SDT_LN_DISPATCH(ln_fwd, true, (const uint4*)x, (const uint4*)res, (const float4*)gamma, (const float4*)beta,
                    (uint4*)xs_out, (uint4*)y, (float2*)stats, M, C8, eps, inv_C);
  else
    SDT_LN_DISPATCH(ln_fwd, false, (const uint4*)x, nullptr, (const float4*)gamma, (const float4*)beta, nullptr,
                    (uint4*)y, (float2*)stats, M, C8, eps, inv_C);
  SDT_LAUNCH_OK("layer_norm_fwd");
  return SDT_OK;
}

extern "C" int sdt_layer_norm_bwd(const void* xs, const void* dy, const void* dres, const float* gamma, const float* stats,
                                  void* dx, int64_t M, int C, void* stream) {
  SDT_REQUIRE(xs && dy && gamma && stats && dx, SDT_ERR_ARG, "sdt_layer_norm_bwd: null pointer");
  SDT_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, SDT_ERR_UNSUPPORTED,
              "sdt_layer_norm_bwd: needs C %% 8 == 0 and C <= 2048 (got M=%lld C=%d)", (long long)M, C);
  SDT_REQUIRE(aligned16(xs) && aligned16(dy) && aligned16(dres) && aligned16(gamma) && aligned16(dx) &&
                  (reinterpret_cast<uintptr_t>(stats) & 7u) == 0, SDT_ERR_ARG, "sdt_layer_norm_bwd: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int C8 = C / 8;
  const float inv_C = 1.0f / (float)C;
  if (dres != nullptr)
    SDT_LN_DISPATCH(ln_bwd, true, (const uint4*)xs, (const uint4*)dy, (const uint4*)dres, (const float4*)gamma,
                    (const float2*)stats, (uint4*)dx, M, C8, inv_C);
  else
    SDT_LN_DISPATCH(ln_bwd, false, (const uint4*)xs, (const uint4*)dy, nullptr, (const float4*)gamma,
                    (const float2*)stats, (uint4*)dx, M, C8, inv_C);
  SDT_LAUNCH_OK("layer_norm_bwd");
  return SDT_OK;
}
