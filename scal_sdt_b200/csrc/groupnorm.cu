// f2: GroupNorm (+ SiLU) on channels-last activations, forward and backward (frozen affine parameters).
//
// torch's CUDA group_norm only takes NCHW-contiguous input, so in a channels-last UNet every GroupNorm costs two
// layout copies in the forward and two more in the backward, next to the separate SiLU kernels.  These kernels work on
// the [B, HW, C] memory directly: one statistics pass (per-(b,group) sum / sum of squares, fp32, reduced in registers
// per channel column, then shared-memory and global red.add) and one apply pass with the activation fused.
// HBM-bound: forward 2 reads + 1 write of the tensor, backward 4 reads + 1 write (the second read of each pair is
// normally an L2 hit: the tensors are tens of MB).
//
// Both passes first fold everything that depends only on (sample, channel) into a small shared-memory table
//     z    = x * A + Bc        A = rstd gamma,  Bc = beta - mean rstd gamma        (pre-activation)
//     xhat = x * Sc + Sh       Sc = rstd,       Sh = -mean rstd
//     dx   = A dz - D - E x    D = rstd (P1 + P2 Sh),  E = rstd^2 P2               (P1, P2: backward group means)
// so the streaming loops carry no per-thread set-up (a UNet step runs 119 of these launches, many on 1-5 MB tensors where
// a chain of dependent global loads per thread costs more than the stream itself) and the grid can fill every SM.
#include "sdt_common.cuh"

namespace sdt {

struct GnShape {
  int64_t HW;
  int C, G, cpg, vecs;          // vecs = C / 8 (16-byte vectors per pixel row)
  int threads;
  int rows_par;                 // pixel rows processed concurrently by one CTA in the statistics pass
  int rows_per_cta;
};

// sigmoid(z) = 0.5 tanh(z/2) + 0.5 with the hardware tanh: ONE special-function op per element instead of ex2 + a full-precision
// division (these kernels were bound by the special-function / issue rate, not by memory); |error| < 2^-11, below bf16 resolution
__device__ __forceinline__ float sigmoidf_(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float bf16_elem(const uint4& u, int j) {
  const uint32_t w = (&u.x)[j >> 1];
  return (j & 1) ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
}

// per-channel table rows (each a float[C] in dynamic shared memory)
enum { T_A = 0, T_BC = 1, T_SC = 2, T_SH = 3, T_D = 2, T_E = 3 };

// Per-(sample, channel) table in shared memory.  The loads of TU channels per thread are all issued before the first use:
// one global-memory latency per CTA instead of C / blockDim of them (C = 2560 with 160 threads was 16 dependent rounds).
//   kStatsBwd: rows A, Bc, Sc, Sh (backward statistics pass)        kApplyBwd: rows A, Bc, D, E        neither: A, Bc
template <bool kStatsBwd, bool kApplyBwd>
__device__ __forceinline__ void build_table(float* tab, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ fstats, const float* __restrict__ bstats, int b,
                                            const GnShape& s, float eps) {
  constexpr int TU = 8;
  const float inv_n = 1.0f / ((float)s.HW * (float)s.cpg);
  for (int cb = threadIdx.x; cb < s.C; cb += TU * blockDim.x) {
    float gm[TU], bt[TU], f0[TU], f1[TU], b0[TU], b1[TU];
#pragma unroll
    for (int t = 0; t < TU; ++t) {
      const int c = cb + t * blockDim.x;
      if (c < s.C) {
        const size_t gi = ((size_t)b * s.G + c / s.cpg) * 2;
        gm[t] = __ldg(gamma + c);
        bt[t] = __ldg(beta + c);
        f0[t] = fstats[gi];
        f1[t] = fstats[gi + 1];
        if (kApplyBwd) { b0[t] = bstats[gi]; b1[t] = bstats[gi + 1]; }
      }
    }
#pragma unroll
    for (int t = 0; t < TU; ++t) {
      const int c = cb + t * blockDim.x;
      if (c < s.C) {
        const float mean = f0[t] * inv_n;
        const float var = fmaxf(f1[t] * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        tab[T_A * s.C + c] = rstd * gm[t];
        tab[T_BC * s.C + c] = bt[t] - mean * rstd * gm[t];
        if (kStatsBwd) {
          tab[T_SC * s.C + c] = rstd;
          tab[T_SH * s.C + c] = -mean * rstd;
        }
        if (kApplyBwd) {
          // group means of dzg and dzg*xhat (the statistics pass accumulated them times rstd)
          const float p1 = b0[t] * inv_n / rstd;
          const float p2 = b1[t] * inv_n / rstd;
          tab[T_D * s.C + c] = rstd * (p1 - p2 * mean * rstd);
          tab[T_E * s.C + c] = rstd * rstd * p2;
        }
      }
    }
  }
}

// ---- statistics: stats[b][g] = (sum x, sum x^2)   or, for the backward, (sum dzg, sum dzg * xhat) ---------------
template <bool BWD, bool SILU>
__global__ void gn_stats_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ fstats, float* __restrict__ out_stats,
                                GnShape s, float eps) {
  extern __shared__ float tab[];                 // BWD: [4][C] (A, Bc, Sc, Sh)
  __shared__ float acc[2 * 128];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) acc[i] = 0.f;
  if (BWD) build_table<true, false>(tab, gamma, beta, fstats, nullptr, b, s, eps);
  __syncthreads();
  const int v = threadIdx.x % s.vecs, rp = threadIdx.x / s.vecs;
  const int c0 = v * 8;
  const int g0 = c0 / s.cpg, g1 = (c0 + 7) / s.cpg;
  float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
  // elements j >= js of this thread's 8-channel vector belong to the next group (cpg >= 8: at most two groups per vector)
  const int js = min(8, (g0 + 1) * s.cpg - c0);
  if (rp < s.rows_par) {
    float tA[8], tB[8], tS[8], tH[8];
    if (BWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tA[j] = tab[T_A * s.C + c0 + j]; tB[j] = tab[T_BC * s.C + c0 + j];
        tS[j] = tab[T_SC * s.C + c0 + j]; tH[j] = tab[T_SH * s.C + c0 + j];
      }
    }
    const int64_t r_begin = (int64_t)blockIdx.x * s.rows_per_cta;
    const int64_t r_end = min(r_begin + (int64_t)s.rows_per_cta, s.HW);
    const uint4* xb = reinterpret_cast<const uint4*>(x) + ((size_t)b * s.HW) * s.vecs + v;
    const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + ((size_t)b * s.HW) * s.vecs + v : nullptr;
    constexpr int U = 4;                      // independent 16-byte loads in flight per thread (x2 in the backward)
    for (int64_t r = r_begin + rp; r < r_end; r += (int64_t)U * s.rows_par) {
      uint4 xvv[U], dvv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t ru = r + (int64_t)u * s.rows_par;
        xvv[u] = make_uint4(0, 0, 0, 0);
        dvv[u] = make_uint4(0, 0, 0, 0);
        if (ru < r_end) {
          xvv[u] = ld_stream(xb + ru * s.vecs);
          if (BWD) dvv[u] = ld_stream(db + ru * s.vecs);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (int64_t)u * s.rows_par >= r_end) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xe = bf16_elem(xvv[u], j);
          float p, pq;
          if (!BWD) {
            p = xe;
            pq = xe * xe;
          } else {
            const float de = bf16_elem(dvv[u], j);
            float dz = de;
            if (SILU) {
              const float z = fmaf(xe, tA[j], tB[j]);
              const float sg = sigmoidf_(z);
              dz = de * sg * (1.0f + z * (1.0f - sg));
            }
            // dzg = dz * gamma = dz * A / rstd; accumulate dz*A and dz*A*xhat, the 1/rstd goes into the apply pass' table
            p = dz * tA[j];
            pq = p * fmaf(xe, tS[j], tH[j]);
          }
          if (j >= js) { a1 += p; q1 += pq; } else { a0 += p; q0 += pq; }
        }
      }
    }
    atomicAdd(&acc[2 * g0], a0);
    atomicAdd(&acc[2 * g0 + 1], q0);
    if (g1 != g0) { atomicAdd(&acc[2 * g1], a1); atomicAdd(&acc[2 * g1 + 1], q1); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) atomicAdd(out_stats + (size_t)b * s.G * 2 + i, acc[i]);
}

// ---- apply: forward y = act(x A + Bc);  backward dx = A dz - D - E x --------------------------------------------------
// bstats of the backward hold (sum dz A, sum dz A xhat) = rstd * (sum dzg, sum dzg xhat): P1 = bstats0 / (rstd n) etc.
template <bool BWD, bool SILU>
__global__ void gn_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ fstats, const float* __restrict__ bstats,
                                uint16_t* __restrict__ out, GnShape s, float eps) {
  extern __shared__ float tab[];                 // forward [2][C] (A, Bc); backward [4][C] (A, Bc, D, E)
  const int b = blockIdx.y;
  build_table<false, BWD>(tab, gamma, beta, fstats, bstats, b, s, eps);
  __syncthreads();
  const int64_t r_begin = (int64_t)blockIdx.x * s.rows_per_cta;
  const int64_t r_end = min(r_begin + (int64_t)s.rows_per_cta, s.HW);
  const int64_t n_vec = (r_end - r_begin) * s.vecs;          // this CTA's contiguous slab of 16-byte vectors
  const size_t base = ((size_t)b * s.HW + r_begin) * s.vecs;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + base;
  const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + base : nullptr;
  uint4* ob = reinterpret_cast<uint4*>(out) + base;
  constexpr int U = 4;
  const int step = blockDim.x;
  // channel vector of slab element i is i % vecs: advanced incrementally (no division in the loop)
  const int dv = step % s.vecs;
  int v = threadIdx.x % s.vecs;
  for (int64_t i = threadIdx.x; i < n_vec; i += (int64_t)U * step) {
    uint4 xvv[U], dvv[U];
    int vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iu = i + (int64_t)u * step;
      vv[u] = v;
      v += dv;
      if (v >= s.vecs) v -= s.vecs;
      xvv[u] = make_uint4(0, 0, 0, 0);
      dvv[u] = make_uint4(0, 0, 0, 0);
      if (iu < n_vec) {
        xvv[u] = ld_stream(xb + iu);
        if (BWD) dvv[u] = ld_stream(db + iu);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iu = i + (int64_t)u * step;
      if (iu >= n_vec) break;
      const int c0 = vv[u] * 8;
      const float4* ta = reinterpret_cast<const float4*>(tab + T_A * s.C + c0);
      const float4* tb = reinterpret_cast<const float4*>(tab + T_BC * s.C + c0);
      const float4 a0 = ta[0], a1 = ta[1], b0 = tb[0], b1 = tb[1];
      const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float Bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
      if (!BWD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(bf16_elem(xvv[u], j), A[j], Bc[j]);
          o[j] = SILU ? z * sigmoidf_(z) : z;
        }
      } else {
        const float4* td = reinterpret_cast<const float4*>(tab + T_D * s.C + c0);
        const float4* te = reinterpret_cast<const float4*>(tab + T_E * s.C + c0);
        const float4 d0 = td[0], d1 = td[1], e0 = te[0], e1 = te[1];
        const float D[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const float E[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xe = bf16_elem(xvv[u], j);
          float dz = bf16_elem(dvv[u], j);
          if (SILU) {
            const float z = fmaf(xe, A[j], Bc[j]);
            const float sg = sigmoidf_(z);
            dz = dz * sg * (1.0f + z * (1.0f - sg));
          }
          o[j] = fmaf(A[j], dz, -fmaf(E[j], xe, D[j]));
        }
      }
      st_stream(ob + iu, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
    }
  }
}

static int gn_shape(GnShape* s, int64_t B, int64_t HW, int C, int G) {
  SDT_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, SDT_ERR_ARG, "group_norm: bad sizes");
  SDT_REQUIRE(C % 8 == 0 && C / G >= 8 && G <= 128 && C / 8 <= 512, SDT_ERR_UNSUPPORTED,
              "group_norm_nhwc: needs C %% 8 == 0, C/G >= 8, G <= 128, C <= 4096 (got C=%d G=%d)", C, G);
  SDT_REQUIRE(B <= 65535, SDT_ERR_UNSUPPORTED, "group_norm_nhwc: batch too large");
  s->HW = HW; s->C = C; s->G = G; s->cpg = C / G; s->vecs = C / 8;
  // 256 threads unless one pixel row alone needs more channel-vector slots
  const int base_threads = s->vecs > 256 ? 512 : 256;
  s->rows_par = base_threads / s->vecs;
  if (s->rows_par > HW) s->rows_par = (int)HW;
  s->threads = ((s->vecs * s->rows_par + 31) / 32) * 32;
  s->rows_per_cta = 0;
  return SDT_OK;
}

// Grid of ONE wave for this kernel: (resident CTAs per SM from the occupancy calculator) x SMs, split over the batch; every
// thread still sees >= 4 vectors so that the per-CTA table is amortised.
template <typename K>
static int gn_grid(K kernel, GnShape* s, int64_t B, size_t smem, dim3* grid) {
  if (smem > 48 * 1024) SDT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  int per_sm = 0;
  SDT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, s->threads, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t want = ((int64_t)num_sms() * per_sm + B - 1) / B;
  if (want < 1) want = 1;
  int64_t rows = (s->HW + want - 1) / want;
  const int64_t min_rows = 4LL * s->rows_par;
  if (rows < min_rows) rows = min_rows;
  s->rows_per_cta = (int)rows;
  *grid = dim3((unsigned)((s->HW + rows - 1) / rows), (unsigned)B);
  return SDT_OK;
}

}  // namespace sdt

using namespace sdt;

#define SDT_GN_LAUNCH(KERNEL, SMEM, WHAT, ...)                               \
  do {                                                                       \
    GnShape sk = s;                                                          \
    dim3 grid;                                                               \
    if ((rc = gn_grid(KERNEL, &sk, B, SMEM, &grid)) != SDT_OK) return rc;    \
    KERNEL<<<grid, sk.threads, SMEM, st>>>(__VA_ARGS__, sk, eps);            \
    SDT_LAUNCH_OK(WHAT);                                                     \
  } while (0)

extern "C" int sdt_group_norm_nhwc(const void* x, const float* gamma, const float* beta, float* stats, void* y, int64_t B,
                                   int64_t HW, int C, int G, float eps, int silu, void* stream) {
  SDT_REQUIRE(x && gamma && beta && stats && y, SDT_ERR_ARG, "sdt_group_norm_nhwc: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(y), SDT_ERR_ARG, "sdt_group_norm_nhwc: pointers must be 16-byte aligned");
  GnShape s;
  int rc = gn_shape(&s, B, HW, C, G);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tab2 = sizeof(float) * 2 * C;
  const uint16_t* xp = (const uint16_t*)x;
  const uint16_t* np = nullptr;
  const float* nf = nullptr;
  SDT_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * B * G, st));
  SDT_GN_LAUNCH((gn_stats_kernel<false, false>), 0, "gn_stats", xp, np, gamma, beta, nf, stats);
  if (silu) SDT_GN_LAUNCH((gn_apply_kernel<false, true>), tab2, "gn_apply", xp, np, gamma, beta, (const float*)stats, nf, (uint16_t*)y);
  else      SDT_GN_LAUNCH((gn_apply_kernel<false, false>), tab2, "gn_apply", xp, np, gamma, beta, (const float*)stats, nf, (uint16_t*)y);
  return SDT_OK;
}

extern "C" int sdt_group_norm_nhwc_bwd(const void* x, const void* dout, const float* gamma, const float* beta,
                                       const float* stats, float* bstats, void* dx, int64_t B, int64_t HW, int C, int G,
                                       float eps, int silu, void* stream) {
  SDT_REQUIRE(x && dout && gamma && beta && stats && bstats && dx, SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(dout) && aligned16(dx), SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: pointers must be 16-byte aligned");
  GnShape s;
  int rc = gn_shape(&s, B, HW, C, G);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tab4 = sizeof(float) * 4 * C;      // <= 64 KiB (C <= 4096)
  const uint16_t* xp = (const uint16_t*)x;
  const uint16_t* dp = (const uint16_t*)dout;
  SDT_CUDA_OK(cudaMemsetAsync(bstats, 0, sizeof(float) * 2 * B * G, st));
  if (silu) {
    SDT_GN_LAUNCH((gn_stats_kernel<true, true>), tab4, "gn_bwd_stats", xp, dp, gamma, beta, stats, bstats);
    SDT_GN_LAUNCH((gn_apply_kernel<true, true>), tab4, "gn_bwd_apply", xp, dp, gamma, beta, stats, (const float*)bstats, (uint16_t*)dx);
  } else {
    SDT_GN_LAUNCH((gn_stats_kernel<true, false>), tab4, "gn_bwd_stats", xp, dp, gamma, beta, stats, bstats);
    SDT_GN_LAUNCH((gn_apply_kernel<true, false>), tab4, "gn_bwd_apply", xp, dp, gamma, beta, stats, (const float*)bstats, (uint16_t*)dx);
  }
  return SDT_OK;
}
