// f2: GroupNorm (+ SiLU) on channels-last activations, forward and backward (frozen affine parameters).
//
// torch's CUDA group_norm only takes NCHW-contiguous input, so in a channels-last UNet every GroupNorm costs two
// layout copies in the forward and two more in the backward, next to the separate SiLU kernels.  These kernels work on
// the [B, HW, C] memory directly: one statistics pass (per-(b,group) sum / sum of squares, fp32, reduced in registers
// per channel column, then shared-memory and global red.add) and one apply pass with the activation fused.
// HBM-bound: forward 2 reads + 1 write of the tensor, backward 4 reads + 1 write (the second read of each pair is
// normally an L2 hit: the tensors are tens of MB).
#include "sdt_common.cuh"

namespace sdt {

constexpr int kGnThreads = 512;

struct GnShape {
  int64_t HW;
  int C, G, cpg, vecs;          // vecs = C / 8 (16-byte vectors per pixel row)
  int rows_par;                 // pixel rows processed concurrently by one CTA
  int rows_per_cta;
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }

// ---- statistics: stats[b][g] = (sum x, sum x^2)   or, for the backward, (sum dzg, sum dzg * xhat) ---------------
template <bool BWD, bool SILU>
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ fstats, float* __restrict__ out_stats, GnShape s,
                float eps) {
  __shared__ float acc[2 * 128];
  const int b = blockIdx.y;
  const int v = threadIdx.x % s.vecs, rp = threadIdx.x / s.vecs;
  const bool active = rp < s.rows_par;
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int c0 = v * 8;
  const int g0 = c0 / s.cpg, g1 = (c0 + 7) / s.cpg;
  float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
  // elements j >= js of this thread's 8-channel vector belong to the next group (cpg >= 8: at most two groups per vector);
  // computed once -- an integer division per element in the streaming loop would dominate the kernel
  const int js = min(8, (g0 + 1) * s.cpg - c0);
  if (active) {
    float gm[8], bt[8], mean[2] = {0.f, 0.f}, rstd[2] = {0.f, 0.f};
    if (BWD) {
      const float inv_n = 1.0f / ((float)s.HW * (float)s.cpg);
#pragma unroll
      for (int j = 0; j < 8; ++j) { gm[j] = __ldg(gamma + c0 + j); bt[j] = __ldg(beta + c0 + j); }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int g = k ? g1 : g0;
        const float m = fstats[((size_t)b * s.G + g) * 2] * inv_n;
        const float var = fmaxf(fstats[((size_t)b * s.G + g) * 2 + 1] * inv_n - m * m, 0.f);
        mean[k] = m;
        rstd[k] = rsqrtf(var + eps);
      }
    }
    const int64_t r_begin = (int64_t)blockIdx.x * s.rows_per_cta;
    const int64_t r_end = min(r_begin + (int64_t)s.rows_per_cta, s.HW);
    const uint4* xb = reinterpret_cast<const uint4*>(x) + ((size_t)b * s.HW) * s.vecs + v;
    const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + ((size_t)b * s.HW) * s.vecs + v : nullptr;
    constexpr int U = 4;                      // independent 16-byte loads in flight per thread
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
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xvv[u]);
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dvv[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xe = bf16_bits_to_f32((j & 1) ? (xw[j >> 1] >> 16) : (xw[j >> 1] & 0xffffu));
          const bool second = j >= js;
          float p, pq;
          if (!BWD) {
            p = xe;
            pq = xe * xe;
          } else {
            const float de = bf16_bits_to_f32((j & 1) ? (dw[j >> 1] >> 16) : (dw[j >> 1] & 0xffffu));
            const float xhat = (xe - mean[second]) * rstd[second];
            float dz = de;
            if (SILU) {
              const float z = xhat * gm[j] + bt[j];
              const float sg = sigmoidf_(z);
              dz = de * sg * (1.0f + z * (1.0f - sg));
            }
            p = dz * gm[j];
            pq = p * xhat;
          }
          if (second) { a1 += p; q1 += pq; } else { a0 += p; q0 += pq; }
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

// ---- apply: forward y = act(xhat * gamma + beta);  backward dx = rstd (dzg - P1/n - xhat P2/n) ---------------------
template <bool BWD, bool SILU>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ fstats, const float* __restrict__ bstats,
                uint16_t* __restrict__ out, GnShape s, float eps) {
  const int b = blockIdx.y;
  const int v = threadIdx.x % s.vecs, rp = threadIdx.x / s.vecs;
  if (rp >= s.rows_par) return;
  const int c0 = v * 8;
  const int g0 = c0 / s.cpg, g1 = (c0 + 7) / s.cpg;
  const float inv_n = 1.0f / ((float)s.HW * (float)s.cpg);
  float gm[8], bt[8], mean[2], rstd[2], p1[2] = {0.f, 0.f}, p2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) { gm[j] = __ldg(gamma + c0 + j); bt[j] = __ldg(beta + c0 + j); }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int g = k ? g1 : g0;
    const float m = fstats[((size_t)b * s.G + g) * 2] * inv_n;
    const float var = fmaxf(fstats[((size_t)b * s.G + g) * 2 + 1] * inv_n - m * m, 0.f);
    mean[k] = m;
    rstd[k] = rsqrtf(var + eps);
    if (BWD) {
      p1[k] = bstats[((size_t)b * s.G + g) * 2] * inv_n;
      p2[k] = bstats[((size_t)b * s.G + g) * 2 + 1] * inv_n;
    }
  }
  // per-element affine of the normalisation, hoisted out of the streaming loop: xhat = x * sc[j] + sh[j]
  const int js = min(8, (g0 + 1) * s.cpg - c0);
  float sc[8], sh[8], q1[8], q2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = j >= js ? 1 : 0;
    sc[j] = rstd[k];
    sh[j] = -mean[k] * rstd[k];
    q1[j] = p1[k];
    q2[j] = p2[k];
  }
  const int64_t r_begin = (int64_t)blockIdx.x * s.rows_per_cta;
  const int64_t r_end = min(r_begin + (int64_t)s.rows_per_cta, s.HW);
  const uint4* xb = reinterpret_cast<const uint4*>(x) + ((size_t)b * s.HW) * s.vecs + v;
  const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + ((size_t)b * s.HW) * s.vecs + v : nullptr;
  uint4* ob = reinterpret_cast<uint4*>(out) + ((size_t)b * s.HW) * s.vecs + v;
  constexpr int U = 4;
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
      const int64_t ru = r + (int64_t)u * s.rows_par;
      if (ru >= r_end) break;
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xvv[u]);
      const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dvv[u]);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xe = bf16_bits_to_f32((j & 1) ? (xw[j >> 1] >> 16) : (xw[j >> 1] & 0xffffu));
        const float xhat = fmaf(xe, sc[j], sh[j]);
        const float z = fmaf(xhat, gm[j], bt[j]);
        if (!BWD) {
          o[j] = SILU ? z * sigmoidf_(z) : z;
        } else {
          const float de = bf16_bits_to_f32((j & 1) ? (dw[j >> 1] >> 16) : (dw[j >> 1] & 0xffffu));
          float dz = de;
          if (SILU) {
            const float sg = sigmoidf_(z);
            dz = de * sg * (1.0f + z * (1.0f - sg));
          }
          o[j] = sc[j] * (dz * gm[j] - q1[j] - xhat * q2[j]);
        }
      }
      st_stream(ob + ru * s.vecs, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
    }
  }
}

static int gn_shape(GnShape* s, int64_t B, int64_t HW, int C, int G, dim3* grid, int* threads) {
  SDT_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, SDT_ERR_ARG, "group_norm: bad sizes");
  SDT_REQUIRE(C % 8 == 0 && C / G >= 8 && G <= 128 && C / 8 <= kGnThreads, SDT_ERR_UNSUPPORTED,
              "group_norm_nhwc: needs C %% 8 == 0, C/G >= 8, G <= 128, C <= %d (got C=%d G=%d)", 8 * kGnThreads, C, G);
  SDT_REQUIRE(B <= 65535, SDT_ERR_UNSUPPORTED, "group_norm_nhwc: batch too large");
  s->HW = HW; s->C = C; s->G = G; s->cpg = C / G; s->vecs = C / 8;
  s->rows_par = kGnThreads / s->vecs;
  if (s->rows_par > HW) s->rows_par = (int)HW;
  *threads = ((s->vecs * s->rows_par + 31) / 32) * 32;
  // ~2 CTAs per SM over the batch (per-thread set-up amortised over more rows); at least 16 rows per row slot
  int64_t want = (2LL * num_sms() + B - 1) / B;
  int64_t rows = (HW + want - 1) / want;
  const int64_t min_rows = 16LL * s->rows_par;
  if (rows < min_rows) rows = min_rows;
  s->rows_per_cta = (int)rows;
  *grid = dim3((unsigned)((HW + rows - 1) / rows), (unsigned)B);
  return SDT_OK;
}

}  // namespace sdt

using namespace sdt;

extern "C" int sdt_group_norm_nhwc(const void* x, const float* gamma, const float* beta, float* stats, void* y, int64_t B,
                                   int64_t HW, int C, int G, float eps, int silu, void* stream) {
  SDT_REQUIRE(x && gamma && beta && stats && y, SDT_ERR_ARG, "sdt_group_norm_nhwc: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(y), SDT_ERR_ARG, "sdt_group_norm_nhwc: pointers must be 16-byte aligned");
  GnShape s; dim3 grid; int threads;
  int rc = gn_shape(&s, B, HW, C, G, &grid, &threads);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SDT_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * B * G, st));
  gn_stats_kernel<false, false><<<grid, threads, 0, st>>>((const uint16_t*)x, nullptr, gamma, beta, nullptr, stats, s, eps);
  SDT_LAUNCH_OK("gn_stats");
  if (silu) gn_apply_kernel<false, true><<<grid, threads, 0, st>>>((const uint16_t*)x, nullptr, gamma, beta, stats, nullptr, (uint16_t*)y, s, eps);
  else      gn_apply_kernel<false, false><<<grid, threads, 0, st>>>((const uint16_t*)x, nullptr, gamma, beta, stats, nullptr, (uint16_t*)y, s, eps);
  SDT_LAUNCH_OK("gn_apply");
  return SDT_OK;
}

extern "C" int sdt_group_norm_nhwc_bwd(const void* x, const void* dout, const float* gamma, const float* beta,
                                       const float* stats, float* bstats, void* dx, int64_t B, int64_t HW, int C, int G,
                                       float eps, int silu, void* stream) {
  SDT_REQUIRE(x && dout && gamma && beta && stats && bstats && dx, SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(dout) && aligned16(dx), SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: pointers must be 16-byte aligned");
  GnShape s; dim3 grid; int threads;
  int rc = gn_shape(&s, B, HW, C, G, &grid, &threads);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SDT_CUDA_OK(cudaMemsetAsync(bstats, 0, sizeof(float) * 2 * B * G, st));
  if (silu) {
    gn_stats_kernel<true, true><<<grid, threads, 0, st>>>((const uint16_t*)x, (const uint16_t*)dout, gamma, beta, stats, bstats, s, eps);
    SDT_LAUNCH_OK("gn_bwd_stats");
    gn_apply_kernel<true, true><<<grid, threads, 0, st>>>((const uint16_t*)x, (const uint16_t*)dout, gamma, beta, stats, bstats, (uint16_t*)dx, s, eps);
  } else {
    gn_stats_kernel<true, false><<<grid, threads, 0, st>>>((const uint16_t*)x, (const uint16_t*)dout, gamma, beta, stats, bstats, s, eps);
    SDT_LAUNCH_OK("gn_bwd_stats");
    gn_apply_kernel<true, false><<<grid, threads, 0, st>>>((const uint16_t*)x, (const uint16_t*)dout, gamma, beta, stats, bstats, (uint16_t*)dx, s, eps);
  }
  SDT_LAUNCH_OK("gn_bwd_apply");
  return SDT_OK;
}
