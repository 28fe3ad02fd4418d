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

// SiLU through the hardware tanh: with h = z/2 and t = tanh(h),  sigmoid(z) = (1 + t)/2,
//     silu(z)  = z sigmoid(z)                 = h + h t                          (FFMA)
//     silu'(z) = sigmoid (1 + z (1 - sigmoid)) = (1 + t)(1 + h - h t) / 2
// ONE special-function op per element instead of ex2 + a full-precision division (ncu: these kernels were bound by the
// special-function / issue rate, not by memory); |error| < 2^-11, below bf16 resolution.  The SiLU variants therefore keep
// HALVED affine rows in the table (h comes straight out of one FFMA) and fold the factor 2 back where A multiplies dz.
__device__ __forceinline__ float tanh_approx(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return t;
}
// 2 * silu'(2h)
__device__ __forceinline__ float silu_grad_x2(float h) {
  const float t = tanh_approx(h);
  return (1.0f + t) * (fmaf(-h, t, h) + 1.0f);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// per-channel table rows (each a float[C] in dynamic shared memory)
enum { T_A = 0, T_BC = 1, T_SC = 2, T_SH = 3, T_D = 2, T_E = 3 };

// Per-(sample, channel) table in shared memory.  The loads of TU channels per thread are all issued before the first use:
// one global-memory latency per CTA instead of C / blockDim of them (C = 2560 with 160 threads was 16 dependent rounds).
//   kStatsBwd: rows A, Bc, Sc, Sh (backward statistics pass)        kApplyBwd: rows A, Bc, D, E        neither: A, Bc
//   kHalf (SiLU variants): A and Bc are stored halved
//   cbias (optional, bf16 [B, C]): the tensor that is normalised is x + cbias[b, c] (the time-embedding add in front of norm2
//   of a ResNet block); every row is an affine function of x, so the bias folds into the constant terms
// Statistics reach the table builder either FINAL (fstats [B,G,2]: the forward's, saved for the backward) or as the per-CTA
// PARTIALS of the statistics launch just before this one (part [B][n_part][2G]); partials are summed here in CTA order, by
// every CTA of the apply launch for itself -- deterministic (no float atomics anywhere), no zero-fill, nothing on the critical
// path of the statistics launch.  CTA (0, b) of the forward apply also writes the final statistics out for the backward.
struct GnStats {
  const float* fstats;      // final forward statistics, or null
  const float* fpart;       // forward partials (used when fstats is null)
  const float* bpart;       // backward partials (apply-backward only)
  float* stats_out;         // where CTA (0, b) stores the summed forward statistics (or null)
  int n_fpart, n_bpart;
};
__device__ __forceinline__ float sum_partials(const float* __restrict__ part, int n_part, int b, int two_g, int i) {
  const float* p = part + (size_t)b * n_part * two_g + i;
  float acc = 0.f;
  int k = 0;
  for (; k + 4 <= n_part; k += 4) {
    const float v0 = __ldcg(p + (size_t)(k + 0) * two_g), v1 = __ldcg(p + (size_t)(k + 1) * two_g);
    const float v2 = __ldcg(p + (size_t)(k + 2) * two_g), v3 = __ldcg(p + (size_t)(k + 3) * two_g);
    acc += v0; acc += v1; acc += v2; acc += v3;
  }
  for (; k < n_part; ++k) acc += __ldcg(p + (size_t)k * two_g);
  return acc;
}
template <bool kStatsBwd, bool kApplyBwd, bool kHalf>
__device__ __forceinline__ void build_table(float* tab, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const GnStats& gs, const uint16_t* __restrict__ cbias, int b, const GnShape& s, float eps) {
  constexpr int TU = 8;
  const float inv_n = 1.0f / ((float)s.HW * (float)s.cpg);
  const float ab = kHalf ? 0.5f : 1.0f;
  __shared__ float gsum[4 * 128];                 // [0, 2G): forward (sum, sum of squares) per group; [2G, 4G): backward sums
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) {
    float f;
    if (gs.fstats != nullptr) f = gs.fstats[(size_t)b * 2 * s.G + i];
    else {
      f = sum_partials(gs.fpart, gs.n_fpart, b, 2 * s.G, i);
      if (gs.stats_out != nullptr && blockIdx.x == 0) gs.stats_out[(size_t)b * 2 * s.G + i] = f;
    }
    gsum[i] = f;
    if (kApplyBwd) gsum[2 * s.G + i] = sum_partials(gs.bpart, gs.n_bpart, b, 2 * s.G, i);
  }
  __syncthreads();
  for (int cb = threadIdx.x; cb < s.C; cb += TU * blockDim.x) {
    float gm[TU], bt[TU], f0[TU], f1[TU], b0[TU], b1[TU], tb[TU];
#pragma unroll
    for (int t = 0; t < TU; ++t) {
      const int c = cb + t * blockDim.x;
      if (c < s.C) {
        const int gi = (c / s.cpg) * 2;
        gm[t] = __ldg(gamma + c);
        bt[t] = __ldg(beta + c);
        tb[t] = cbias != nullptr ? __uint_as_float((uint32_t)cbias[(size_t)b * s.C + c] << 16) : 0.f;
        f0[t] = gsum[gi];
        f1[t] = gsum[gi + 1];
        if (kApplyBwd) { b0[t] = gsum[2 * s.G + gi]; b1[t] = gsum[2 * s.G + gi + 1]; }
      }
    }
#pragma unroll
    for (int t = 0; t < TU; ++t) {
      const int c = cb + t * blockDim.x;
      if (c < s.C) {
        const float mean = f0[t] * inv_n;
        const float var = fmaxf(f1[t] * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        // with x' = x + tb:  z = x' A + Bc = x A + (Bc + tb A);  xhat = x Sc + (Sh + tb Sc);  dx = A dz - (D + E tb) - E x
        tab[T_A * s.C + c] = ab * rstd * gm[t];
        tab[T_BC * s.C + c] = ab * (bt[t] + (tb[t] - mean) * rstd * gm[t]);
        if (kStatsBwd) {
          tab[T_SC * s.C + c] = rstd;
          tab[T_SH * s.C + c] = (tb[t] - mean) * rstd;
        }
        if (kApplyBwd) {
          // group means of dzg and dzg*xhat (the statistics pass accumulated them times rstd)
          const float p1 = b0[t] * inv_n / rstd;
          const float p2 = b1[t] * inv_n / rstd;
          tab[T_D * s.C + c] = rstd * (p1 + p2 * (tb[t] - mean) * rstd);
          tab[T_E * s.C + c] = rstd * rstd * p2;
        }
      }
    }
  }
}

// ---- statistics: stats[b][g] = (sum x, sum x^2)   or, for the backward, (sum dz A, sum dz A xhat) = rstd (sum dzg, sum dzg xhat)
template <bool BWD, bool SILU>
__global__ void gn_stats_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ fstats, float* __restrict__ out_part,
                                const uint16_t* __restrict__ cbias, GnShape s, float eps) {
  extern __shared__ float tab[];                 // BWD: [4][C] (A, Bc, Sc, Sh)
  __shared__ float4 tpart[512];                  // per-thread (sum, sq) of its first / second group: summed in thread order below
  const int b = blockIdx.y;
  pdl_wait();                 // PDL (sdt_common.cuh): x / dout come from the kernel in front of us
  pdl_launch_dependents();
  if (BWD) {
    GnStats gs{fstats, nullptr, nullptr, nullptr, 0, 0};
    build_table<true, false, SILU>(tab, gamma, beta, gs, cbias, b, s, eps);
  }
  __syncthreads();
  const int v = threadIdx.x % s.vecs, rp = threadIdx.x / s.vecs;
  const int c0 = v * 8;
  if (rp < s.rows_par) {
    float tA[8], tB[8], tS[8], tH[8], tb[8];
    if (!BWD) {                 // forward statistics of x + cbias
      uint4 tbv = make_uint4(0, 0, 0, 0);
      if (cbias != nullptr) tbv = *reinterpret_cast<const uint4*>(cbias + (size_t)b * s.C + c0);
      unpack8(tbv, tb);
    }
    if (BWD) {
      load8(tab + T_A * s.C + c0, tA);
      load8(tab + T_BC * s.C + c0, tB);
      load8(tab + T_SC * s.C + c0, tS);
      load8(tab + T_SH * s.C + c0, tH);
    }
    // one accumulator pair per element position: the streaming loop is pure FMAs, the split into groups happens once at the end
    float S[8], Q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { S[j] = 0.f; Q[j] = 0.f; }
    const int r_begin = blockIdx.x * s.rows_per_cta;
    const int n_rows = min(s.rows_per_cta, (int)(s.HW - r_begin));
    // this thread's rows are rp, rp + rows_par, ...: consecutive ones are rstep vectors apart in the slab
    const size_t base = ((size_t)b * s.HW + r_begin + rp) * s.vecs + v;
    const uint4* xb = reinterpret_cast<const uint4*>(x) + base;
    const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + base : nullptr;
    const int rstep = s.rows_par * s.vecs;
    const int nk = rp < n_rows ? (n_rows - rp + s.rows_par - 1) / s.rows_par : 0;
    auto accumulate = [&](const uint4& xu, const uint4& du) {
      float xe[8], de[8];
      unpack8(xu, xe);
      if (BWD) unpack8(du, de);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!BWD) {
          const float xb = xe[j] + tb[j];
          S[j] += xb;
          Q[j] = fmaf(xb, xb, Q[j]);
        } else {
          float dz = de[j];
          if (SILU) dz *= silu_grad_x2(fmaf(xe[j], tA[j], tB[j]));      // 2 dz, against the halved A
          const float p = dz * tA[j];
          S[j] += p;
          Q[j] = fmaf(p, fmaf(xe[j], tS[j], tH[j]), Q[j]);
        }
      }
    };
    constexpr int U = 4;                      // independent 16-byte loads in flight per thread (x2 in the backward)
    int k = 0;
    for (; k + U <= nk; k += U) {
      uint4 xvv[U], dvv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xvv[u] = ld_stream(xb + (k + u) * rstep);
        if (BWD) dvv[u] = ld_stream(db + (k + u) * rstep);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) accumulate(xvv[u], BWD ? dvv[u] : xvv[u]);
    }
    for (; k < nk; ++k) {
      const uint4 xu = ld_stream(xb + k * rstep);
      const uint4 du = BWD ? ld_stream(db + k * rstep) : xu;
      accumulate(xu, du);
    }
    // elements j >= js of this thread's 8-channel vector belong to the next group (cpg >= 8: at most two groups per vector)
    const int g0 = c0 / s.cpg;
    const int js = min(8, (g0 + 1) * s.cpg - c0);
    float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < js) { a0 += S[j]; q0 += Q[j]; } else { a1 += S[j]; q1 += Q[j]; }
    }
    tpart[threadIdx.x] = make_float4(a0, q0, a1, q1);
  }
  __syncthreads();
  // group totals of this CTA in a FIXED order (pixel-row slot, then channel vector): thread i owns value i of [G][2]; the CTA's
  // partial goes to out_part[b][blockIdx.x][2G] and is summed, in CTA order, by whoever consumes the statistics (build_table)
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) {
    const int g = i >> 1, sq = i & 1;
    const int v_lo = (g * s.cpg) / 8, v_hi = min(s.vecs - 1, ((g + 1) * s.cpg - 1) / 8);
    float tot = 0.f;
    for (int r = 0; r < s.rows_par; ++r)
      for (int vv = v_lo; vv <= v_hi; ++vv) {
        const float4 t = tpart[r * s.vecs + vv];
        const int gg0 = (vv * 8) / s.cpg, gg1 = (vv * 8 + 7) / s.cpg;
        if (gg0 == g) tot += sq ? t.y : t.x;
        else if (gg1 == g) tot += sq ? t.w : t.z;
      }
    out_part[((size_t)b * gridDim.x + blockIdx.x) * 2 * s.G + i] = tot;
  }
}

// ---- apply: forward y = act(x A + Bc);  backward dx = A dz - D - E x --------------------------------------------------
// bstats of the backward hold (sum dz A, sum dz A xhat) = rstd * (sum dzg, sum dzg xhat): P1 = bstats0 / (rstd n) etc.
template <bool BWD, bool SILU>
__global__ void gn_apply_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dout, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const GnStats gs,
                                uint16_t* __restrict__ out, const uint16_t* __restrict__ cbias, GnShape s, float eps) {
  extern __shared__ float tab[];                 // forward [2][C] (A, Bc); backward [4][C] (A, Bc, D, E); A, Bc halved for SiLU
  const int b = blockIdx.y;
  pdl_wait();                 // PDL: the statistics (and x) come from the kernels in front of us
  pdl_launch_dependents();
  build_table<false, BWD, SILU>(tab, gamma, beta, gs, cbias, b, s, eps);
  __syncthreads();
  const int r_begin = blockIdx.x * s.rows_per_cta;
  const int n_rows = min(s.rows_per_cta, (int)(s.HW - r_begin));
  const int n_vec = n_rows * s.vecs;                         // this CTA's contiguous slab of 16-byte vectors
  const size_t base = ((size_t)b * s.HW + r_begin) * s.vecs;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + base;
  const uint4* db = BWD ? reinterpret_cast<const uint4*>(dout) + base : nullptr;
  uint4* ob = reinterpret_cast<uint4*>(out) + base;
  auto process = [&](int i, int vv, const uint4& xu, const uint4& du) {
    const int c0 = vv * 8;
    float A[8], Bc[8], xe[8], o[8];
    load8(tab + T_A * s.C + c0, A);
    load8(tab + T_BC * s.C + c0, Bc);
    unpack8(xu, xe);
    if (!BWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = fmaf(xe[j], A[j], Bc[j]);              // SiLU: z / 2, otherwise z
        o[j] = SILU ? fmaf(h, tanh_approx(h), h) : h;
      }
    } else {
      float D[8], E[8], de[8];
      load8(tab + T_D * s.C + c0, D);
      load8(tab + T_E * s.C + c0, E);
      unpack8(du, de);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = de[j];
        if (SILU) dz *= silu_grad_x2(fmaf(xe[j], A[j], Bc[j]));
        o[j] = fmaf(A[j], dz, -fmaf(E[j], xe[j], D[j]));
      }
    }
    st_stream(ob + i, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
  };
  constexpr int U = 4;
  const int step = blockDim.x;
  // channel vector of slab element i is i % vecs: advanced incrementally (no division in the loop)
  const int dv = step % s.vecs;
  int v = threadIdx.x % s.vecs;
  int i = threadIdx.x;
  for (; i + (U - 1) * step < n_vec; i += U * step) {
    uint4 xvv[U], dvv[U];
    int vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      vv[u] = v;
      v += dv;
      if (v >= s.vecs) v -= s.vecs;
      xvv[u] = ld_stream(xb + i + u * step);
      if (BWD) dvv[u] = ld_stream(db + i + u * step);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) process(i + u * step, vv[u], xvv[u], BWD ? dvv[u] : xvv[u]);
  }
  for (; i < n_vec; i += step) {
    const uint4 xu = ld_stream(xb + i);
    const uint4 du = BWD ? ld_stream(db + i) : xu;
    process(i, v, xu, du);
    v += dv;
    if (v >= s.vecs) v -= s.vecs;
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
  // static shared memory (per-thread partials, group sums: ~10 KB) counts against the 48 KB default limit too
  if (smem > 32 * 1024) SDT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
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

// one launch: grid from the occupancy calculator, PDL attribute; *n_ctas_x receives gridDim.x (the number of partials per sample
// a statistics launch produces)
#define SDT_GN_LAUNCH(KERNEL, SMEM, WHAT, NX, ...)                           \
  do {                                                                       \
    GnShape sk = s;                                                          \
    dim3 grid;                                                               \
    if ((rc = gn_grid(KERNEL, &sk, B, SMEM, &grid)) != SDT_OK) return rc;    \
    if ((NX) != nullptr) *(NX) = (int)grid.x;                                \
    SDT_CUDA_OK(launch_kernel(KERNEL, grid, dim3(sk.threads), SMEM, st, true, __VA_ARGS__, cb, sk, eps)); \
    SDT_LAUNCH_OK(WHAT);                                                     \
  } while (0)

// floats of partial-statistics workspace that any launch for this batch size / group count can need (grid.x <= 16 CTAs per SM)
extern "C" int64_t sdt_group_norm_workspace_floats(int64_t B, int G) {
  return ((int64_t)num_sms() * 16 + B) * 2 * G;
}

extern "C" int sdt_group_norm_nhwc(const void* x, const void* chan_bias, const float* gamma, const float* beta, float* stats, void* y,
                                   int64_t B, int64_t HW, int C, int G, float eps, int silu, float* ws, int64_t ws_floats,
                                   void* stream) {
  SDT_REQUIRE(x && gamma && beta && stats && y && ws, SDT_ERR_ARG, "sdt_group_norm_nhwc: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(y) && aligned16(chan_bias), SDT_ERR_ARG, "sdt_group_norm_nhwc: pointers must be 16-byte aligned");
  const uint16_t* cb = (const uint16_t*)chan_bias;
  GnShape s;
  int rc = gn_shape(&s, B, HW, C, G);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tab2 = sizeof(float) * 2 * C;
  const uint16_t* xp = (const uint16_t*)x;
  const uint16_t* np = nullptr;
  const float* nf = nullptr;
  {   // the statistics launch writes grid.x partials per sample: check the workspace before launching
    GnShape sk = s;
    dim3 grid;
    if ((rc = gn_grid(gn_stats_kernel<false, false>, &sk, B, 0, &grid)) != SDT_OK) return rc;
    SDT_REQUIRE((int64_t)grid.x * B * 2 * G <= ws_floats, SDT_ERR_ARG, "sdt_group_norm_nhwc: workspace of %lld floats, %lld needed",
                (long long)ws_floats, (long long)grid.x * B * 2 * G);
  }
  int n_part = 0;
  SDT_GN_LAUNCH((gn_stats_kernel<false, false>), 0, "gn_stats", &n_part, xp, np, gamma, beta, nf, ws);
  const GnStats gs{nullptr, ws, nullptr, stats, n_part, 0};       // sum the partials; CTA (0, b) stores the totals for the backward
  int* none = nullptr;
  if (silu) SDT_GN_LAUNCH((gn_apply_kernel<false, true>), tab2, "gn_apply", none, xp, np, gamma, beta, gs, (uint16_t*)y);
  else      SDT_GN_LAUNCH((gn_apply_kernel<false, false>), tab2, "gn_apply", none, xp, np, gamma, beta, gs, (uint16_t*)y);
  return SDT_OK;
}

extern "C" int sdt_group_norm_nhwc_bwd(const void* x, const void* chan_bias, const void* dout, const float* gamma, const float* beta,
                                       const float* stats, float* ws, int64_t ws_floats, void* dx, int64_t B, int64_t HW, int C, int G,
                                       float eps, int silu, void* stream) {
  SDT_REQUIRE(x && dout && gamma && beta && stats && ws && dx, SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: null pointer");
  SDT_REQUIRE(aligned16(x) && aligned16(dout) && aligned16(dx) && aligned16(chan_bias), SDT_ERR_ARG,
              "sdt_group_norm_nhwc_bwd: pointers must be 16-byte aligned");
  const uint16_t* cb = (const uint16_t*)chan_bias;
  GnShape s;
  int rc = gn_shape(&s, B, HW, C, G);
  if (rc != SDT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tab4 = sizeof(float) * 4 * C;      // <= 64 KiB (C <= 4096)
  const uint16_t* xp = (const uint16_t*)x;
  const uint16_t* dp = (const uint16_t*)dout;
  SDT_REQUIRE(sdt_group_norm_workspace_floats(B, G) <= ws_floats, SDT_ERR_ARG, "sdt_group_norm_nhwc_bwd: workspace of %lld floats, %lld needed",
              (long long)ws_floats, (long long)sdt_group_norm_workspace_floats(B, G));
  int n_part = 0;
  int* none = nullptr;
  if (silu) {
    SDT_GN_LAUNCH((gn_stats_kernel<true, true>), tab4, "gn_bwd_stats", &n_part, xp, dp, gamma, beta, stats, ws);
    const GnStats gs{stats, nullptr, ws, nullptr, 0, n_part};
    SDT_GN_LAUNCH((gn_apply_kernel<true, true>), tab4, "gn_bwd_apply", none, xp, dp, gamma, beta, gs, (uint16_t*)dx);
  } else {
    SDT_GN_LAUNCH((gn_stats_kernel<true, false>), tab4, "gn_bwd_stats", &n_part, xp, dp, gamma, beta, stats, ws);
    const GnStats gs{stats, nullptr, ws, nullptr, 0, n_part};
    SDT_GN_LAUNCH((gn_apply_kernel<true, false>), tab4, "gn_bwd_apply", none, xp, dp, gamma, beta, gs, (uint16_t*)dx);
  }
  return SDT_OK;
}
