// Elementwise / reduction kernels of the LoRA train-step path (sm_100a):
//   K3 sdt_noise_target   -- DDPM add_noise + v-target      (modules/model.py:302,306-314)
//   K4 sdt_mse_loss       -- MSE, two-segment mean, dPred   (modules/model.py:316,338-342)
//   K5 sdt_ema_update_*   -- EMA lerp                       (modules/ema.py:56-61)
//   f1 sdt_adamw_flat     -- AdamW over the flat LoRA arena (modules/model.py:33-64)
//      sdt_lora_pack      -- f32 LoRA masters -> bf16 tensor-core operand layouts
// All are HBM-bound: 128-bit coalesced streaming accesses, grid sized in multiples of the SM
// count with a grid-stride loop, warp-shuffle reductions, no float atomics in the loss.
#include "sdt_common.cuh"

namespace sdt {

constexpr int kThreads = 256;

static inline int grid_for(int64_t work_items, int per_block, int ctas_per_sm) {
  int64_t want = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// =============================================================================================
// K3: noising + target
// =============================================================================================
template <bool BF16> struct Coef;
template <> struct Coef<false> {   // f32: a = sqrt(abar), s = sqrt(1 - abar)
  static __device__ __forceinline__ void get(float ac, float& a, float& s) {
    a = __fsqrt_rn(ac);
    s = __fsqrt_rn(__fsub_rn(1.0f, ac));
  }
};
template <> struct Coef<true> {    // bf16: every torch op rounds its result to bf16
  static __device__ __forceinline__ void get(float ac, float& a, float& s) {
    float acb = round_bf16(ac);
    a = round_bf16(__fsqrt_rn(acb));
    s = round_bf16(__fsqrt_rn(round_bf16(__fsub_rn(1.0f, acb))));
  }
};

__device__ __forceinline__ int clamp_t(int64_t t, int T, int32_t* oob) {
  if (t < 0 || t >= T) {
    if (oob) *oob = 1;
    t = t < 0 ? 0 : T - 1;
  }
  return (int)t;
}

// f32, 4 elements / 16 bytes per access.
template <bool WITH_V>
__global__ void __launch_bounds__(kThreads)
noise_target_f32_kernel(const float* __restrict__ x0, const float* __restrict__ eps, const int64_t* __restrict__ t,
                        const float* __restrict__ abar, int T, float* __restrict__ noisy, float* __restrict__ target,
                        int64_t nvec, int64_t chw_vec, int32_t* oob) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / chw_vec;
    float a, s;
    Coef<false>::get(__ldg(abar + clamp_t(__ldg(t + b), T, oob)), a, s);
    uint4 xv = ld_stream(reinterpret_cast<const uint4*>(x0) + i);
    uint4 ev = ld_stream(reinterpret_cast<const uint4*>(eps) + i);
    const float* x = reinterpret_cast<const float*>(&xv);
    const float* e = reinterpret_cast<const float*>(&ev);
    uint4 nv, tv;
    float* n = reinterpret_cast<float*>(&nv);
    float* v = reinterpret_cast<float*>(&tv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      n[j] = __fadd_rn(__fmul_rn(a, x[j]), __fmul_rn(s, e[j]));
      if (WITH_V) v[j] = __fsub_rn(__fmul_rn(a, e[j]), __fmul_rn(s, x[j]));
    }
    st_stream(reinterpret_cast<uint4*>(noisy) + i, nv);
    if (WITH_V) st_stream(reinterpret_cast<uint4*>(target) + i, tv);
  }
}

// bf16, 8 elements / 16 bytes per access.
template <bool WITH_V>
__global__ void __launch_bounds__(kThreads)
noise_target_bf16_kernel(const uint16_t* __restrict__ x0, const uint16_t* __restrict__ eps, const int64_t* __restrict__ t,
                         const float* __restrict__ abar, int T, uint16_t* __restrict__ noisy, uint16_t* __restrict__ target,
                         int64_t nvec, int64_t chw_vec, int32_t* oob) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / chw_vec;
    float a, s;
    Coef<true>::get(__ldg(abar + clamp_t(__ldg(t + b), T, oob)), a, s);
    uint4 xv = ld_stream(reinterpret_cast<const uint4*>(x0) + i);
    uint4 ev = ld_stream(reinterpret_cast<const uint4*>(eps) + i);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xv);
    const uint32_t* ew = reinterpret_cast<const uint32_t*>(&ev);
    uint4 nv, tv;
    uint32_t* nw = reinterpret_cast<uint32_t*>(&nv);
    uint32_t* tw = reinterpret_cast<uint32_t*>(&tv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xl = bf16_bits_to_f32(xw[j] & 0xffffu), xh = bf16_bits_to_f32(xw[j] >> 16);
      float el = bf16_bits_to_f32(ew[j] & 0xffffu), eh = bf16_bits_to_f32(ew[j] >> 16);
      float nl = __fadd_rn(round_bf16(__fmul_rn(a, xl)), round_bf16(__fmul_rn(s, el)));
      float nh = __fadd_rn(round_bf16(__fmul_rn(a, xh)), round_bf16(__fmul_rn(s, eh)));
      nw[j] = pack_bf16x2(nl, nh);
      if (WITH_V) {
        float vl = __fsub_rn(round_bf16(__fmul_rn(a, el)), round_bf16(__fmul_rn(s, xl)));
        float vh = __fsub_rn(round_bf16(__fmul_rn(a, eh)), round_bf16(__fmul_rn(s, xh)));
        tw[j] = pack_bf16x2(vl, vh);
      }
    }
    st_stream(reinterpret_cast<uint4*>(noisy) + i, nv);
    if (WITH_V) st_stream(reinterpret_cast<uint4*>(target) + i, tv);
  }
}

// scalar variant for chw not divisible by the vector width (never the case for latents, kept so that
// odd shapes are served by CUDA rather than rejected)
template <bool BF16, bool WITH_V>
__global__ void __launch_bounds__(kThreads)
noise_target_scalar_kernel(const void* __restrict__ x0, const void* __restrict__ eps, const int64_t* __restrict__ t,
                           const float* __restrict__ abar, int T, void* __restrict__ noisy, void* __restrict__ target,
                           int64_t n, int64_t chw, int32_t* oob) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / chw;
    float a, s;
    Coef<BF16>::get(__ldg(abar + clamp_t(__ldg(t + b), T, oob)), a, s);
    float x, e;
    if (BF16) {
      x = bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(x0)[i]);
      e = bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(eps)[i]);
      float nn = __fadd_rn(round_bf16(__fmul_rn(a, x)), round_bf16(__fmul_rn(s, e)));
      reinterpret_cast<uint16_t*>(noisy)[i] = (uint16_t)f32_to_bf16_bits(nn);
      if (WITH_V) {
        float vv = __fsub_rn(round_bf16(__fmul_rn(a, e)), round_bf16(__fmul_rn(s, x)));
        reinterpret_cast<uint16_t*>(target)[i] = (uint16_t)f32_to_bf16_bits(vv);
      }
    } else {
      x = reinterpret_cast<const float*>(x0)[i];
      e = reinterpret_cast<const float*>(eps)[i];
      reinterpret_cast<float*>(noisy)[i] = __fadd_rn(__fmul_rn(a, x), __fmul_rn(s, e));
      if (WITH_V) reinterpret_cast<float*>(target)[i] = __fsub_rn(__fmul_rn(a, e), __fmul_rn(s, x));
    }
  }
}

// =============================================================================================
// K4: MSE loss
// =============================================================================================
constexpr int kLossMaxBlocks = 1024;
struct LossWorkspace {
  float partial[kLossMaxBlocks][2];
  unsigned int counter;
  unsigned int pad[3];
};

template <typename T> struct Load8;
template <> struct Load8<float> {
  static __device__ __forceinline__ void ld(const void* base, int64_t i, float* out) {
    const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * i;
    uint4 a = ld_stream(p), b = ld_stream(p + 1);
    out[0] = __uint_as_float(a.x); out[1] = __uint_as_float(a.y); out[2] = __uint_as_float(a.z); out[3] = __uint_as_float(a.w);
    out[4] = __uint_as_float(b.x); out[5] = __uint_as_float(b.y); out[6] = __uint_as_float(b.z); out[7] = __uint_as_float(b.w);
  }
  static __device__ __forceinline__ void st(void* base, int64_t i, const float* v) {
    uint4* p = reinterpret_cast<uint4*>(base) + 2 * i;
    st_stream(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
    st_stream(p + 1, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
  }
  static __device__ __forceinline__ float ld1(const void* base, int64_t i) { return reinterpret_cast<const float*>(base)[i]; }
  static __device__ __forceinline__ void st1(void* base, int64_t i, float v) { reinterpret_cast<float*>(base)[i] = v; }
};
template <> struct Load8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const void* base, int64_t i, float* out) {
    uint4 a = ld_stream(reinterpret_cast<const uint4*>(base) + i);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) { out[2 * j] = bf16_bits_to_f32(w[j] & 0xffffu); out[2 * j + 1] = bf16_bits_to_f32(w[j] >> 16); }
  }
  static __device__ __forceinline__ void st(void* base, int64_t i, const float* v) {
    st_stream(reinterpret_cast<uint4*>(base) + i,
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
  }
  static __device__ __forceinline__ float ld1(const void* base, int64_t i) {
    return bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(base)[i]);
  }
  static __device__ __forceinline__ void st1(void* base, int64_t i, float v) {
    reinterpret_cast<uint16_t*>(base)[i] = (uint16_t)f32_to_bf16_bits(v);
  }
};

template <> struct Load8<__half> {      // fp16 predictions (the reference's stock `trainer.precision: 16`)
  static __device__ __forceinline__ void ld(const void* base, int64_t i, float* out) {
    uint4 a = ld_stream(reinterpret_cast<const uint4*>(base) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); out[2 * j] = f.x; out[2 * j + 1] = f.y; }
  }
  static __device__ __forceinline__ void st(void* base, int64_t i, const float* v) {
    st_stream(reinterpret_cast<uint4*>(base) + i,
              make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7])));
  }
  static __device__ __forceinline__ float ld1(const void* base, int64_t i) { return __half2float(reinterpret_cast<const __half*>(base)[i]); }
  static __device__ __forceinline__ void st1(void* base, int64_t i, float v) { reinterpret_cast<__half*>(base)[i] = __float2half_rn(v); }
};

// VEC = 8 (128-bit path) or 1 (odd shapes).  One block-level partial per segment, the last block to
// finish adds the partials in index order: the result does not depend on scheduling.
template <typename PT, typename TT, int VEC>
__global__ void __launch_bounds__(kThreads)
mse_loss_kernel(const void* __restrict__ pred, const void* __restrict__ target, float* __restrict__ loss_out,
                void* __restrict__ dpred, float* __restrict__ loss_elem, int32_t* __restrict__ nan_flag,
                int64_t nunits, int64_t split_units, float coef0, float coef1, float inv_n0, float inv_n1, float w_prior,
                float grad_scale, LossWorkspace* __restrict__ ws) {
  float acc0 = 0.f, acc1 = 0.f;
  bool saw_nan = false;
  const float g0 = 2.0f * grad_scale * coef0, g1 = 2.0f * grad_scale * coef1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nunits; i += (int64_t)gridDim.x * blockDim.x) {
    const bool second = i >= split_units;
    float p[VEC], q[VEC], d[VEC], l[VEC];
    if (VEC == 8) { Load8<PT>::ld(pred, i, p); Load8<TT>::ld(target, i, q); }
    else { p[0] = Load8<PT>::ld1(pred, i); q[0] = Load8<TT>::ld1(target, i); }
    float local = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      d[j] = p[j] - q[j];
      l[j] = d[j] * d[j];
      local += l[j];
      saw_nan |= (l[j] != l[j]);
    }
    if (second) acc1 += local; else acc0 += local;
    if (dpred != nullptr) {
      const float g = second ? g1 : g0;
#pragma unroll
      for (int j = 0; j < VEC; ++j) d[j] *= g;
      if (VEC == 8) Load8<PT>::st(dpred, i, d); else Load8<PT>::st1(dpred, i, d[0]);
    }
    if (loss_elem != nullptr) {
      if (VEC == 8) Load8<float>::st(loss_elem, i, l); else loss_elem[i] = l[0];
    }
  }
  if (saw_nan && nan_flag != nullptr) *nan_flag = 1;

  __shared__ float s0[kThreads / 32], s1[kThreads / 32];
  __shared__ bool is_last;
  acc0 = warp_sum(acc0);
  acc1 = warp_sum(acc1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s0[warp] = acc0; s1[warp] = acc1; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < kThreads / 32 ? s0[lane] : 0.f, b = lane < kThreads / 32 ? s1[lane] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      ws->partial[blockIdx.x][0] = a;
      ws->partial[blockIdx.x][1] = b;
      __threadfence();
      unsigned int done = atomicAdd(&ws->counter, 1u);
      is_last = (done == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    float a = 0.f, b = 0.f;
    for (int i = lane; i < (int)gridDim.x; i += 32) {   // fixed order: lane-strided, then shuffle tree
      a += __ldcg(&ws->partial[i][0]);
      b += __ldcg(&ws->partial[i][1]);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      const float m0 = a * inv_n0, m1 = b * inv_n1;
      loss_out[0] = m0 + w_prior * m1;
      loss_out[1] = m0;
      loss_out[2] = m1;
      ws->counter = 0;   // leave the workspace ready for the next launch
    }
  }
}

// =============================================================================================
// K5: EMA
// =============================================================================================
__device__ __forceinline__ float ema_f32(float s, float p, float omd) {
  return __fsub_rn(s, __fmul_rn(__fsub_rn(s, p), omd));            // tmp=s-p; tmp*=omd; s-=tmp
}
__device__ __forceinline__ float ema_bf16(float s, float p, float omd) {
  return round_bf16(__fsub_rn(s, round_bf16(__fmul_rn(round_bf16(__fsub_rn(s, p)), omd))));
}
__device__ __forceinline__ uint4 ema_vec_f32(uint4 sv, uint4 pv, float omd) {
  uint4 r;
  r.x = __float_as_uint(ema_f32(__uint_as_float(sv.x), __uint_as_float(pv.x), omd));
  r.y = __float_as_uint(ema_f32(__uint_as_float(sv.y), __uint_as_float(pv.y), omd));
  r.z = __float_as_uint(ema_f32(__uint_as_float(sv.z), __uint_as_float(pv.z), omd));
  r.w = __float_as_uint(ema_f32(__uint_as_float(sv.w), __uint_as_float(pv.w), omd));
  return r;
}
__device__ __forceinline__ uint32_t ema_pair_bf16(uint32_t s, uint32_t p, float omd) {
  float lo = ema_bf16(bf16_bits_to_f32(s & 0xffffu), bf16_bits_to_f32(p & 0xffffu), omd);
  float hi = ema_bf16(bf16_bits_to_f32(s >> 16), bf16_bits_to_f32(p >> 16), omd);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ uint4 ema_vec_bf16(uint4 sv, uint4 pv, float omd) {
  return make_uint4(ema_pair_bf16(sv.x, pv.x, omd), ema_pair_bf16(sv.y, pv.y, omd),
                    ema_pair_bf16(sv.z, pv.z, omd), ema_pair_bf16(sv.w, pv.w, omd));
}

// One contiguous range [0,n) of elements; 4 independent 128-bit loads per array in flight per thread.
template <bool BF16>
__device__ __forceinline__ void ema_range(void* __restrict__ shadow, const void* __restrict__ param, int64_t n, float omd,
                                          int64_t tid, int64_t nthreads) {
  constexpr int EPV = BF16 ? 8 : 4;     // elements per 16-byte vector
  constexpr int UNROLL = 4;
  uint4* s = reinterpret_cast<uint4*>(shadow);
  const uint4* p = reinterpret_cast<const uint4*>(param);
  const int64_t nvec = n / EPV;
  int64_t i = tid;
  for (; i + (UNROLL - 1) * nthreads < nvec; i += UNROLL * nthreads) {
    uint4 sv[UNROLL], pv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { sv[u] = ld_rw(s + i + u * nthreads); pv[u] = ld_stream(p + i + u * nthreads); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      st_stream(s + i + u * nthreads, BF16 ? ema_vec_bf16(sv[u], pv[u], omd) : ema_vec_f32(sv[u], pv[u], omd));
  }
  for (; i < nvec; i += nthreads) {
    uint4 sv = ld_rw(s + i), pv = ld_stream(p + i);
    st_stream(s + i, BF16 ? ema_vec_bf16(sv, pv, omd) : ema_vec_f32(sv, pv, omd));
  }
  for (int64_t k = nvec * EPV + tid; k < n; k += nthreads) {   // tail (< one vector)
    if (BF16) {
      uint16_t* ss = reinterpret_cast<uint16_t*>(shadow);
      const uint16_t* pp = reinterpret_cast<const uint16_t*>(param);
      ss[k] = (uint16_t)f32_to_bf16_bits(ema_bf16(bf16_bits_to_f32(ss[k]), bf16_bits_to_f32(pp[k]), omd));
    } else {
      float* ss = reinterpret_cast<float*>(shadow);
      const float* pp = reinterpret_cast<const float*>(param);
      ss[k] = ema_f32(ss[k], pp[k], omd);
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads)
ema_flat_kernel(void* __restrict__ shadow, const void* __restrict__ param, int64_t n, float omd,
                const float* __restrict__ omd_dev) {
  if (omd_dev != nullptr) omd = __ldg(omd_dev);
  ema_range<BF16>(shadow, param, n, omd, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads)
ema_multi_kernel(void* const* __restrict__ shadow_ptrs, const void* const* __restrict__ param_ptrs,
                 const int64_t* __restrict__ numels, const sdt_chunk* __restrict__ chunks, int n_chunks, int chunk_elems,
                 float omd, const float* __restrict__ omd_dev) {
  if (omd_dev != nullptr) omd = __ldg(omd_dev);
  constexpr int ES = BF16 ? 2 : 4;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const sdt_chunk ch = chunks[c];
    const int64_t total = numels[ch.tensor];
    int64_t len = total - ch.offset;
    if (len > chunk_elems) len = chunk_elems;
    char* s = reinterpret_cast<char*>(shadow_ptrs[ch.tensor]) + ch.offset * ES;
    const char* p = reinterpret_cast<const char*>(param_ptrs[ch.tensor]) + ch.offset * ES;
    // chunk offsets are multiples of chunk_elems (a multiple of 8), so 16-byte alignment holds
    // whenever the tensor base is 16-byte aligned (checked on the host).
    ema_range<BF16>(s, p, len, omd, threadIdx.x, blockDim.x);
  }
}

// =============================================================================================
// f1: AdamW (+ optional EMA) over the flat arena
// =============================================================================================
struct AdamHyper { float lr, beta1, beta2, eps, wd, bc1, bc2; };

__global__ void __launch_bounds__(kThreads)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, AdamHyper h, const float* __restrict__ hyper_dev, float grad_scale,
                  float* __restrict__ ema_shadow, float ema_omd, const float* __restrict__ ema_omd_dev) {
  if (hyper_dev != nullptr) {
    h.lr = __ldg(hyper_dev + 0); h.beta1 = __ldg(hyper_dev + 1); h.beta2 = __ldg(hyper_dev + 2); h.eps = __ldg(hyper_dev + 3);
    h.wd = __ldg(hyper_dev + 4); h.bc1 = __ldg(hyper_dev + 5); h.bc2 = __ldg(hyper_dev + 6);
  }
  if (ema_omd_dev != nullptr) ema_omd = __ldg(ema_omd_dev);
  const float decay_mul = 1.0f - h.lr * h.wd;
  const float step_size = h.lr / h.bc1;
  const float inv_bc2_sqrt = 1.0f / sqrtf(h.bc2);
  const float omb1 = 1.0f - h.beta1, omb2 = 1.0f - h.beta2;
  const int64_t nvec = n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    pp *= decay_mul;
    mm = mm + omb1 * (gg - mm);                 // exp_avg.lerp_(grad, 1-beta1)
    vv = vv * h.beta2 + omb2 * gg * gg;         // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    const float denom = sqrtf(vv) * inv_bc2_sqrt + h.eps;
    pp = pp - step_size * (mm / denom);
  };
  for (int64_t i = tid; i < nvec; i += nth) {
    uint4 pv = ld_rw(reinterpret_cast<uint4*>(p) + i), gv = ld_stream(reinterpret_cast<const uint4*>(g) + i);
    uint4 mv = ld_rw(reinterpret_cast<uint4*>(m) + i), vv = ld_rw(reinterpret_cast<uint4*>(v) + i);
    float* pf = reinterpret_cast<float*>(&pv); const float* gf = reinterpret_cast<const float*>(&gv);
    float* mf = reinterpret_cast<float*>(&mv); float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) one(pf[j], gf[j], mf[j], vf[j]);
    st_stream(reinterpret_cast<uint4*>(p) + i, pv);
    st_stream(reinterpret_cast<uint4*>(m) + i, mv);
    st_stream(reinterpret_cast<uint4*>(v) + i, vv);
    if (ema_shadow != nullptr) {
      uint4 sv = ld_rw(reinterpret_cast<uint4*>(ema_shadow) + i);
      st_stream(reinterpret_cast<uint4*>(ema_shadow) + i, ema_vec_f32(sv, pv, ema_omd));
    }
  }
  for (int64_t k = (nvec << 2) + tid; k < n; k += nth) {
    float pp = p[k], mm = m[k], vv = v[k];
    one(pp, g[k], mm, vv);
    p[k] = pp; m[k] = mm; v[k] = vv;
    if (ema_shadow != nullptr) ema_shadow[k] = ema_f32(ema_shadow[k], pp, ema_omd);
  }
}

// =============================================================================================
// LoRA operand packing
// =============================================================================================
__device__ __forceinline__ uint16_t f32_to_act_bits(float f, bool f16) {
  return f16 ? __half_as_ushort(__float2half_rn(f)) : (uint16_t)f32_to_bf16_bits(f);
}
__global__ void __launch_bounds__(kThreads)
lora_pack_kernel(const sdt_pack_site* __restrict__ sites, int f16) {
  const sdt_pack_site s = sites[blockIdx.y];
  const int K = s.K, N = s.N, r = s.r, rt = s.r_true;
  const int64_t nA = (int64_t)r * K, nB = (int64_t)N * r;
  uint16_t* A_p = reinterpret_cast<uint16_t*>(s.A_p);
  uint16_t* At_p = reinterpret_cast<uint16_t*>(s.At_p);
  uint16_t* B_p = reinterpret_cast<uint16_t*>(s.B_p);
  uint16_t* Bt_p = reinterpret_cast<uint16_t*>(s.Bt_p);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nA + nB; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nA) {                         // (j,k) in A_p order
      const int j = (int)(e / K), k = (int)(e - (int64_t)j * K);
      const uint16_t b = j < rt ? f32_to_act_bits(__ldg(s.A + (int64_t)j * K + k), f16 != 0) : (uint16_t)0;
      A_p[e] = b;
      At_p[(int64_t)k * r + j] = b;
    } else {                              // (n,j) in B_p order
      const int64_t f = e - nA;
      const int n = (int)(f / r), j = (int)(f - (int64_t)n * r);
      const uint16_t b = j < rt ? f32_to_act_bits(__ldg(s.B + (int64_t)n * rt + j), f16 != 0) : (uint16_t)0;
      B_p[f] = b;
      Bt_p[(int64_t)j * N + n] = b;
    }
  }
}

// =============================================================================================
// f2: GEGLU (the activation that follows ff.net.0.proj): out = h * gelu(gate), proj = [h | gate]
// =============================================================================================
// gelu_erf / gelu_erf_grad: sdt_common.cuh (shared with the GEGLU epilogue of the fused projection)

// proj [M, 2I] bf16 -> out [M, I] bf16.  A thread handles 16-byte vectors (8 elements) of h and of gate, TWO of them per
// iteration with all four loads issued before the math (the erf GELU is ~20 instructions per element: with one vector per
// iteration the loads of the next iteration waited behind it), and 32-bit index arithmetic (a 64-bit division per vector cost as
// much as two elements of GELU).
__device__ __forceinline__ uint4 geglu_fwd_vec(const uint4& hv, const uint4& gv) {
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(&hv);
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(&gv);
  uint4 ov;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float h0 = bf16_bits_to_f32(hw[j] & 0xffffu), h1 = bf16_bits_to_f32(hw[j] >> 16);
    const float g0 = bf16_bits_to_f32(gw[j] & 0xffffu), g1 = bf16_bits_to_f32(gw[j] >> 16);
    ow[j] = pack_bf16x2(h0 * gelu_erf(g0), h1 * gelu_erf(g1));
  }
  return ov;
}
__global__ void __launch_bounds__(kThreads)
geglu_fwd_bf16_kernel(const uint16_t* __restrict__ proj, uint16_t* __restrict__ out, uint32_t total, uint32_t I8) {
  pdl_wait();                 // PDL (sdt_common.cuh)
  pdl_launch_dependents();
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint4* pv = reinterpret_cast<const uint4*>(proj);
  uint4* ov = reinterpret_cast<uint4*>(out);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const uint32_t i2 = i + stride;
    const bool two = i2 < total;
    const uint32_t r0 = i / I8, r1 = two ? i2 / I8 : 0u;
    const uint4* b0 = pv + (size_t)r0 * (2 * I8) + (i - r0 * I8);
    const uint4* b1 = pv + (size_t)r1 * (2 * I8) + (two ? i2 - r1 * I8 : 0u);
    const uint4 h0 = ld_stream(b0), g0 = ld_stream(b0 + I8);
    uint4 h1 = make_uint4(0u, 0u, 0u, 0u), g1 = h1;
    if (two) { h1 = ld_stream(b1); g1 = ld_stream(b1 + I8); }
    st_stream(ov + i, geglu_fwd_vec(h0, g0));
    if (two) st_stream(ov + i2, geglu_fwd_vec(h1, g1));
  }
}

// dproj[:, :I] = dout * gelu(gate) ; dproj[:, I:] = dout * h * gelu'(gate); same structure (six loads in flight per thread)
__device__ __forceinline__ void geglu_bwd_vec(const uint4& hv, const uint4& gv, const uint4& dv, uint4& dhv, uint4& dgv) {
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(&hv);
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(&gv);
  const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dv);
  uint32_t* dhw = reinterpret_cast<uint32_t*>(&dhv);
  uint32_t* dgw = reinterpret_cast<uint32_t*>(&dgv);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float h0 = bf16_bits_to_f32(hw[j] & 0xffffu), h1 = bf16_bits_to_f32(hw[j] >> 16);
    const float g0 = bf16_bits_to_f32(gw[j] & 0xffffu), g1 = bf16_bits_to_f32(gw[j] >> 16);
    const float d0 = bf16_bits_to_f32(dw[j] & 0xffffu), d1 = bf16_bits_to_f32(dw[j] >> 16);
    dhw[j] = pack_bf16x2(d0 * gelu_erf(g0), d1 * gelu_erf(g1));
    dgw[j] = pack_bf16x2(d0 * h0 * gelu_erf_grad(g0), d1 * h1 * gelu_erf_grad(g1));
  }
}
__global__ void __launch_bounds__(kThreads)
geglu_bwd_bf16_kernel(const uint16_t* __restrict__ proj, const uint16_t* __restrict__ dout, uint16_t* __restrict__ dproj,
                      uint32_t total, uint32_t I8) {
  pdl_wait();                 // PDL (sdt_common.cuh)
  pdl_launch_dependents();
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint4* pv = reinterpret_cast<const uint4*>(proj);
  const uint4* dov = reinterpret_cast<const uint4*>(dout);
  uint4* dpv = reinterpret_cast<uint4*>(dproj);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const uint32_t i2 = i + stride;
    const bool two = i2 < total;
    const uint32_t r0 = i / I8, r1 = two ? i2 / I8 : 0u;
    const size_t o0 = (size_t)r0 * (2 * I8) + (i - r0 * I8), o1 = (size_t)r1 * (2 * I8) + (two ? i2 - r1 * I8 : 0u);
    const uint4 h0 = ld_stream(pv + o0), g0 = ld_stream(pv + o0 + I8), d0 = ld_stream(dov + i);
    uint4 h1 = make_uint4(0u, 0u, 0u, 0u), g1 = h1, d1 = h1;
    if (two) { h1 = ld_stream(pv + o1); g1 = ld_stream(pv + o1 + I8); d1 = ld_stream(dov + i2); }
    uint4 dh, dg;
    geglu_bwd_vec(h0, g0, d0, dh, dg);
    st_stream(dpv + o0, dh);
    st_stream(dpv + o0 + I8, dg);
    if (two) {
      geglu_bwd_vec(h1, g1, d1, dh, dg);
      st_stream(dpv + o1, dh);
      st_stream(dpv + o1 + I8, dg);
    }
  }
}

// f32 variant (parity path), scalar
__global__ void __launch_bounds__(kThreads)
geglu_f32_kernel(const float* __restrict__ proj, const float* __restrict__ dout, float* __restrict__ out_or_dproj,
                 int64_t M, int64_t I, int backward) {
  const int64_t total = M * I;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / I, c = i - row * I;
    const float h = proj[row * 2 * I + c], g = proj[row * 2 * I + I + c];
    if (!backward) {
      out_or_dproj[i] = h * gelu_erf(g);
    } else {
      const float d = dout[i];
      out_or_dproj[row * 2 * I + c] = d * gelu_erf(g);
      out_or_dproj[row * 2 * I + I + c] = d * h * gelu_erf_grad(g);
    }
  }
}

}  // namespace sdt

// =============================================================================================
// C ABI
// =============================================================================================
using namespace sdt;

extern "C" int sdt_noise_target(const void* x0, const void* eps, const int64_t* t, const float* alphas_cumprod,
                                int num_train_timesteps, void* noisy, void* target, int mode,
                                int64_t B, int64_t chw, int dtype, int32_t* oob_flag, void* stream) {
  SDT_REQUIRE(x0 && eps && t && alphas_cumprod && noisy, SDT_ERR_ARG, "sdt_noise_target: null pointer");
  SDT_REQUIRE(B > 0 && chw > 0 && num_train_timesteps > 0, SDT_ERR_ARG, "sdt_noise_target: bad sizes B=%lld chw=%lld T=%d",
              (long long)B, (long long)chw, num_train_timesteps);
  SDT_REQUIRE(mode == SDT_TARGET_EPSILON || mode == SDT_TARGET_SAMPLE || mode == SDT_TARGET_V, SDT_ERR_ARG,
              "Unknown prediction type (mode=%d)", mode);   // modules/model.py:313-314
  SDT_REQUIRE((mode == SDT_TARGET_V) == (target != nullptr), SDT_ERR_ARG,
              "sdt_noise_target: target must be given for mode V and NULL otherwise");
  SDT_REQUIRE(dtype == SDT_F32 || dtype == SDT_BF16, SDT_ERR_UNSUPPORTED, "sdt_noise_target: unsupported dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = dtype == SDT_BF16, wv = mode == SDT_TARGET_V;
  const int vec = bf ? 8 : 4;
  const bool vec_ok = (chw % vec == 0) && aligned16(x0) && aligned16(eps) && aligned16(noisy) && (!wv || aligned16(target));
  const int64_t n = B * chw;
  if (vec_ok) {
    const int64_t nvec = n / vec;
    const int grid = grid_for(nvec, kThreads, 8);
    if (bf) {
      if (wv) noise_target_bf16_kernel<true><<<grid, kThreads, 0, st>>>((const uint16_t*)x0, (const uint16_t*)eps, t, alphas_cumprod, num_train_timesteps, (uint16_t*)noisy, (uint16_t*)target, nvec, chw / vec, oob_flag);
      else    noise_target_bf16_kernel<false><<<grid, kThreads, 0, st>>>((const uint16_t*)x0, (const uint16_t*)eps, t, alphas_cumprod, num_train_timesteps, (uint16_t*)noisy, nullptr, nvec, chw / vec, oob_flag);
    } else {
      if (wv) noise_target_f32_kernel<true><<<grid, kThreads, 0, st>>>((const float*)x0, (const float*)eps, t, alphas_cumprod, num_train_timesteps, (float*)noisy, (float*)target, nvec, chw / vec, oob_flag);
      else    noise_target_f32_kernel<false><<<grid, kThreads, 0, st>>>((const float*)x0, (const float*)eps, t, alphas_cumprod, num_train_timesteps, (float*)noisy, nullptr, nvec, chw / vec, oob_flag);
    }
  } else {
    const int grid = grid_for(n, kThreads, 8);
    if (bf) {
      if (wv) noise_target_scalar_kernel<true, true><<<grid, kThreads, 0, st>>>(x0, eps, t, alphas_cumprod, num_train_timesteps, noisy, target, n, chw, oob_flag);
      else    noise_target_scalar_kernel<true, false><<<grid, kThreads, 0, st>>>(x0, eps, t, alphas_cumprod, num_train_timesteps, noisy, nullptr, n, chw, oob_flag);
    } else {
      if (wv) noise_target_scalar_kernel<false, true><<<grid, kThreads, 0, st>>>(x0, eps, t, alphas_cumprod, num_train_timesteps, noisy, target, n, chw, oob_flag);
      else    noise_target_scalar_kernel<false, false><<<grid, kThreads, 0, st>>>(x0, eps, t, alphas_cumprod, num_train_timesteps, noisy, nullptr, n, chw, oob_flag);
    }
  }
  SDT_LAUNCH_OK("noise_target");
  return SDT_OK;
}

extern "C" size_t sdt_mse_loss_workspace_bytes(void) { return sizeof(LossWorkspace); }

template <typename PT, typename TT>
static int launch_mse(const void* pred, const void* target, float* loss_out, void* dpred, float* loss_elem,
                      int32_t* nan_flag, int64_t B, int64_t chw, int64_t split, float w_prior, float grad_scale,
                      void* workspace, cudaStream_t st) {
  const int64_t n = B * chw;
  const bool two = split < B;
  const double n0 = (double)(two ? split : B) * (double)chw, n1 = two ? (double)(B - split) * (double)chw : 1.0;
  const float inv_n0 = (float)(1.0 / n0), inv_n1 = two ? (float)(1.0 / n1) : 0.f;
  const float coef0 = inv_n0, coef1 = two ? (float)((double)w_prior / n1) : 0.f;
  const float w = two ? w_prior : 0.f;
  const bool vec_ok = (chw % 8 == 0) && aligned16(pred) && aligned16(target) && (!dpred || aligned16(dpred)) &&
                      (!loss_elem || aligned16(loss_elem));
  LossWorkspace* ws = reinterpret_cast<LossWorkspace*>(workspace);
  if (vec_ok) {
    const int64_t units = n / 8;
    int grid = grid_for(units, kThreads, 4);
    if (grid > kLossMaxBlocks) grid = kLossMaxBlocks;
    mse_loss_kernel<PT, TT, 8><<<grid, kThreads, 0, st>>>(pred, target, loss_out, dpred, loss_elem, nan_flag, units,
                                                          (two ? split : B) * (chw / 8), coef0, coef1, inv_n0, inv_n1, w,
                                                          grad_scale, ws);
  } else {
    int grid = grid_for(n, kThreads, 4);
    if (grid > kLossMaxBlocks) grid = kLossMaxBlocks;
    mse_loss_kernel<PT, TT, 1><<<grid, kThreads, 0, st>>>(pred, target, loss_out, dpred, loss_elem, nan_flag, n,
                                                          (two ? split : B) * chw, coef0, coef1, inv_n0, inv_n1, w,
                                                          grad_scale, ws);
  }
  SDT_LAUNCH_OK("mse_loss");
  return SDT_OK;
}

extern "C" int sdt_mse_loss(const void* pred, int pred_dtype, const void* target, int target_dtype,
                            float* loss_out, void* dpred, float* loss_elem, int32_t* nan_flag,
                            int64_t B, int64_t chw, int64_t split, float w_prior, float grad_scale,
                            void* workspace, void* stream) {
  SDT_REQUIRE(pred && target && loss_out && workspace, SDT_ERR_ARG, "sdt_mse_loss: null pointer");
  SDT_REQUIRE(B > 0 && chw > 0, SDT_ERR_ARG, "sdt_mse_loss: bad sizes");
  SDT_REQUIRE(split > 0 && split <= B, SDT_ERR_ARG, "sdt_mse_loss: split=%lld outside (0,B=%lld]", (long long)split, (long long)B);
  cudaStream_t st = (cudaStream_t)stream;
#define SDT_MSE(PT, TT) return launch_mse<PT, TT>(pred, target, loss_out, dpred, loss_elem, nan_flag, B, chw, split, w_prior, grad_scale, workspace, st)
  if (pred_dtype == SDT_F32 && target_dtype == SDT_F32) SDT_MSE(float, float);
  if (pred_dtype == SDT_BF16 && target_dtype == SDT_F32) SDT_MSE(__nv_bfloat16, float);
  if (pred_dtype == SDT_BF16 && target_dtype == SDT_BF16) SDT_MSE(__nv_bfloat16, __nv_bfloat16);
  if (pred_dtype == SDT_F32 && target_dtype == SDT_BF16) SDT_MSE(float, __nv_bfloat16);
  if (pred_dtype == SDT_F16 && target_dtype == SDT_F32) SDT_MSE(__half, float);
  if (pred_dtype == SDT_F16 && target_dtype == SDT_F16) SDT_MSE(__half, __half);
#undef SDT_MSE
  set_error("sdt_mse_loss: unsupported dtypes pred=%d target=%d", pred_dtype, target_dtype);
  return SDT_ERR_UNSUPPORTED;
}

extern "C" int sdt_ema_update_flat(void* shadow, const void* param, int64_t n, float one_minus_decay,
                                   const float* one_minus_decay_dev, int dtype, void* stream) {
  SDT_REQUIRE(shadow && param, SDT_ERR_ARG, "sdt_ema_update_flat: null pointer");
  SDT_REQUIRE(n >= 0, SDT_ERR_ARG, "sdt_ema_update_flat: n < 0");
  SDT_REQUIRE(dtype == SDT_F32 || dtype == SDT_BF16, SDT_ERR_UNSUPPORTED, "sdt_ema_update_flat: unsupported dtype %d", dtype);
  SDT_REQUIRE(aligned16(shadow) && aligned16(param), SDT_ERR_ARG, "sdt_ema_update_flat: pointers must be 16-byte aligned");
  if (n == 0) return SDT_OK;
  const int epv = dtype == SDT_BF16 ? 8 : 4;
  const int grid = grid_for(n / epv / 4 + 1, kThreads, 8);
  if (dtype == SDT_BF16) ema_flat_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(shadow, param, n, one_minus_decay, one_minus_decay_dev);
  else                   ema_flat_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(shadow, param, n, one_minus_decay, one_minus_decay_dev);
  SDT_LAUNCH_OK("ema_flat");
  return SDT_OK;
}

extern "C" int sdt_ema_update_multi(void* const* shadow_ptrs, const void* const* param_ptrs, const int64_t* numels,
                                    const sdt_chunk* chunks, int n_chunks, int chunk_elems, float one_minus_decay,
                                    const float* one_minus_decay_dev, int dtype, void* stream) {
  SDT_REQUIRE(shadow_ptrs && param_ptrs && numels && chunks, SDT_ERR_ARG, "sdt_ema_update_multi: null pointer");
  SDT_REQUIRE(chunk_elems > 0 && chunk_elems % 8 == 0, SDT_ERR_ARG, "sdt_ema_update_multi: chunk_elems must be a positive multiple of 8");
  SDT_REQUIRE(dtype == SDT_F32 || dtype == SDT_BF16, SDT_ERR_UNSUPPORTED, "sdt_ema_update_multi: unsupported dtype %d", dtype);
  if (n_chunks <= 0) return SDT_OK;
  const int grid = grid_for(n_chunks, 1, 8);
  if (dtype == SDT_BF16) ema_multi_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(shadow_ptrs, param_ptrs, numels, chunks, n_chunks, chunk_elems, one_minus_decay, one_minus_decay_dev);
  else                   ema_multi_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(shadow_ptrs, param_ptrs, numels, chunks, n_chunks, chunk_elems, one_minus_decay, one_minus_decay_dev);
  SDT_LAUNCH_OK("ema_multi");
  return SDT_OK;
}

extern "C" int sdt_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                              const float* hyper_dev, float grad_scale, float* ema_shadow, float ema_one_minus_decay,
                              const float* ema_one_minus_decay_dev, void* stream) {
  SDT_REQUIRE(p && g && m && v && hyper, SDT_ERR_ARG, "sdt_adamw_flat: null pointer");
  SDT_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!ema_shadow || aligned16(ema_shadow)),
              SDT_ERR_ARG, "sdt_adamw_flat: pointers must be 16-byte aligned");
  if (n <= 0) return SDT_OK;
  AdamHyper h{hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], hyper[6]};
  const int grid = grid_for(n / 4 + 1, kThreads, 8);
  adamw_flat_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, h, hyper_dev, grad_scale, ema_shadow,
                                                                 ema_one_minus_decay, ema_one_minus_decay_dev);
  SDT_LAUNCH_OK("adamw_flat");
  return SDT_OK;
}

extern "C" int sdt_lora_pack(const sdt_pack_site* sites, int n_sites, int64_t max_site_elems, int dtype, void* stream) {
  SDT_REQUIRE(sites, SDT_ERR_ARG, "sdt_lora_pack: null pointer");
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_pack: operands are bf16 or fp16 (got dtype %d)", dtype);
  if (n_sites <= 0) return SDT_OK;
  SDT_REQUIRE(n_sites <= 65535, SDT_ERR_ARG, "sdt_lora_pack: too many sites");
  int gx = (int)((max_site_elems + kThreads * 4 - 1) / (kThreads * 4));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  lora_pack_kernel<<<dim3(gx, n_sites), kThreads, 0, (cudaStream_t)stream>>>(sites, dtype == SDT_F16 ? 1 : 0);
  SDT_LAUNCH_OK("lora_pack");
  return SDT_OK;
}

extern "C" int sdt_geglu(const void* proj, const void* dout, void* out_or_dproj, int64_t M, int64_t I, int backward,
                         int dtype, void* stream) {
  SDT_REQUIRE(proj && out_or_dproj && (!backward || dout), SDT_ERR_ARG, "sdt_geglu: null pointer");
  SDT_REQUIRE(M > 0 && I > 0, SDT_ERR_ARG, "sdt_geglu: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SDT_BF16) {
    SDT_REQUIRE(I % 8 == 0 && aligned16(proj) && aligned16(out_or_dproj) && aligned16(dout), SDT_ERR_UNSUPPORTED,
                "sdt_geglu(bf16): inner width must be a multiple of 8 and pointers 16-byte aligned");
    SDT_REQUIRE(M * (I / 8) < (1ll << 31), SDT_ERR_UNSUPPORTED, "sdt_geglu(bf16): more than 2^31 vectors (M = %lld, I = %lld)", (long long)M,
                (long long)I);
    const uint32_t total = (uint32_t)(M * (I / 8));
    const int grid = grid_for((int64_t)(total + 1) / 2, kThreads, 8);          // two vectors per thread and iteration
    if (!backward)
      SDT_CUDA_OK(launch_kernel(geglu_fwd_bf16_kernel, dim3(grid), dim3(kThreads), 0, st, true, (const uint16_t*)proj,
                                (uint16_t*)out_or_dproj, total, (uint32_t)(I / 8)));
    else
      SDT_CUDA_OK(launch_kernel(geglu_bwd_bf16_kernel, dim3(grid), dim3(kThreads), 0, st, true, (const uint16_t*)proj,
                                (const uint16_t*)dout, (uint16_t*)out_or_dproj, total, (uint32_t)(I / 8)));
  } else if (dtype == SDT_F32) {
    const int grid = grid_for(M * I, kThreads, 8);
    geglu_f32_kernel<<<grid, kThreads, 0, st>>>((const float*)proj, (const float*)dout, (float*)out_or_dproj, M, I, backward);
  } else {
    set_error("sdt_geglu: unsupported dtype %d", dtype);
    return SDT_ERR_UNSUPPORTED;
  }
  SDT_LAUNCH_OK("geglu");
  return SDT_OK;
}


// ---- f2: residual add with a folded per-channel bias, channels-last / token-major bf16 [rows, C] --------------------------
// out = a + b + bias[c].  diffusers ResnetBlock2D ends with  hidden = conv2(hidden); out = shortcut(x) + hidden , and torch adds
// every convolution's bias as a separate broadcast kernel (aten::add_ with a [1,C,1,1] operand: non-vectorised, 22 us on a
// 21 MB tensor).  The convolutions run without bias and the (frozen) biases are added here, in the pass that the residual add
// needs anyway.  HBM-bound: 3 tensors of rows*C bf16.
namespace sdt {
__global__ void __launch_bounds__(kThreads)
residual_bias_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const float4* __restrict__ bias,
                         uint4* __restrict__ out, int64_t n_vec, int C8) {
  pdl_wait();                 // PDL (sdt_common.cuh)
  pdl_launch_dependents();
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  constexpr int U = 4;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_vec; i += U * stride) {
    uint4 av[U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n_vec) { av[u] = ld_stream(a + i + u * stride); bv[u] = ld_stream(b + i + u * stride); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iu = i + u * stride;
      if (iu >= n_vec) break;
      const int c = (int)(iu % C8);
      const float4 b0 = __ldg(bias + 2 * c), b1 = __ldg(bias + 2 * c + 1);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const uint32_t* aw = reinterpret_cast<const uint32_t*>(&av[u]);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv[u]);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(aw[j] << 16) + __uint_as_float(bw[j] << 16) + bb[2 * j];
        const float hi = __uint_as_float(aw[j] & 0xffff0000u) + __uint_as_float(bw[j] & 0xffff0000u) + bb[2 * j + 1];
        o[j] = pack_bf16x2(lo, hi);
      }
      st_stream(out + iu, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}
}  // namespace sdt

extern "C" int sdt_residual_bias_add(const void* a, const void* b, const float* bias, void* out, int64_t rows, int C, void* stream) {
  SDT_REQUIRE(a && b && bias && out, SDT_ERR_ARG, "sdt_residual_bias_add: null pointer");
  SDT_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, SDT_ERR_UNSUPPORTED, "sdt_residual_bias_add: needs C %% 8 == 0 (got C=%d)", C);
  SDT_REQUIRE(aligned16(a) && aligned16(b) && aligned16(bias) && aligned16(out), SDT_ERR_ARG,
              "sdt_residual_bias_add: pointers must be 16-byte aligned");
  const int64_t n_vec = rows * (C / 8);
  const int grid = grid_for(n_vec, kThreads, 8);
  SDT_CUDA_OK(sdt::launch_kernel(sdt::residual_bias_add_kernel, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, true,
                                 (const uint4*)a, (const uint4*)b, (const float4*)bias, (uint4*)out, n_vec, C / 8));
  SDT_LAUNCH_OK("residual_bias_add");
  return SDT_OK;
}
