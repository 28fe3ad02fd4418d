// LoRA dropout on the rank path (modules/lora.py:12 -> loralib 0.1: result += (dropout(x) A^T) B^T * scaling).
//
// No shipped optim_target uses it (`dropout: 0.`), so this is the compatibility path, built from the fused kernels instead of
// a third kernel variant: the projection runs on the CONCATENATED contraction
//     X' = [x | xd] [M,2K],  W' = [W | 0],  A' = [0 | A]        xd = x * keep / (1 - p)
// so that  X' W'^T + s (X' A'^T) B^T = x W^T + s (xd A^T) B^T  and, in the backward, dX' = [dY W | G A]: the second half is the
// gradient w.r.t. xd.  The two kernels here build X' (one pass over x) and fold dX' back into dx (one pass), regenerating the
// keep mask from a counter-based generator (Philox4x32-10 keyed by a device-resident seed and a per-call salt), so no mask is
// stored and a captured CUDA graph draws fresh masks on every replay (the seed lives in device memory).
#include "sdt_common.cuh"

namespace sdt {

constexpr int kDropThreads = 256;

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep bits of the 8 elements of vector `vec` (16 random bits per element against a 16-bit threshold)
__device__ __forceinline__ void keep8(uint64_t seed, uint64_t salt, int64_t vec, uint32_t thresh16, bool (&keep)[8]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)vec, (uint32_t)((uint64_t)vec >> 32), (uint32_t)salt, (uint32_t)(salt >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    keep[2 * j] = (w[j] & 0xffffu) >= thresh16;
    keep[2 * j + 1] = (w[j] >> 16) >= thresh16;
  }
}

__device__ __forceinline__ void unpack_act8(const uint4& u, float (&f)[8], bool f16) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (f16) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      f[2 * j] = t.x; f[2 * j + 1] = t.y;
    } else {
      f[2 * j] = bf16_bits_to_f32(w[j] & 0xffffu); f[2 * j + 1] = bf16_bits_to_f32(w[j] >> 16);
    }
  }
}

// forward: xcat[m, 0:K] = x[m, :],  xcat[m, K:2K] = x[m, :] * keep / (1 - p)
__global__ void __launch_bounds__(kDropThreads)
lora_dropout_concat_kernel(const uint4* __restrict__ x, uint4* __restrict__ xcat, int64_t M, int K8, const int64_t* __restrict__ seed,
                           uint64_t salt, uint32_t thresh16, float inv_keep, int f16) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t sd = (uint64_t)*seed;
  const int64_t total = M * K8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / K8;
    const int c = (int)(i - row * K8);
    const uint4 xv = ld_stream(x + i);
    float f[8];
    bool keep[8];
    unpack_act8(xv, f, f16 != 0);
    keep8(sd, salt, i, thresh16, keep);
    uint4 dv;
    dv.x = pack_act2(keep[0] ? f[0] * inv_keep : 0.f, keep[1] ? f[1] * inv_keep : 0.f, f16 != 0);
    dv.y = pack_act2(keep[2] ? f[2] * inv_keep : 0.f, keep[3] ? f[3] * inv_keep : 0.f, f16 != 0);
    dv.z = pack_act2(keep[4] ? f[4] * inv_keep : 0.f, keep[5] ? f[5] * inv_keep : 0.f, f16 != 0);
    dv.w = pack_act2(keep[6] ? f[6] * inv_keep : 0.f, keep[7] ? f[7] * inv_keep : 0.f, f16 != 0);
    uint4* dst = xcat + row * 2 * K8 + c;
    st_stream(dst, xv);
    st_stream(dst + K8, dv);
  }
}

// backward: dx[m, :] = dxcat[m, 0:K] + dxcat[m, K:2K] * keep / (1 - p)      (same seed / salt as the forward)
__global__ void __launch_bounds__(kDropThreads)
lora_dropout_fold_kernel(const uint4* __restrict__ dxcat, uint4* __restrict__ dx, int64_t M, int K8, const int64_t* __restrict__ seed,
                         uint64_t salt, uint32_t thresh16, float inv_keep, int f16) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t sd = (uint64_t)*seed;
  const int64_t total = M * K8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / K8;
    const int c = (int)(i - row * K8);
    const uint4* src = dxcat + row * 2 * K8 + c;
    const uint4 av = ld_stream(src), bv = ld_stream(src + K8);
    float a[8], b[8];
    bool keep[8];
    unpack_act8(av, a, f16 != 0);
    unpack_act8(bv, b, f16 != 0);
    keep8(sd, salt, i, thresh16, keep);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] + (keep[j] ? b[j] * inv_keep : 0.f);
    st_stream(dx + i, make_uint4(pack_act2(o[0], o[1], f16 != 0), pack_act2(o[2], o[3], f16 != 0), pack_act2(o[4], o[5], f16 != 0),
                                 pack_act2(o[6], o[7], f16 != 0)));
  }
}

}  // namespace sdt

using namespace sdt;

extern "C" int sdt_lora_dropout(const void* in, void* out, int64_t M, int64_t K, float p, const int64_t* seed_dev, uint64_t salt,
                                int backward, int dtype, void* stream) {
  SDT_REQUIRE(in && out && seed_dev, SDT_ERR_ARG, "sdt_lora_dropout: null pointer");
  SDT_REQUIRE(M > 0 && K > 0 && K % 8 == 0, SDT_ERR_UNSUPPORTED, "sdt_lora_dropout: K must be a positive multiple of 8 (K=%lld)", (long long)K);
  SDT_REQUIRE(p > 0.f && p < 1.f, SDT_ERR_ARG, "sdt_lora_dropout: p=%g outside (0,1)", (double)p);
  SDT_REQUIRE(dtype == SDT_BF16 || dtype == SDT_F16, SDT_ERR_UNSUPPORTED, "sdt_lora_dropout: bf16 / fp16 only (got dtype %d)", dtype);
  SDT_REQUIRE(aligned16(in) && aligned16(out), SDT_ERR_ARG, "sdt_lora_dropout: pointers must be 16-byte aligned");
  uint32_t thresh = (uint32_t)(p * 65536.0f + 0.5f);
  if (thresh < 1) thresh = 1;
  if (thresh > 65535) thresh = 65535;
  const float inv_keep = 1.0f / (1.0f - (float)thresh / 65536.0f);      // the probability actually realised by the threshold
  const int K8 = (int)(K / 8);
  int64_t want = (M * K8 + kDropThreads - 1) / kDropThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  cudaStream_t st = (cudaStream_t)stream;
  if (!backward)
    SDT_CUDA_OK(launch_kernel(lora_dropout_concat_kernel, dim3(grid), dim3(kDropThreads), 0, st, true, (const uint4*)in, (uint4*)out, M,
                              K8, seed_dev, salt, thresh, inv_keep, dtype == SDT_F16 ? 1 : 0));
  else
    SDT_CUDA_OK(launch_kernel(lora_dropout_fold_kernel, dim3(grid), dim3(kDropThreads), 0, st, true, (const uint4*)in, (uint4*)out, M,
                              K8, seed_dev, salt, thresh, inv_keep, dtype == SDT_F16 ? 1 : 0));
  SDT_LAUNCH_OK("lora_dropout");
  return SDT_OK;
}
