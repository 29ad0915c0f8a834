// Element math shared by all kernels: exact-erf GELU and its derivative, bf16 helpers, vector row
// load/store, and the counter-based (Philox4x32-7) dropout mask that every kernel can regenerate
// from (seed, layer, row, column) so no mask is ever stored.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace p2t {

// ------------------------------------------------------------------------------------------------
// GELU (torch.nn.GELU() default = exact erf form; reference models/modeling_esm2llama_instruct.py:54)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * z * z);
  return fmaf(z, pdf, cdf);
}
// value and derivative together (they share the erf)
__device__ __forceinline__ void gelu_erf_both(float z, float& val, float& der) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * z * z);
  val = z * cdf;
  der = fmaf(z, pdf, cdf);
}
// Fast erf-form GELU pair for the GEMM epilogues.  Same function as above (z*Phi(z), Phi via erf — NOT the
// tanh approximation), with erfc evaluated by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, below fp32 erff's
// own error for O(1) arguments and 4 orders of magnitude below the bf16/fp16 resolution of the stored result).
// exp(-z^2/2) is shared between erfc and the Gaussian density, so value + derivative cost one ex2 + one rcp.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_erf_both_fast(float z, float& val, float& der) {
  const float ax = fabsf(z) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  const float e = ex2_approx(z * z * -0.72134752044448170f);  // exp(-z^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * e;  // 0.5*erfc(|z|/sqrt2)
  const float cdf = (z >= 0.f) ? 1.0f - half_erfc : half_erfc;
  val = z * cdf;
  der = fmaf(z * 0.39894228040143268f, e, cdf);
}
// The same pair already multiplied by the dropout multiplier s = 1/(1-p): val = s z Phi(z), der = s (Phi(z) + z phi(z)).
// s rides on the shared exponential (one FMUL) and on the "1 -" of the cdf; the 0.5 of erfc and the 1/sqrt2 of its
// argument are folded into immediates.  Two instructions fewer per element than gelu_erf_both_fast followed by two
// scalings, and no constant beyond s lives in a register.
__device__ __forceinline__ void gelu_erf_both_scaled(float z, float s, float& val, float& der) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(z), 1.0f));
  const float es = ex2_approx((z * -0.72134752044448170f) * z) * s;  // s exp(-z^2/2)
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float h = poly * t * es;                 // s/2 erfc(|z|/sqrt2)
  const float cdf = (z >= 0.f) ? s - h : h;      // s Phi(z)
  val = z * cdf;
  der = fmaf(z * 0.39894228040143268f, es, cdf);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// dropout: keep multiplier for element (row, col) of layer `layer`
//   one Philox4x32-7 block = 8 x 16-bit lanes = columns [8g, 8g+8) of a row
//   counter = (row, g, layer, 0), key = seed; keep iff u16 >= threshold (threshold = round(p*65536))
// ------------------------------------------------------------------------------------------------
struct DropoutParams {
  unsigned long long seed;
  const unsigned long long* seed_dev;  // optional: the seed is read from device memory and ADDED to `seed`
                                       // (CUDA-graph replays bump it on the device; kernel parameters are frozen)
  float scale;         // 1/(1-p)
  uint32_t threshold;  // 0 -> dropout disabled
  uint32_t layer;      // 1 = after fc1 GELU, 2 = after fc2 GELU
};

// Philox4x32 with R rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3": 7 rounds is the
// shortest variant that passes BigCrush; 10 is the library default).  The dropout masks use 7.
constexpr int kDropoutPhiloxRounds = 7;
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// keep multipliers for the 8 columns [8g, 8g+8) of `row`
__device__ __forceinline__ void dropout_keep8(const DropoutParams& d, uint32_t row, uint32_t g, float (&keep)[8]) {
  if (d.threshold == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) keep[i] = 1.0f;
    return;
  }
  const uint4 r = philox4x32<kDropoutPhiloxRounds>(make_uint4(row, g, d.layer, 0u),
                                                  make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep[2 * i] = ((w[i] & 0xFFFFu) >= d.threshold) ? d.scale : 0.0f;
    keep[2 * i + 1] = ((w[i] >> 16) >= d.threshold) ? d.scale : 0.0f;
  }
}

// same, from an already generated Philox block
__device__ __forceinline__ void dropout_keep8_from(const DropoutParams& d, const uint4& r, float (&keep)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep[2 * i] = ((w[i] & 0xFFFFu) >= d.threshold) ? d.scale : 0.0f;
    keep[2 * i + 1] = ((w[i] >> 16) >= d.threshold) ? d.scale : 0.0f;
  }
}

// same, from plain (threshold, scale) values held in registers; threshold 0 -> dropout disabled
__device__ __forceinline__ void dropout_keep8_raw(uint32_t threshold, float scale, const uint4& r, float (&keep)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep[2 * i] = (threshold == 0 || (w[i] & 0xFFFFu) >= threshold) ? scale : 0.0f;
    keep[2 * i + 1] = (threshold == 0 || (w[i] >> 16) >= threshold) ? scale : 0.0f;
  }
}

// keep multipliers for 32 consecutive columns starting at `col` (col % 8 == 0)
struct DropoutRow {
  float k[32];
  __device__ __forceinline__ DropoutRow(const DropoutParams& d, int row, int col) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float k8[8];
      dropout_keep8(d, static_cast<uint32_t>(row), static_cast<uint32_t>((col >> 3) + j), k8);
#pragma unroll
      for (int i = 0; i < 8; ++i) k[8 * j + i] = k8[i];
    }
  }
  __device__ __forceinline__ float keep(int i) const { return k[i]; }
};

// ------------------------------------------------------------------------------------------------
// row fragments: 32 consecutive elements owned by one thread
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
template <bool F16>
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  if constexpr (F16) return unpack_f16x2(u); else return unpack_bf16x2(u);
}
__device__ __forceinline__ float f16_round(float x) { return __half2float(__float2half_rn(x)); }

// fp16 flavour of the 32-element row fragment helpers (activations private to our own kernels are
// kept in fp16: same bytes as bf16, 8x finer mantissa; GEMM operands stay bf16)
__device__ __forceinline__ void store_row_f16(__half* dst, const float (&f)[32], int ncols) {
  if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      u.x = pack_f16x2(f[8 * q + 0], f[8 * q + 1]);
      u.y = pack_f16x2(f[8 * q + 2], f[8 * q + 3]);
      u.z = pack_f16x2(f[8 * q + 4], f[8 * q + 5]);
      u.w = pack_f16x2(f[8 * q + 6], f[8 * q + 7]);
      reinterpret_cast<uint4*>(dst)[q] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = __float2half_rn(f[i]);
  }
}
__device__ __forceinline__ void load_row_f16(const __half* src, float (&f)[32], int ncols) {
  if (ncols == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + q);
      float2 a = unpack_f16x2(u.x), b = unpack_f16x2(u.y), c = unpack_f16x2(u.z), d = unpack_f16x2(u.w);
      f[8 * q + 0] = a.x; f[8 * q + 1] = a.y; f[8 * q + 2] = b.x; f[8 * q + 3] = b.y;
      f[8 * q + 4] = c.x; f[8 * q + 5] = c.y; f[8 * q + 6] = d.x; f[8 * q + 7] = d.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? __half2float(src[i]) : 0.f;
  }
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&f)[32], int ncols) {
  if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      u.x = pack_bf16x2(f[8 * q + 0], f[8 * q + 1]);
      u.y = pack_bf16x2(f[8 * q + 2], f[8 * q + 3]);
      u.z = pack_bf16x2(f[8 * q + 4], f[8 * q + 5]);
      u.w = pack_bf16x2(f[8 * q + 6], f[8 * q + 7]);
      reinterpret_cast<uint4*>(dst)[q] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = __float2bfloat16_rn(f[i]);
  }
}
__device__ __forceinline__ void store_row_f32(float* dst, const float (&f)[32], int ncols) {
  if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      reinterpret_cast<float4*>(dst)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = f[i];
  }
}
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* src, float (&f)[32], int ncols) {
  if (ncols == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + q);
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      f[8 * q + 0] = a.x; f[8 * q + 1] = a.y; f[8 * q + 2] = b.x; f[8 * q + 3] = b.y;
      f[8 * q + 4] = c.x; f[8 * q + 5] = c.y; f[8 * q + 6] = d.x; f[8 * q + 7] = d.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? __bfloat162float(src[i]) : 0.f;
  }
}
// same values for every thread of the warp (bias): the loads coalesce into broadcasts
__device__ __forceinline__ void load_row_bf16_bcast(const __nv_bfloat16* src, float (&f)[32], int ncols) {
  if (src == nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = 0.f;
    return;
  }
  load_row_bf16(src, f, ncols);
}

// ------------------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace p2t
