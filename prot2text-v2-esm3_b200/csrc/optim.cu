// Gradient-norm clipping + AdamW for the adapter parameters, as two kernels over a table of tensors
// (reference: torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW(eps=1e-6, betas=(0.9, 0.999)) + zero_grad,
//  scripts/train_contrast.py:455-465, 621-626).
//
// HBM-bound: per parameter element the update reads grad (2 B), m, v (4 + 4 B) and the fp32 master weight (4 B) and
// writes m, v, master and the bf16 parameter: 32 B; the norm pass reads 2 B more.  Nothing synchronises with the host:
// the step count, the learning rate and the clip coefficient live in device memory, so the optimizer step can sit
// inside the same CUDA graph as the contrastive step.
#include "common.h"
#include "mathfn.cuh"
#include "rows.h"

namespace p2t {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = kOptThreads * 8 * 4;  // elements per CTA: 4 x (8 bf16 = 16 B) per thread

__device__ __forceinline__ int find_tensor(const AdamTable& t, int block, int& local_block) {
  int i = 0;
#pragma unroll 1
  while (i + 1 < t.count && block >= t.block_start[i + 1]) ++i;
  local_block = block - t.block_start[i];
  return i;
}

template <int THREADS>
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < THREADS / 32 ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// partial[block] = sum of squares of this block's chunk of gradient elements (fp32 accumulate, fixed order).
// A tensor with an fp32 source (the mean over ranks left by the gradient all-reduce) is rounded to bf16 HERE — the
// one rounding of the sharded step's gradients — into the bf16 gradient buffer the update kernel (and param.grad)
// reads; the norm is that of the rounded values, i.e. what clip_grad_norm_ sees on the bf16 .grad.
__global__ void __launch_bounds__(kOptThreads)
grad_sqnorm_partial_kernel(AdamTable t, float* __restrict__ partial) {
  __shared__ float sh[kOptThreads / 32];
  int lb;
  const int ti = find_tensor(t, blockIdx.x, lb);
  __nv_bfloat16* g = static_cast<__nv_bfloat16*>(t.grad[ti]);
  const float* src = t.grad_f32[ti];
  const long long n = t.numel[ti];
  const long long base = (long long)lb * kOptChunk;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long e = base + ((long long)k * kOptThreads + threadIdx.x) * 8;
    if (e + 8 <= n) {
      uint4 u;
      if (src != nullptr) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(src + e)), hi = __ldg(reinterpret_cast<const float4*>(src + e + 4));
        u = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
        *reinterpret_cast<uint4*>(g + e) = u;
      } else {
        u = __ldg(reinterpret_cast<const uint4*>(g + e));
      }
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    } else {
      for (long long j = e; j < n; ++j) {
        if (src != nullptr) g[j] = __float2bfloat16_rn(src[j]);
        const float x = __bfloat162float(g[j]);
        s += x * x;
      }
    }
  }
  s = block_reduce_sum<kOptThreads>(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// One CTA: total norm from the partials (fixed order), clip coefficient, step counter, bias corrections.
//   scal[0] = ||g||            (what clip_grad_norm_ returns)
//   scal[1] = clip coefficient min(1, max_norm / (||g|| + 1e-6))   (1 when max_norm <= 0 or inf)
//   scal[2] = lr / (1 - beta1^t),  scal[3] = 1 / sqrt(1 - beta2^t)
__global__ void __launch_bounds__(kOptThreads)
adamw_prepare_kernel(const float* __restrict__ partial, int n_partial, float max_norm, float beta1, float beta2,
                     const float* __restrict__ lr_dev, long long* __restrict__ step_dev, float* __restrict__ scal) {
  __shared__ float sh[kOptThreads / 32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += kOptThreads) s += partial[i];
  s = block_reduce_sum<kOptThreads>(s, sh);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s);
    float coef = 1.f;
    if (max_norm > 0.f && !isinf(max_norm)) coef = fminf(1.f, max_norm / (norm + 1e-6f));
    const long long step = *step_dev + 1;
    *step_dev = step;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    scal[0] = norm;
    scal[1] = coef;
    scal[2] = (float)((double)*lr_dev / bc1);
    scal[3] = (float)(1.0 / sqrt(bc2));
  }
}

// AdamW (decoupled weight decay), torch.optim.AdamW semantics:
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// g is the clipped gradient.  With a master copy the arithmetic runs on the fp32 weight and the bf16 parameter is
// its rounding; without one the bf16 parameter itself is the state (the reference's configuration).
__global__ void __launch_bounds__(kOptThreads)
adamw_update_kernel(AdamTable t, const float* __restrict__ scal, const float* __restrict__ lr_dev, float beta1, float beta2,
                    float eps, float weight_decay, int zero_grad) {
  int lb;
  const int ti = find_tensor(t, blockIdx.x, lb);
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(t.param[ti]);
  __nv_bfloat16* g = static_cast<__nv_bfloat16*>(t.grad[ti]);
  float* m = t.exp_avg[ti];
  float* v = t.exp_avg_sq[ti];
  float* w = t.master[ti];
  const long long n = t.numel[ti];
  const float clip = scal[1], step_size = scal[2], inv_sqrt_bc2 = scal[3];
  const float decay = 1.f - *lr_dev * weight_decay;
  const long long base = (long long)lb * kOptChunk;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const long long e0 = base + ((long long)k * kOptThreads + threadIdx.x) * 8;
    if (e0 >= n) continue;
    const int cnt = (int)min((long long)8, n - e0);
    float gv[8], mv[8], vv[8], wv[8];
    if (cnt == 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(g + e0);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      gv[0] = a.x; gv[1] = a.y; gv[2] = b.x; gv[3] = b.y; gv[4] = c.x; gv[5] = c.y; gv[6] = d.x; gv[7] = d.y;
      const float4 m0 = *reinterpret_cast<const float4*>(m + e0), m1 = *reinterpret_cast<const float4*>(m + e0 + 4);
      const float4 v0 = *reinterpret_cast<const float4*>(v + e0), v1 = *reinterpret_cast<const float4*>(v + e0 + 4);
      mv[0] = m0.x; mv[1] = m0.y; mv[2] = m0.z; mv[3] = m0.w; mv[4] = m1.x; mv[5] = m1.y; mv[6] = m1.z; mv[7] = m1.w;
      vv[0] = v0.x; vv[1] = v0.y; vv[2] = v0.z; vv[3] = v0.w; vv[4] = v1.x; vv[5] = v1.y; vv[6] = v1.z; vv[7] = v1.w;
      if (w) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + e0), w1 = *reinterpret_cast<const float4*>(w + e0 + 4);
        wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
      } else {
        const uint4 pu = *reinterpret_cast<const uint4*>(p + e0);
        const float2 a2 = unpack_bf16x2(pu.x), b2 = unpack_bf16x2(pu.y), c2 = unpack_bf16x2(pu.z), d2 = unpack_bf16x2(pu.w);
        wv[0] = a2.x; wv[1] = a2.y; wv[2] = b2.x; wv[3] = b2.y; wv[4] = c2.x; wv[5] = c2.y; wv[6] = d2.x; wv[7] = d2.y;
      }
    } else {
      for (int j = 0; j < 8; ++j) {
        const bool ok = j < cnt;
        gv[j] = ok ? __bfloat162float(g[e0 + j]) : 0.f;
        mv[j] = ok ? m[e0 + j] : 0.f;
        vv[j] = ok ? v[e0 + j] : 0.f;
        wv[j] = ok ? (w ? w[e0 + j] : __bfloat162float(p[e0 + j])) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = gv[j] * clip;
      mv[j] = fmaf(beta1, mv[j], (1.f - beta1) * gj);
      vv[j] = fmaf(beta2, vv[j], (1.f - beta2) * gj * gj);
      const float denom = fmaf(sqrtf(vv[j]), inv_sqrt_bc2, eps);
      wv[j] = fmaf(-step_size, mv[j] / denom, wv[j] * decay);
    }
    if (cnt == 8) {
      *reinterpret_cast<float4*>(m + e0) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(m + e0 + 4) = make_float4(mv[4], mv[5], mv[6], mv[7]);
      *reinterpret_cast<float4*>(v + e0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      *reinterpret_cast<float4*>(v + e0 + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
      if (w) {
        *reinterpret_cast<float4*>(w + e0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        *reinterpret_cast<float4*>(w + e0 + 4) = make_float4(wv[4], wv[5], wv[6], wv[7]);
      }
      *reinterpret_cast<uint4*>(p + e0) = make_uint4(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]),
                                                     pack_bf16x2(wv[4], wv[5]), pack_bf16x2(wv[6], wv[7]));
      if (zero_grad) *reinterpret_cast<uint4*>(g + e0) = make_uint4(0, 0, 0, 0);
    } else {
      for (int j = 0; j < cnt; ++j) {
        m[e0 + j] = mv[j];
        v[e0 + j] = vv[j];
        if (w) w[e0 + j] = wv[j];
        p[e0 + j] = __float2bfloat16_rn(wv[j]);
        if (zero_grad) g[e0 + j] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

}  // namespace

int adamw_blocks(long long numel) { return (int)((numel + kOptChunk - 1) / kOptChunk); }

int adamw_step(AdamTable t, float* partial_ws, float* scal, const float* lr_dev, long long* step_dev, float beta1,
               float beta2, float eps, float weight_decay, float max_norm, int zero_grad, cudaStream_t st) {
  if (t.count < 1 || t.count > kAdamMaxTensors) return set_error(-1, "p2t_adamw_step: 1..%d tensors", kAdamMaxTensors);
  int blocks = 0;
  for (int i = 0; i < t.count; ++i) {
    if (!t.param[i] || !t.grad[i] || !t.exp_avg[i] || !t.exp_avg_sq[i] || t.numel[i] <= 0)
      return set_error(-1, "p2t_adamw_step: tensor %d has a null pointer or no elements", i);
    if ((reinterpret_cast<uintptr_t>(t.param[i]) | reinterpret_cast<uintptr_t>(t.grad[i])) & 15 ||
        (reinterpret_cast<uintptr_t>(t.exp_avg[i]) | reinterpret_cast<uintptr_t>(t.exp_avg_sq[i]) |
         reinterpret_cast<uintptr_t>(t.master[i]) | reinterpret_cast<uintptr_t>(t.grad_f32[i])) & 15)
      return set_error(-1, "p2t_adamw_step: tensor %d is not 16-byte aligned", i);
    t.block_start[i] = blocks;
    blocks += adamw_blocks(t.numel[i]);
  }
  grad_sqnorm_partial_kernel<<<blocks, kOptThreads, 0, st>>>(t, partial_ws);
  if (int r = check_launch("grad_sqnorm_partial_kernel", st)) return r;
  adamw_prepare_kernel<<<1, kOptThreads, 0, st>>>(partial_ws, blocks, max_norm, beta1, beta2, lr_dev, step_dev, scal);
  if (int r = check_launch("adamw_prepare_kernel", st)) return r;
  adamw_update_kernel<<<blocks, kOptThreads, 0, st>>>(t, scal, lr_dev, beta1, beta2, eps, weight_decay, zero_grad);
  return check_launch("adamw_update_kernel", st);
}

}  // namespace p2t
