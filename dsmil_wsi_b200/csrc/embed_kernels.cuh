// Embedder kernels of the patch-embedding loop: the reference builds its backbone as a torchvision
// ResNet with norm_layer = nn.InstanceNorm2d (compute_feats.py:146-170; affine = False, track_running_stats = False,
// eps = 1e-5), i.e. after every convolution the framework runs  instance_norm -> (+ identity) -> relu  as two or three
// memory-bound passes over the activation tensor.  One kernel does all of it in ONE read and ONE write:
//
//   k_instnorm_plane   one CTA per (n, c) plane of H*W > 1024 elements: the plane is staged in shared memory (<= 16 K
//                      floats: 112x112 = 12 544 is the largest plane of a 224x224 ResNet), mean and biased variance by
//                      two passes over the staged copy (the two-pass form, not E[x^2] - mean^2), then
//                      y = (x - mean) * rsqrt(var + eps) (+ residual) (relu) written out with 16-byte stores
//   k_instnorm_warp    one WARP per plane of <= 1024 elements (56x56 is handled by the CTA kernel; 28x28, 14x14, 7x7
//                      here): the plane lives in registers, 8 planes per CTA
//
// HBM-bound: algorithmic bytes = 8 B per element (+ 4 with a residual).  NCHW fp32 contiguous; in place allowed.
#pragma once
#include "common.cuh"

namespace dsmil {

constexpr int kInPlaneMax = 16384;      // floats staged per plane (64 KB of shared memory)

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {   // fixed-order: warp shuffles, then 8 partials
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
}

template <bool VEC>
__global__ void __launch_bounds__(256)
k_instnorm_plane(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y, int HW, float eps,
                 int relu) {
  extern __shared__ __align__(16) float s_plane[];
  __shared__ float s_red[8];
  const size_t base = static_cast<size_t>(blockIdx.x) * HW;
  const float* xp = x + base;
  float acc = 0.f;
  if (VEC) {
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    float4* s4 = reinterpret_cast<float4*>(s_plane);
    for (int i = threadIdx.x; i < (HW >> 2); i += 256) {
      const float4 v = __ldg(x4 + i);
      s4[i] = v;
      acc += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float v = __ldg(xp + i);
      s_plane[i] = v;
      acc += v;
    }
  }
  const float mean = block_sum_256(acc, s_red) / static_cast<float>(HW);
  float sq = 0.f;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const float d = s_plane[i] - mean;
    sq = fmaf(d, d, sq);
  }
  const float var = block_sum_256(sq, s_red) / static_cast<float>(HW);       // biased, as F.instance_norm
  const float rstd = rsqrtf(var + eps);
  if (VEC) {
    const float4* s4 = reinterpret_cast<const float4*>(s_plane);
    const float4* r4 = res ? reinterpret_cast<const float4*>(res + base) : nullptr;
    float4* y4 = reinterpret_cast<float4*>(y + base);
    for (int i = threadIdx.x; i < (HW >> 2); i += 256) {
      float4 v = s4[i];
      v.x = (v.x - mean) * rstd; v.y = (v.y - mean) * rstd; v.z = (v.z - mean) * rstd; v.w = (v.w - mean) * rstd;
      if (r4) { const float4 r = __ldg(r4 + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      y4[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += 256) {
      float v = (s_plane[i] - mean) * rstd;
      if (res) v += __ldg(res + base + i);
      if (relu) v = fmaxf(v, 0.f);
      y[base + i] = v;
    }
  }
}

// one warp per plane, HW <= 1024: element lane + 32 j lives in register j
__global__ void __launch_bounds__(256)
k_instnorm_warp(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y, long long planes, int HW,
                float eps, int relu) {
  const int lane = threadIdx.x & 31;
  const long long plane = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (plane >= planes) return;
  const size_t base = static_cast<size_t>(plane) * HW;
  float v[32];
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int i = lane + 32 * j;
    v[j] = i < HW ? __ldg(x + base + i) : 0.f;
    acc += v[j];
  }
  const float mean = warp_sum(acc) / static_cast<float>(HW);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = (lane + 32 * j < HW) ? v[j] - mean : 0.f;
    sq = fmaf(d, d, sq);
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(HW) + eps);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int i = lane + 32 * j;
    if (i < HW) {
      float o = (v[j] - mean) * rstd;
      if (res) o += __ldg(res + base + i);
      if (relu) o = fmaxf(o, 0.f);
      y[base + i] = o;
    }
  }
}

inline int launch_instnorm(const float* x, const float* res, float* y, long long planes, int HW, float eps, int relu,
                           cudaStream_t st) {
  if (HW <= 1024) {
    const long long grid = (planes + 7) / 8;
    k_instnorm_warp<<<static_cast<unsigned>(grid), 256, 0, st>>>(x, res, y, planes, HW, eps, relu);
    DSMIL_LAUNCH_OK("k_instnorm_warp");
    return 0;
  }
  const size_t smem = sizeof(float) * static_cast<size_t>(HW);
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                   (res == nullptr || (reinterpret_cast<uintptr_t>(res) & 15) == 0);
  if (vec) {
    if (smem > 48 * 1024)
      DSMIL_CUDA_OK(cudaFuncSetAttribute(k_instnorm_plane<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k_instnorm_plane<true><<<static_cast<unsigned>(planes), 256, smem, st>>>(x, res, y, HW, eps, relu);
  } else {
    if (smem > 48 * 1024)
      DSMIL_CUDA_OK(cudaFuncSetAttribute(k_instnorm_plane<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k_instnorm_plane<false><<<static_cast<unsigned>(planes), 256, smem, st>>>(x, res, y, HW, eps, relu);
  }
  DSMIL_LAUNCH_OK("k_instnorm_plane");
  return 0;
}

}  // namespace dsmil
