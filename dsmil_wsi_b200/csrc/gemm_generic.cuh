// Generic fp32 FFMA GEMM building blocks (any N, K, M; no alignment assumptions).
// Used by the generic forward for shapes the tcgen05 kernel does not take (D % 64 != 0, e.g. the
// classic-MIL 166/230 features of train_mil.py:127-141), by the V projection (dsmil.py:35-39) and
// by the backward GEMMs.  Deterministic: split reductions go through partial buffers summed in a
// fixed order, never float atomics.
#pragma once
#include "common.cuh"

namespace dsmil {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_MASK_POS = 3 };

constexpr int LBM = 64, LBN = 128, LBK = 16;

// Y[N,M] (+)= act(X[N,K] * op(W) + b);  op(W)[k,m] = WT ? W[k*M+m] : W[m*K+k].
// ACT_MASK_POS: y = aux[n,m] > 0 ? y : 0  (ReLU backward through a saved activation).
template <int ACT, bool WT>
__global__ void __launch_bounds__(256)
k_linear(const float* __restrict__ X, int64_t N, int K, const float* __restrict__ W,
         const float* __restrict__ b, int M, float* __restrict__ Y, const float* __restrict__ aux,
         int accumulate) {
  __shared__ __align__(16) float As[LBK][LBM + 4];
  __shared__ __align__(16) float Bs[LBK][LBN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * LBM;
  const int m0 = blockIdx.y * LBN;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int ar = tid >> 2, akq = (tid & 3) * 4;  // A loader: row, k-quad
  const bool a_vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (int k0 = 0; k0 < K; k0 += LBK) {
    {  // ---- A tile
      const int64_t n = n0 + ar;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < N) {
        const float* src = X + n * K + k0 + akq;
        if (a_vec && k0 + akq + 3 < K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k0 + akq + i < K) v[i] = __ldg(src + i);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[akq + i][ar] = v[i];
    }
    if (!WT) {  // ---- B tile from W[m,k]
      const int m = tid >> 1, kq = (tid & 1) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = 0.f;
        if (m0 + m < M && k0 + kq + i < K) v = __ldg(W + static_cast<int64_t>(m0 + m) * K + k0 + kq + i);
        Bs[kq + i][m] = v;
      }
    } else {  // ---- B tile from W[k,m]
      const int kk = tid >> 4, mq = (tid & 15) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = 0.f;
        if (k0 + kk < K && m0 + mq + i < M) v = __ldg(W + static_cast<int64_t>(k0 + kk) * M + m0 + mq + i);
        Bs[kk][mq + i] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (m >= M) continue;
      float y = acc[i][j] + (b ? __ldg(b + m) : 0.f);
      if (ACT == ACT_RELU) y = fmaxf(y, 0.f);
      if (ACT == ACT_TANH) y = tanhf(y);
      if (ACT == ACT_MASK_POS) y = (__ldg(aux + n * M + m) > 0.f) ? y : 0.f;
      float* dst = Y + n * M + m;
      *dst = accumulate ? (*dst + y) : y;
    }
  }
}

template <int ACT, bool WT>
inline int launch_linear(const float* X, int64_t N, int K, const float* W, const float* b, int M, float* Y,
                         const float* aux, int accumulate, cudaStream_t st) {
  if (N <= 0) return 0;
  dim3 grid(ceil_div(N, LBM), ceil_div(M, LBN));
  k_linear<ACT, WT><<<grid, 256, 0, st>>>(X, N, K, W, b, M, Y, aux, accumulate);
  DSMIL_LAUNCH_OK("k_linear");
  return 0;
}

// ---- out[M1,M2] = sum_n P[n,M1] * R[n,M2]  (reduction over rows; split over grid.z into partials)
constexpr int TBM = 64, TBK = 16;
__global__ void __launch_bounds__(256)
k_gemm_tn(const float* __restrict__ P, int M1, const float* __restrict__ R, int M2, int64_t N,
          int64_t rows_per_split, float* __restrict__ part) {
  __shared__ __align__(16) float Ps[TBK][TBM + 4];
  __shared__ __align__(16) float Rs[TBK][TBM + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int a0 = blockIdx.y * TBM, b0 = blockIdx.x * TBM;
  const int64_t nb = static_cast<int64_t>(blockIdx.z) * rows_per_split;
  const int64_t ne = min(N, nb + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ln = tid >> 4, lq = (tid & 15) * 4;
  for (int64_t n0 = nb; n0 < ne; n0 += TBK) {
    const int64_t n = n0 + ln;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float pv = 0.f, rv = 0.f;
      if (n < ne) {
        if (a0 + lq + i < M1) pv = __ldg(P + n * M1 + a0 + lq + i);
        if (b0 + lq + i < M2) rv = __ldg(R + n * M2 + b0 + lq + i);
      }
      Ps[ln][lq + i] = pv;
      Rs[ln][lq + i] = rv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Ps[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Rs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + static_cast<int64_t>(blockIdx.z) * M1 * M2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= M1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < M2) out[static_cast<int64_t>(a) * M2 + b] = acc[i][j];
    }
  }
}

// part[z][M] = sum over the z-th row chunk of P[n, m]
__global__ void __launch_bounds__(256)
k_colsum(const float* __restrict__ P, int M, int64_t N, int64_t rows_per_split, float* __restrict__ part) {
  const int64_t nb = static_cast<int64_t>(blockIdx.x) * rows_per_split;
  const int64_t ne = min(N, nb + rows_per_split);
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    float s = 0.f;
    for (int64_t n = nb; n < ne; ++n) s += __ldg(P + n * M + m);
    part[static_cast<int64_t>(blockIdx.x) * M + m] = s;
  }
}

// out[i] = sum_z part[z][i]  (fixed order)
__global__ void __launch_bounds__(256)
k_sum_partials(const float* __restrict__ part, int S, int64_t L, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= L) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += part[static_cast<int64_t>(z) * L + i];
  out[i] = s;
}

inline int tn_splits(int M1, int M2, int64_t N) {
  const int tiles = ceil_div(M1, TBM) * ceil_div(M2, TBM);
  int s = 296 / (tiles > 0 ? tiles : 1);
  const int maxs = ceil_div(N, 128);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}
inline size_t tn_partial_floats(int M1, int M2, int64_t N) {
  return static_cast<size_t>(tn_splits(M1, M2, N)) * M1 * M2;
}
// out[M1,M2] = P^T R, via partials in `part` (>= tn_partial_floats floats).
inline int launch_gemm_tn(const float* P, int M1, const float* R, int M2, int64_t N, float* part, float* out,
                          cudaStream_t st) {
  if (N <= 0) {
    DSMIL_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * M1 * M2, st));
    return 0;
  }
  const int S = tn_splits(M1, M2, N);
  int64_t rps = (N + S - 1) / S;
  rps = (rps + TBK - 1) / TBK * TBK;
  dim3 grid(ceil_div(M2, TBM), ceil_div(M1, TBM), S);
  k_gemm_tn<<<grid, 256, 0, st>>>(P, M1, R, M2, N, rps, part);
  DSMIL_LAUNCH_OK("k_gemm_tn");
  const int64_t L = static_cast<int64_t>(M1) * M2;
  k_sum_partials<<<ceil_div(L, 256), 256, 0, st>>>(part, S, L, out);
  DSMIL_LAUNCH_OK("k_sum_partials");
  return 0;
}
inline int colsum_splits(int64_t N) {
  int s = ceil_div(N, 64);
  return s > 296 ? 296 : (s < 1 ? 1 : s);
}
inline int launch_colsum(const float* P, int M, int64_t N, float* part, float* out, cudaStream_t st) {
  if (N <= 0) {
    DSMIL_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * M, st));
    return 0;
  }
  const int S = colsum_splits(N);
  const int64_t rps = (N + S - 1) / S;
  k_colsum<<<S, 256, 0, st>>>(P, M, N, rps, part);
  DSMIL_LAUNCH_OK("k_colsum");
  k_sum_partials<<<ceil_div(M, 256), 256, 0, st>>>(part, S, M, out);
  DSMIL_LAUNCH_OK("k_sum_partials");
  return 0;
}

}  // namespace dsmil
