// Access-pattern microbenchmark: how fast can 148 persistent CTAs stream a [N,512] fp32 matrix when each
// CTA owns 128-row tiles and reads them (a) in 64-column chunks (256 B pieces at 2 KB stride, the tensor-core
// kernel's pattern) or (b) as whole 2 KB rows.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <int PATTERN, int DEPTH, int SYNC = 0>
__global__ void __launch_bounds__(256, 1) k_stream(const float* X, long long N, float* out) {
  const int tid = threadIdx.x, seg = tid & 15, r0 = tid >> 4;
  const long long ntiles = N / 128;
  float acc = 0.f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float4* base = reinterpret_cast<const float4*>(X + tile * 128 * 512);
    if (PATTERN == 0) {  // chunked: half-chunk h: 4 float4 per thread, DEPTH half-chunks in flight
      float4 buf[DEPTH][4];
      auto load = [&](int h, float4* d) {
        const int kc = h >> 1, part = h & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = ldg_stream(base + (long long)(r0 + 16 * (i + 4 * part)) * 128 + kc * 16 + seg);
      };
#pragma unroll
      for (int d = 0; d < DEPTH - 1; ++d) load(d, buf[d]);
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        if (h + DEPTH - 1 < 16) load(h + DEPTH - 1, buf[(h + DEPTH - 1) % DEPTH]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += buf[h % DEPTH][i].x + buf[h % DEPTH][i].w;
        if (SYNC && (h & 1)) __syncthreads();   // couple all 8 warps once per 32 KB chunk, like a shared stage
      }
    } else {  // row-contiguous: warp w reads rows w*16..+15, 4 float4 per lane per row (2 KB per row)
      const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        const float4* row = base + (long long)(warp * 16 + r) * 128;
        float4 a = ldg_stream(row + lane), b = ldg_stream(row + 32 + lane), c = ldg_stream(row + 64 + lane), d = ldg_stream(row + 96 + lane);
        acc += a.x + b.x + c.x + d.w;
      }
    }
  }
  if (acc == 123456.f) out[0] = acc;
}
template <int P, int DPT, int SY = 0> void run(const char* name, const float* X, long long N, float* out, int grid) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k_stream<P, DPT, SY><<<grid, 256>>>(X, N, out);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) k_stream<P, DPT, SY><<<grid, 256>>>(X, N, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s grid=%d  %.1f us  %.2f TB/s\n", name, grid, ms / reps * 1e3, N * 2048.0 / (ms / reps * 1e-3) / 1e12);
}
int main() {
  const long long N = 400000;   // 819 MB > L2
  float *X, *out; cudaMalloc(&X, N * 2048); cudaMalloc(&out, 4); cudaMemset(X, 0, N * 2048);
  for (int grid : {148, 296}) {
    run<0, 2>("chunked depth2 (16KB/SM)", X, N, out, grid);
    run<0, 3>("chunked depth3", X, N, out, grid);
    run<0, 4>("chunked depth4 (48KB/SM)", X, N, out, grid);
    run<0, 8>("chunked depth8", X, N, out, grid);
    run<1, 1>("row-contiguous", X, N, out, grid);
    run<0, 2, 1>("chunked depth2 + sync/chunk", X, N, out, grid);
    run<0, 4, 1>("chunked depth4 + sync/chunk", X, N, out, grid);
    run<0, 8, 1>("chunked depth8 + sync/chunk", X, N, out, grid);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
