// Producer/consumer chain microbenchmark for the phase-1 kernel's load side.
//
// The product kernel's "skeleton" (loads + handshakes only) streams X at 2.6 TB/s while the same access pattern
// without handshakes streams at 6.7 TB/s (tools/membench.cu).  This tool rebuilds only the chain
//
//   16 loader warps: LDG (DEPTH chunks of register prefetch) -> [bf16 hi/lo split] -> STS into a ring stage
//        -> mbarrier WRITTEN (one arrive per warp)
//   fence thread:    WRITTEN -> fence.proxy.async -> FULL           (optional hop, as in the product)
//   consumer thread: FULL -> [busy for MMA_CYCLES] -> EMPTY          (stands in for MMA issue + tcgen05.commit)
//
// with the ring depth, the polling back-off, the prefetch depth and the stand-in MMA time as template knobs, and
// prints the streaming rate of each combination, so the pacing hop can be found without touching the product.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/chainbench tools/chainbench.cu && tools/chainbench
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kLoaderWarps = 16, kThreads = 32 * (kLoaderWarps + 2);
constexpr int kChunkBytes = 32 * 1024;      // [128 rows x 64 k] bf16 hi tile + lo tile
constexpr int kTileBytes = 16 * 1024;
constexpr int D = 512, kChunks = D / 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int SLEEP>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && SLEEP > 0) __nanosleep(SLEEP);
    if (spins > (1u << 26)) __trap();       // a protocol bug must not hang the box
  }
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__host__ __device__ inline uint32_t swz_off(int row, int k) {
  return static_cast<uint32_t>(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1));
}

// STAGES: ring depth; SLEEP: ns of back-off in every poll loop; FENCE: 1 = extra fence-thread hop;
// DEPTH: chunks of loads in flight per loader thread (2 or 3); CONVERT: 1 = bf16 hi/lo split + STS, 0 = loads only
// (values are folded into a checksum); MMA: cycles the consumer stays busy per chunk
template <int STAGES, int SLEEP, int FENCE, int DEPTH, int CONVERT, int MMA>
__global__ void __launch_bounds__(kThreads, 1) k_chain(const float* __restrict__ X, long long ntiles, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto WRITTEN = [&](int s) { return smem_u32(&bars[s]); };
  auto FULL = [&](int s) { return smem_u32(&bars[STAGES + s]); };
  auto EMPTY = [&](int s) { return smem_u32(&bars[2 * STAGES + s]); };
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(WRITTEN(s), kLoaderWarps);
      mbar_init(FULL(s), FENCE ? 1 : kLoaderWarps);
      mbar_init(EMPTY(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long my_chunks = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) my_chunks += kChunks;

  if (warp < kLoaderWarps) {
    const int seg = tid & 15, r0 = tid >> 4;
    const uint32_t off0 = swz_off(r0, seg * 4), ring = smem_u32(smem);
    float4 buf[DEPTH][4];
    // flat chunk counter c -> (tile, kc); loads run DEPTH-1 chunks ahead of the convert, across tile boundaries
    auto load = [&](long long c, float4* d) {
      const long long tile = blockIdx.x + (c / kChunks) * gridDim.x;
      const int kc = static_cast<int>(c % kChunks);
      const float4* base = reinterpret_cast<const float4*>(X + tile * 128 * D) + static_cast<long long>(r0) * (D / 4) + kc * 16 + seg;
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = ldg_stream(base + static_cast<long long>(i) * 32 * (D / 4));
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d)
      if (d < my_chunks) load(d, buf[d]);
    uint32_t stage = 0, phase = 0;
    float acc = 0.f;
    for (long long c0 = 0; c0 < my_chunks; c0 += DEPTH) {
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) {           // unrolled so that buf[] indices are compile-time constants
        const long long c = c0 + j;
        if (c >= my_chunks) break;
        if (c + DEPTH - 1 < my_chunks) load(c + DEPTH - 1, buf[(j + DEPTH - 1) % DEPTH]);
        mbar_wait<SLEEP>(EMPTY(stage), phase ^ 1);
        const uint32_t hi_tile = ring + stage * kChunkBytes + off0, lo_tile = hi_tile + kTileBytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = buf[j][i];
          if (CONVERT) {
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
            const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(u01 << 16), x.y - __uint_as_float(u01 & 0xffff0000u));
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(u23 << 16), x.w - __uint_as_float(u23 & 0xffff0000u));
            sts64(hi_tile + i * 4096, u01, u23);
            sts64(lo_tile + i * 4096, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          } else {
            acc += x.x + x.y + x.z + x.w;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(FENCE ? WRITTEN(stage) : FULL(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (acc == 123456.f) out[1] = acc;
  } else if (warp == kLoaderWarps) {
    if (FENCE && lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long c = 0; c < my_chunks; ++c) {
        mbar_wait<SLEEP>(WRITTEN(stage), phase);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(FULL(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (lane == 0) {
    uint32_t stage = 0, phase = 0;
    unsigned sink = 0;
    for (long long c = 0; c < my_chunks; ++c) {
      mbar_wait<SLEEP>(FULL(stage), phase);
      if (MMA > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < MMA) sink += 1;
      }
      if (CONVERT) sink += smem[stage * kChunkBytes + (c & 1023)];   // touch the stage like a consumer would
      mbar_arrive(EMPTY(stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (sink == 0xdeadbeefu) out[0] = 1.f;
  }
}

template <int STAGES, int SLEEP, int FENCE, int DEPTH, int CONVERT, int MMA>
void run(const float* X, long long N, float* out) {
  auto k = k_chain<STAGES, SLEEP, FENCE, DEPTH, CONVERT, MMA>;
  const size_t smem = static_cast<size_t>(STAGES) * kChunkBytes;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
    printf("stages=%d: %zu B of shared memory not available\n", STAGES, smem);
    return;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const long long ntiles = N / 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k<<<sms, kThreads, smem>>>(X, ntiles, out);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) k<<<sms, kThreads, smem>>>(X, ntiles, out);
  cudaEventRecord(e1);
  const cudaError_t err = cudaEventSynchronize(e1);
  if (err != cudaSuccess) {
    printf("stages=%d sleep=%d fence=%d depth=%d convert=%d mma=%d: %s\n", STAGES, SLEEP, FENCE, DEPTH, CONVERT, MMA,
           cudaGetErrorString(err));
    return;
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms / reps * 1e3;
  printf("stages=%d sleep=%3dns fence_hop=%d prefetch=%d convert=%d mma=%4d cyc : %7.1f us  %5.2f TB/s  %6.0f cycles/chunk @1.9GHz\n",
         STAGES, SLEEP, FENCE, DEPTH, CONVERT, MMA, us, N * 2048.0 / (us * 1e-6) / 1e12,
         us * 1e-6 * 1.9e9 / (static_cast<double>(ntiles) * kChunks / sms));
}

int main() {
  const long long N = 400000 / 128 * 128;   // 819 MB > L2
  float *X, *out;
  if (cudaMalloc(&X, N * 2048) != cudaSuccess || cudaMalloc(&out, 16) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMemset(X, 0, N * 2048);
  printf("# chainbench: %lld rows x 512 fp32, one persistent CTA per SM, 16 loader warps\n", N);
  // loads + handshakes only (the product's "skeleton"), product settings: 4 stages, 128 ns back-off, fence hop
  run<4, 128, 1, 2, 0, 0>(X, N, out);
  run<4, 0, 1, 2, 0, 0>(X, N, out);
  run<4, 32, 1, 2, 0, 0>(X, N, out);
  run<4, 128, 0, 2, 0, 0>(X, N, out);
  run<4, 0, 0, 2, 0, 0>(X, N, out);
  run<4, 0, 0, 3, 0, 0>(X, N, out);
  run<6, 0, 0, 2, 0, 0>(X, N, out);
  run<2, 0, 0, 2, 0, 0>(X, N, out);
  // with the bf16 split + STS
  run<4, 128, 1, 2, 1, 0>(X, N, out);
  run<4, 0, 1, 2, 1, 0>(X, N, out);
  run<4, 0, 0, 2, 1, 0>(X, N, out);
  run<4, 0, 0, 3, 1, 0>(X, N, out);
  run<6, 0, 1, 2, 1, 0>(X, N, out);
  // with a consumer that is busy for one chunk's MMA time (12 x 64 cycles) before releasing the stage
  run<4, 128, 1, 2, 1, 768>(X, N, out);
  run<4, 0, 1, 2, 1, 768>(X, N, out);
  run<4, 0, 1, 3, 1, 768>(X, N, out);
  run<6, 0, 1, 2, 1, 768>(X, N, out);
  run<6, 0, 1, 3, 1, 768>(X, N, out);
  cudaFree(X);
  cudaFree(out);
  return 0;
}
