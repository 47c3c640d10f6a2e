// sm_100a tensor-core path of phase 1 (dsmil.py:11 scores + :49 Q-MLP), D % 64 == 0.
//
//   k_prep_wimg     packs W1 / W2 into bf16 hi/lo "images" that are byte-for-byte the shared-memory
//                   operand tiles tcgen05.mma reads (K-major, SWIZZLE_128B), so they can be brought
//                   in with plain 1-D bulk copies (cp.async.bulk -> SASS UBLKCP), no tensor map.
//   k_qmlp_sm100    persistent, warp-specialised, one CTA per SM, 128-row tiles:
//        converter warps : coalesced float4 loads of X from HBM -> fp32 FFMA instance scores (+ arg-max
//                          key) -> split x = hi + lo (two bf16) -> swizzled st.shared operand tiles
//        TMA warp        : streams the W1 image chunk by chunk (L2-resident) into a smem ring
//        MMA warp        : one thread issues tcgen05.mma (M=128,N=128,K=16, bf16 in, fp32 accumulate in
//                          TMEM); 3 products per K-step: hi*Whi + lo*Whi + hi*Wlo  ("3xBF16", error at
//                          the fp32 noise floor -- SURVEY A.4, tests/test_oracle.py)
//        epilogue warps  : H1 = relu(acc + b1) -> bf16 hi/lo written BACK to TMEM (tcgen05.st) as the A
//                          operand of layer 2 (A-from-TMEM MMA); Q = tanh(acc2 + b2) -> global
//   TMEM columns (512): H1 accumulators 2 x 128 | Q accumulator 128 | A2 hi 64 | A2 lo 64
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dsmil {
namespace sm100 {

constexpr int kTileM = 128;            // rows per tile == UMMA M
constexpr int kChunkK = 64;            // k per smem operand chunk: 64 bf16 = 128 B = one swizzle row
constexpr int kTileBytes = kTileM * kChunkK * 2;       // 16 KiB: one [128 x 64] bf16 operand tile
constexpr int kChunkBytes = 2 * kTileBytes;            // hi tile + lo tile
constexpr int kAStages = 2;
constexpr int kWStages = 2;
constexpr int kConvWarps = 8;
constexpr int kThreads = 32 * (4 + 1 + 1 + kConvWarps);  // epilogue x4, MMA, TMA, converters
constexpr uint32_t kSpinLimit = 1u << 28;

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();  // a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
#define DSMIL_TMEM_LD32(taddr, v)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                   \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                   \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                  \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),       \
                 "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),   \
                 "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),\
                 "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),\
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                        \
               : "r"(taddr) : "memory")
#define DSMIL_TMEM_ST16(taddr, v)                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                             \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                 \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),  \
                 "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),          \
                 "r"(v[14]), "r"(v[15]) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version [46,48) = 1 (sm_100) | layout [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of element (row, k) inside one [rows x 64] bf16 SWIZZLE_128B tile
__host__ __device__ inline uint32_t swz_off(int row, int k) {
  return static_cast<uint32_t>(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1));
}

// ---- weight images -----------------------------------------------------------------------
// img layout: for each 64-wide k chunk: [hi tile 16 KiB][lo tile 16 KiB], rows = output features (128)
__global__ void __launch_bounds__(256)
k_prep_wimg(const float* __restrict__ W, int K, uint8_t* __restrict__ img) {
  const int total = 128 * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i % K;
    const float w = W[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    uint8_t* chunk = img + static_cast<size_t>(k / kChunkK) * kChunkBytes;
    const uint32_t off = swz_off(n, k % kChunkK);
    *reinterpret_cast<__nv_bfloat16*>(chunk + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(chunk + kTileBytes + off) = lo;
  }
}

struct QmlpArgs {
  const float* X;
  int64_t N;
  int D;
  int C;
  const float* Wi;
  const float* bi;
  const float* b1;
  const float* b2;
  const uint8_t* w1img;   // D/64 chunks
  const uint8_t* w2img;   // 2 chunks
  float* classes;         // [N,C]
  unsigned long long* keys;
  float* Q;               // [N,128]
  float* H1;              // [N,128] or NULL
};

// dynamic smem carve (bytes, from a 1024-aligned base)
constexpr int kOffW2 = 0;                                   // 2 chunks x 32 KiB
constexpr int kOffWRing = kOffW2 + 2 * kChunkBytes;         // kWStages x 32 KiB
constexpr int kOffARing = kOffWRing + kWStages * kChunkBytes;
constexpr int kOffWi = kOffARing + kAStages * kChunkBytes;  // up to 8 x 2048 floats?  (C*D floats)
constexpr int kSmemFixed = kOffWi;

template <int CT>
__global__ void __launch_bounds__(kThreads, 1)
k_qmlp_sm100(const QmlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_b1[kQ], s_b2[kQ];
  __shared__ unsigned long long s_best[kConvWarps][kMaxC];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D, C = a.C;
  const int nchunks = D / kChunkK;
  const int64_t ntiles = (a.N + kTileM - 1) / kTileM;

  // barrier indices
  enum { A_FULL = 0, A_EMPTY = A_FULL + kAStages, W_FULL = A_EMPTY + kAStages, W_EMPTY = W_FULL + kWStages,
         H1_FULL = W_EMPTY + kWStages, H1_EMPTY = H1_FULL + 2, A2_FULL = H1_EMPTY + 2, A2_EMPTY, Q_FULL, Q_EMPTY,
         W2_FULL, NBARS };
  static_assert(NBARS <= 32, "too many barriers");
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  float* sWi = reinterpret_cast<float*>(smem + kOffWi);
  const bool do_scores = a.classes != nullptr;   // bag form (scores given): Wi/bi may be NULL
  if (do_scores)
    for (int i = tid; i < C * D; i += kThreads) sWi[i] = a.Wi[i];
  if (tid < kQ) { s_b1[tid] = a.b1[tid]; s_b2[tid] = a.b2[tid]; }
  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(bar(A_FULL + s), kConvWarps); mbar_init(bar(A_EMPTY + s), 1); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(bar(W_FULL + s), 1); mbar_init(bar(W_EMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar(H1_FULL + b), 1); mbar_init(bar(H1_EMPTY + b), 128); }
    mbar_init(bar(A2_FULL), 128); mbar_init(bar(A2_EMPTY), 1);
    mbar_init(bar(Q_FULL), 1); mbar_init(bar(Q_EMPTY), 128);
    mbar_init(bar(W2_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {  // TMEM allocation (whole 512 columns; one CTA per SM by construction)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  const uint32_t tm_h1[2] = {tmem + 0, tmem + 128};
  const uint32_t tm_q = tmem + 256, tm_a2hi = tmem + 384, tm_a2lo = tmem + 448;

  if (warp >= 6) {
    // =============================== converter warps =========================================
    const int ct = tid - 6 * 32;                 // 0..255
    const int seg = ct & 15, r0 = ct >> 4;       // float4 index in the 64-float chunk row; base row
    unsigned long long best[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) best[k] = 0ull;
    uint32_t stage = 0, phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t row_base = tile * kTileM;
      float sc[8][CT];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k) sc[i][k] = 0.f;
      float4 cur[8], nxt[8];
      auto load_chunk = [&](int kc, float4* dst) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t n = row_base + r0 + 16 * i;
          dst[i] = (n < a.N) ? __ldg(reinterpret_cast<const float4*>(a.X + n * D + kc * kChunkK) + seg)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_chunk(0, cur);
      for (int kc = 0; kc < nchunks; ++kc) {
        if (kc + 1 < nchunks) load_chunk(kc + 1, nxt);
        mbar_wait(bar(A_EMPTY + stage), phase ^ 1);
        uint8_t* hi_tile = smem + kOffARing + stage * kChunkBytes;
        uint8_t* lo_tile = hi_tile + kTileBytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const float4 x = cur[i];
#pragma unroll
          for (int k = 0; k < CT; ++k)
            if (do_scores && k < C) {
              const float4 w = *reinterpret_cast<const float4*>(sWi + k * D + kc * kChunkK + seg * 4);
              float s = sc[i][k];
              s = fmaf(x.x, w.x, s); s = fmaf(x.y, w.y, s); s = fmaf(x.z, w.z, s); s = fmaf(x.w, w.w, s);
              sc[i][k] = s;
            }
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __low2float(h01), x.y - __high2float(h01));
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __low2float(h23), x.w - __high2float(h23));
          const uint32_t off = swz_off(r, seg * 4);
          uint2 hv, lv;
          hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
          lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(hi_tile + off) = hv;
          *reinterpret_cast<uint2*>(lo_tile + off) = lv;
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(A_FULL + stage));
        if (++stage == kAStages) { stage = 0; phase ^= 1; }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      }
      // instance scores of this tile: reduce over the 16 threads (seg) that share a row
#pragma unroll
      for (int i = 0; do_scores && i < 8; ++i) {
        const int64_t n = row_base + r0 + 16 * i;
#pragma unroll
        for (int k = 0; k < CT; ++k) {
          float v = sc[i][k];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if (seg == 0 && k < C && n < a.N) {
            v += a.bi[k];
            a.classes[n * C + k] = v;
            const unsigned long long key = pack_key(v, static_cast<uint32_t>(n));
            best[k] = key > best[k] ? key : best[k];
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      const unsigned long long b = warp_max_u64(best[k]);
      if (lane == 0) s_best[warp - 6][k] = b;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConvWarps * 32));   // converter-only named barrier
    if (ct < C && do_scores) {
      unsigned long long b = 0ull;
      for (int w = 0; w < kConvWarps; ++w) b = s_best[w][ct] > b ? s_best[w][ct] : b;
      if (b) atomicMax(a.keys + ct, b);
    }
  } else if (warp == 5) {
    // =============================== W1 image producer (bulk copies) ==========================
    if (lane == 0) {
      mbar_expect_tx(bar(W2_FULL), 2 * kChunkBytes);
      bulk_g2s(smem_u32(smem + kOffW2), a.w2img, 2 * kChunkBytes, bar(W2_FULL));
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc) {
          mbar_wait(bar(W_EMPTY + stage), phase ^ 1);
          mbar_expect_tx(bar(W_FULL + stage), kChunkBytes);
          bulk_g2s(smem_u32(smem + kOffWRing + stage * kChunkBytes), a.w1img + static_cast<size_t>(kc) * kChunkBytes,
                   kChunkBytes, bar(W_FULL + stage));
          if (++stage == kWStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // =============================== MMA issuer ================================================
    if (lane == 0) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      const uint32_t w2base = smem_u32(smem + kOffW2);
      auto issue_l2 = [&](int64_t j) {   // layer 2 of the j-th local tile: Qacc = A2(tmem) * W2^T
        mbar_wait(bar(A2_FULL), j & 1);
        mbar_wait(bar(Q_EMPTY), (j & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t wb = w2base + (ks >> 2) * kChunkBytes + (ks & 3) * 32;
          const uint64_t bhi = make_desc(wb), blo = make_desc(wb + kTileBytes);
          mma_ts(tm_q, tm_a2hi + ks * 8, bhi, kIdesc, ks > 0);
          mma_ts(tm_q, tm_a2lo + ks * 8, bhi, kIdesc, 1);
          mma_ts(tm_q, tm_a2hi + ks * 8, blo, kIdesc, 1);
        }
        tc_commit(bar(Q_FULL));
        tc_commit(bar(A2_EMPTY));
      };
      mbar_wait(bar(W2_FULL), 0);
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(bar(H1_EMPTY + b), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < nchunks; ++kc) {
          mbar_wait(bar(A_FULL + as), aph);
          mbar_wait(bar(W_FULL + ws), wph);
          tc_fence_after();
          const uint32_t ab = smem_u32(smem + kOffARing + as * kChunkBytes);
          const uint32_t wb = smem_u32(smem + kOffWRing + ws * kChunkBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ahi = make_desc(ab + ks * 32), alo = make_desc(ab + kTileBytes + ks * 32);
            const uint64_t bhi = make_desc(wb + ks * 32), blo = make_desc(wb + kTileBytes + ks * 32);
            mma_ss(tm_h1[b], ahi, bhi, kIdesc, (kc | ks) != 0);
            mma_ss(tm_h1[b], alo, bhi, kIdesc, 1);
            mma_ss(tm_h1[b], ahi, blo, kIdesc, 1);
          }
          tc_commit(bar(A_EMPTY + as));
          tc_commit(bar(W_EMPTY + ws));
          if (++as == kAStages) { as = 0; aph ^= 1; }
          if (++ws == kWStages) { ws = 0; wph ^= 1; }
        }
        tc_commit(bar(H1_FULL + b));
        if (it > 0) issue_l2(it - 1);
      }
      if (it > 0) issue_l2(it - 1);
    }
  } else {
    // =============================== epilogue warps (TMEM lane quadrant = warp) ===============
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const int row_in_tile = warp * 32 + lane;
    auto q_epilogue = [&](int64_t j, int64_t tile) {
      mbar_wait(bar(Q_FULL), j & 1);
      tc_fence_after();
      const int64_t n = tile * kTileM + row_in_tile;
#pragma unroll 1
      for (int c0 = 0; c0 < kQ; c0 += 32) {
        uint32_t v[32];
        DSMIL_TMEM_LD32(tm_q + lane_sel + c0, v);
        tmem_wait_ld();
        if (c0 == kQ - 32) { tc_fence_before(); mbar_arrive(bar(Q_EMPTY)); }
        if (n < a.N) {
          float4* dst = reinterpret_cast<float4*>(a.Q + n * kQ + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o;
            o.x = tanhf(__uint_as_float(v[4 * q + 0]) + s_b2[c0 + 4 * q + 0]);
            o.y = tanhf(__uint_as_float(v[4 * q + 1]) + s_b2[c0 + 4 * q + 1]);
            o.z = tanhf(__uint_as_float(v[4 * q + 2]) + s_b2[c0 + 4 * q + 2]);
            o.w = tanhf(__uint_as_float(v[4 * q + 3]) + s_b2[c0 + 4 * q + 3]);
            dst[q] = o;
          }
        }
      }
    };
    int64_t it = 0, prev_tile = -1;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const int64_t n = tile * kTileM + row_in_tile;
      mbar_wait(bar(H1_FULL + b), (it >> 1) & 1);
      mbar_wait(bar(A2_EMPTY), (it & 1) ^ 1);     // layer 2 of the previous tile has consumed A2
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < kQ; c0 += 32) {
        uint32_t v[32];
        DSMIL_TMEM_LD32(tm_h1[b] + lane_sel + c0, v);
        tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float h0 = fmaxf(__uint_as_float(v[2 * q]) + s_b1[c0 + 2 * q], 0.f);
          const float h1 = fmaxf(__uint_as_float(v[2 * q + 1]) + s_b1[c0 + 2 * q + 1], 0.f);
          v[2 * q] = __float_as_uint(h0);
          v[2 * q + 1] = __float_as_uint(h1);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(h0 - __low2float(hh), h1 - __high2float(hh));
          hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        DSMIL_TMEM_ST16(tm_a2hi + lane_sel + (c0 >> 1), hi);
        DSMIL_TMEM_ST16(tm_a2lo + lane_sel + (c0 >> 1), lo);
        if (a.H1 != nullptr && n < a.N) {
          float4* dst = reinterpret_cast<float4*>(a.H1 + n * kQ + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                 __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(H1_EMPTY + b));
      mbar_arrive(bar(A2_FULL));
      if (it > 0) q_epilogue(it - 1, prev_tile);
      prev_tile = tile;
    }
    if (it > 0) q_epilogue(it - 1, prev_tile);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

inline size_t qmlp_smem_bytes(int C, int D) { return kSmemFixed + sizeof(float) * C * D + 1024; }
inline size_t wimg_bytes(int D) { return static_cast<size_t>(D / kChunkK) * kChunkBytes + 2 * kChunkBytes; }
inline bool qmlp_supported(const dsmil_params_t* p) {
  return p->nonlinear && p->D % kChunkK == 0 && p->D >= kChunkK && qmlp_smem_bytes(p->C, p->D) <= 232448 &&
         (reinterpret_cast<uintptr_t>(p->W1) % 4 == 0);
}

// scores + arg-max keys + Q (+H1) for N rows.  wimg: >= wimg_bytes(D) bytes of workspace.
inline int launch_qmlp(const dsmil_params_t* p, const float* X, int64_t N, float* classes,
                       unsigned long long* keys, float* Q, float* H1, uint8_t* wimg, int num_sms,
                       cudaStream_t st) {
  const int D = p->D, C = p->C;
  uint8_t* w1img = wimg;
  uint8_t* w2img = wimg + static_cast<size_t>(D / kChunkK) * kChunkBytes;
  k_prep_wimg<<<64, 256, 0, st>>>(p->W1, D, w1img);
  DSMIL_LAUNCH_OK("k_prep_wimg(W1)");
  k_prep_wimg<<<16, 256, 0, st>>>(p->W2, kQ, w2img);
  DSMIL_LAUNCH_OK("k_prep_wimg(W2)");
  QmlpArgs a{X, N, D, C, p->Wi, p->bi, p->b1, p->b2, w1img, w2img, classes, keys, Q, H1};
  const size_t smem = qmlp_smem_bytes(C, D);
  const int64_t tiles = (N + kTileM - 1) / kTileM;
  const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  auto go = [&](auto kern) -> int {
    DSMIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    prof_begin(PROF_FUSED, st);
    kern<<<grid, kThreads, smem, st>>>(a);
    prof_end(PROF_FUSED, st);
    DSMIL_LAUNCH_OK("k_qmlp_sm100");
    return 0;
  };
  if (C == 1) return go(k_qmlp_sm100<1>);
  if (C == 2) return go(k_qmlp_sm100<2>);
  if (C <= 4) return go(k_qmlp_sm100<4>);
  return go(k_qmlp_sm100<8>);
}

}  // namespace sm100
}  // namespace dsmil
