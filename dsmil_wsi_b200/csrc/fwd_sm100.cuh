// sm_100a tensor-core path of phase 1 (dsmil.py:11 scores + :49 Q-MLP), D % 128 == 0.
//
//   k_prep_wimg2    packs W1 / W2 into bf16 hi/lo "images" that are byte-for-byte the shared-memory
//                   operand tiles tcgen05.mma reads (K-major, SWIZZLE_128B), so they can be brought
//                   in with plain 1-D bulk copies (cp.async.bulk -> SASS UBLKCP), no tensor map.
//   k_qmlp_sm100    persistent, warp-specialised, one CTA per SM, 128-row tiles walked through a bag table
//                   (ragged batch of bags), 896 threads in warpgroup-aligned roles (setmaxnreg 80/72/40):
//        converters x16  : streaming 16-byte loads of X (one 32 KB chunk ahead, across tile boundaries) ->
//                          fp32 FFMA instance scores (+ packed arg-max key per bag) -> x = hi + lo (two bf16)
//                          -> swizzled st.shared operand tiles; release to the fence warp (no MEMBAR here)
//        fence warp      : fence.proxy.async per stage, then publishes A_FULL to the MMA issuer
//        W producer      : streams the W1 chunks and, per tile, the two W2 chunks into a 2-stage smem ring
//        MMA issuer      : one thread issues tcgen05.mma (M=128,N=128,K=16, bf16 in, fp32 accumulate in
//                          TMEM); 3 products per K-step: hi*Whi + lo*Whi + hi*Wlo  ("3xBF16", error at
//                          the fp32 noise floor -- SURVEY A.4, tests/test_oracle.py); layer 2: A from TMEM
//        epilogue x8     : H1 = relu(acc + b1) -> bf16 hi/lo written BACK to TMEM (tcgen05.st) as the A
//                          operand of layer 2; Q = tanh(acc2 + b2) (packed f32x2 math) -> global, in tile
//                          blocks (column-major, coalesced) or row-major when training keeps it
//   smem (196 KB): W ring 2 x 32 KB | A ring 4 x 32 KB (hi tile + lo tile) | Wi rows
//   TMEM columns (512): H1 accumulators 2 x 128 | Q accumulator 128 | A2 hi 64 | A2 lo 64
//   Measured state and what limits it: DESIGN.md section 4.1, profiles/.  Round-2 findings (profiles/r2_bench_history.md):
//   neither ring depth (A 2-4 / W 2-4 stages), nor the load chain in isolation (tools/chainbench: 6.5 TB/s), nor deeper
//   epilogue unrolling changes the 117 us per 16-bag step; the CTA-0 timeline (tools/ktrace.py) shows the eight epilogue
//   warps busy ~14 k of the ~20 k cycles of a tile (H1 epilogue 4.2 k, Q epilogue incl. the 64 KB of tile-blocked stores
//   9.5 k): the epilogue role paces the pipeline.
#pragma once
#include <cuda_bf16.h>
#include <type_traits>
#include <cstdlib>

#include "common.cuh"

namespace dsmil {
namespace sm100 {

constexpr int kTileM = 128;            // rows per tile == UMMA M
constexpr int kChunkK = 64;            // k per smem operand chunk: 64 bf16 = 128 B = one swizzle row
constexpr int kTileBytes = kTileM * kChunkK * 2;       // 16 KiB: one [128 x 64] bf16 operand tile
constexpr int kChunkBytes = 2 * kTileBytes;            // hi tile + lo tile
constexpr int kAStages = 4;            // converters may run 4 chunks (128 KB) ahead of the tensor core
constexpr int kWStages = 2;
constexpr int kConvWarps = 16;           // 512 converter threads: 4 float4 per thread per 32 KB chunk
constexpr int kEpiWarps = 8;             // two warps per TMEM lane quadrant, each takes 64 of the 128 columns
// Roles are aligned to warpgroups (4 warps) so that setmaxnreg can move registers between them:
// WG0-1 epilogue | WG2-3 converters | WG4 = MMA issuer, W producer, 2 idle warps
constexpr int kWarpConv0 = kEpiWarps, kWarpMma = kEpiWarps + kConvWarps, kWarpTma = kWarpMma + 1;
constexpr int kThreads = 32 * (kEpiWarps + kConvWarps + 4);
// The pool is what the CTA got at launch: 72 regs x 896 threads = 64512 (an .inc blocks until .dec frees enough)
constexpr int kRegsConv = 80, kRegsEpi = 72, kRegsCtl = 40;    // 512*80 + 256*72 + 128*40 = 64512
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {   // read-once data: keep it out of L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
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
// Waiting roles share warp schedulers with the busy ones and the kernel is issue-bound, so a failed probe
// backs off with nanosleep (sleep_ns > 0) instead of re-polling at full rate.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = 0) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (sleep_ns) __nanosleep(sleep_ns);
    if (++spins > (1u << 26)) __trap();  // a protocol bug must not hang the GPU
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
#define DSMIL_TMEM_LD16(taddr, v)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                            \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),       \
                 "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),   \
                 "=r"(v[14]), "=r"(v[15])                                                                  \
               : "r"(taddr) : "memory")
#define DSMIL_TMEM_ST8(taddr, v)                                                                           \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                    \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),  \
                 "r"(v[7]) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version [46,48) = 1 (sm_100) | layout [61,64) = 2 (SWIZZLE_128B)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// The same descriptor as two 32-bit words: only the low word (start address) changes between MMAs.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3ffffu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_ss2(uint32_t d, uint32_t alo, uint32_t blo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %3, 0;\n\t"
      "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d), "r"(alo), "r"(blo), "r"(acc), "r"(kDescHi), "r"(kIdesc) : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %3, 0;\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(acc), "r"(kDescHi), "r"(kIdesc) : "memory");
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)

// byte offset of element (row, k) inside one [rows x 64] bf16 SWIZZLE_128B tile
__host__ __device__ inline uint32_t swz_off(int row, int k) {
  return static_cast<uint32_t>(row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1));
}

// ---- weight images -----------------------------------------------------------------------
// img layout: for each 64-wide k chunk: [hi tile 16 KiB][lo tile 16 KiB], rows = output features (128)
// (written by k_prep_wimg2 below, once per parameter version)

// One bag of a batch, as the kernels see it (device memory, built by the host per call).
struct BagDev {
  const float* X;       // [N, D]
  long long N;
  long long row_off;    // first row of this bag in the packed outputs (classes, A, Q, H1)
  int tile_off;         // first 128-row tile of this bag in the batch-wide tile numbering
  int rec_off;          // first partial record of this bag
  int nrec;             // partial records (== attend CTAs) of this bag
  int pad_;
};

extern long long* g_trace_buf;   // device buffer (3*8*64 int64) or NULL
inline int debug_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("DSMIL_B200_DEBUG_MODE"); m = e ? atoi(e) : 0; }
  return m;
}

constexpr int kSmemBags = 96;

struct QmlpArgs {
  const BagDev* bags;
  int bag0, nb;           // bags [bag0, bag0+nb) are covered by this launch ...
  int tile0, ntiles;      // ... which are tiles [tile0, tile0+ntiles)
  int D, C;
  const float* Wi;
  const float* bi;
  const float* b1;
  const float* b2;
  const uint8_t* w1img;   // D/64 chunks
  const uint8_t* w2img;   // 2 chunks
  float* classes;         // packed [sumN, C] or NULL (scores given by the caller)
  unsigned long long* keys;  // [nbags][kMaxC]
  float* Q;               // packed row-major [sumN,128], or (q_blocked) per-tile column-major blocks [tile][128 col][128 row]
  float* H1;              // packed [sumN,128] or NULL
  long long* dbg;         // optional timeline of CTA 0 (clock64 stamps), see DSMIL_B200_TRACE in abi.cu
  int q_blocked;          // 1: Q is written in tile blocks (coalesced epilogue stores; inference path); 2: tile blocks of the
                          // PRE-activation z2 = acc + b2 -- the tanh moves to the readers of Q (k_attend_b, k_gather_cand_b):
                          // the epilogue role paces this kernel (profiles/r2_ktrace_old.txt), its readers have idle issue slots
  int mode;               // timing experiments only (DSMIL_B200_DEBUG_MODE): bit0 no Q stores, bit1 no scores,
                          // bit2 converter skips convert+store, bit3 no MMAs, bit4 epilogue skips math
};
// trace slots: [role][event][index] -> role 0 converter (warp 0 of the role), 1 mma, 2 epilogue
#define DSMIL_TRACE(role, ev, idx)                                                            \
  do {                                                                                        \
    if (a.dbg != nullptr && blockIdx.x == 0 && (idx) < 64)                                    \
      a.dbg[((role) * 8 + (ev)) * 64 + (idx)] = clock64();                                    \
  } while (0)

// Walks the bag table as a role's tile index increases monotonically.
struct TileCursor {
  const BagDev* bags;
  int bag, last;
  __device__ TileCursor(const BagDev* b, int bag0, int nb) : bags(b), bag(bag0), last(bag0 + nb - 1) {}
  __device__ __forceinline__ void seek(int tile) {
    while (bag < last && tile >= bags[bag + 1].tile_off) ++bag;
  }
};

// Packed fp32x2 arithmetic (sm_100: add/mul/fma.f32x2 issue one instruction for two lanes' values) -- the
// epilogue is issue-bound, so halving its FADD/FMUL/FFMA count matters.  Results are bit-identical to scalar ops.
struct f2 { float x, y; };
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
// two tanh at once: 1 - 2/(exp2(x * 2log2e) + 1); same formula and intrinsics as fast_tanh
__device__ __forceinline__ f2 fast_tanh2(f2 x) {
  const f2 t = mul2(x, f2{2.8853900817779268f, 2.8853900817779268f});
  f2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const f2 d = add2(e, f2{1.f, 1.f});
  f2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return fma2(r, f2{-2.f, -2.f}, f2{1.f, 1.f});
}
__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 1 - 2/(exp(2x)+1); ex2.approx + rcp.approx: |err| < ~3e-7 absolute, saturates correctly
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

// L2 eviction policy for cache-hinted loads (the pair kernel keeps X resident for the attention pass)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol;
}

// dynamic smem carve (bytes, from a 1024-aligned base)
constexpr int kOffWRing = 0;                                // kWStages x 32 KiB: W1 chunks AND the two W2 chunks stream here
constexpr int kOffARing = kOffWRing + kWStages * kChunkBytes;
constexpr int kOffWi = kOffARing + kAStages * kChunkBytes;  // C*D floats
constexpr int kSmemFixed = kOffWi;

// CT: classes rounded up to 1/2/4/8.  DT: compile-time feature size (512 = every shipped configuration:
// the chunk loops unroll and the load offsets become immediates) or 0 = run-time D.
template <int CT, int DT>
__global__ void __launch_bounds__(kThreads, 1)
k_qmlp_sm100(const QmlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_b1[kQ], s_b2[kQ];
  __shared__ BagDev s_bags[kSmemBags];           // the launch's slice of the bag table (tile-boundary lookups)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = DT ? DT : a.D, C = a.C;
  const int nchunks = DT ? DT / kChunkK : a.D / kChunkK;
  const int tile_end = a.tile0 + a.ntiles;

  // barrier indices
  enum { A_FULL = 0, A_EMPTY = A_FULL + kAStages, A_WRITTEN = A_EMPTY + kAStages, W_FULL = A_WRITTEN + kAStages, W_EMPTY = W_FULL + kWStages,
         H1_FULL = W_EMPTY + kWStages, H1_EMPTY = H1_FULL + 2, A2_FULL = H1_EMPTY + 2, A2_EMPTY, Q_FULL, Q_EMPTY,
         NBARS };
  static_assert(NBARS <= 32, "too many barriers");
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  float* sWi = reinterpret_cast<float*>(smem + kOffWi);
  const bool do_scores = a.classes != nullptr && !(a.mode & 2);   // bag form (scores given): Wi/bi may be NULL
  if (do_scores)
    for (int i = tid; i < CT * D; i += kThreads) sWi[i] = (i < C * D) ? a.Wi[i] : 0.f;
  if (tid < kQ) { s_b1[tid] = a.b1[tid]; s_b2[tid] = a.b2[tid]; }
  const bool tbl_in_smem = a.nb <= kSmemBags;
  if (tbl_in_smem)
    for (int i = tid; i < a.nb; i += kThreads) s_bags[i] = a.bags[a.bag0 + i];
  // cursors below index the table relative to bag0 when it is cached in shared memory
  const BagDev* tbl = tbl_in_smem ? s_bags : a.bags + a.bag0;
  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(bar(A_WRITTEN + s), kConvWarps); mbar_init(bar(A_FULL + s), 1); mbar_init(bar(A_EMPTY + s), 1);
    }
    for (int s = 0; s < kWStages; ++s) { mbar_init(bar(W_FULL + s), 1); mbar_init(bar(W_EMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar(H1_FULL + b), 1); mbar_init(bar(H1_EMPTY + b), kEpiWarps * 32); }
    mbar_init(bar(A2_FULL), kEpiWarps * 32); mbar_init(bar(A2_EMPTY), 1);
    mbar_init(bar(Q_FULL), 1); mbar_init(bar(Q_EMPTY), kEpiWarps * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {  // TMEM allocation (whole 512 columns; one CTA per SM by construction)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) DSMIL_TRACE(2, 7, 0);   // prologue done
  const uint32_t tmem = s_tmem_base;
  const uint32_t tm_h1_0 = tmem, tm_h1_1 = tmem + 128;
  const uint32_t tm_q = tmem + 256, tm_a2hi = tmem + 384, tm_a2lo = tmem + 448;

  if (warp >= kWarpConv0 && warp < kWarpMma) {
    // =============================== converter warps =========================================
    reg_inc<kRegsConv>();                        // fed by the control warps' .dec
    const int ct = tid - kWarpConv0 * 32;        // 0..511
    const int seg = ct & 15, r0 = ct >> 4;       // float4 index in the 64-float chunk row; base row (0..31)
    const uint32_t off0 = swz_off(r0, seg * 4);  // rows r0 + 32 i share (row & 7): offset_i = off0 + 4096 i
    const uint32_t a_ring_u32 = smem_u32(smem + kOffARing), swi_u32 = smem_u32(smem + kOffWi);
    uint32_t stage = 0, phase = 0;
    TileCursor cur_bag(tbl, 0, a.nb);
    // The loop is flat over (tile, k-chunk): `nxt` always holds the NEXT chunk -- of this tile or the first
    // chunk of the CTA's next tile -- so the HBM stream never drains at a tile boundary.
    int tile = a.tile0 + blockIdx.x;
    // state of the tile whose chunks are being LOADED (may already be the next tile): 32-bit row numbers
    uint32_t ld_N = 0, ld_row = 0;               // rows in the bag, first row of the tile (+ r0)
    bool ld_full = false;                        // whole 128-row tile inside the bag: unpredicated loads
    long long ld_rowoff = 0;
    const float4* xrow = nullptr;
    auto open_tile = [&](int t) {
      cur_bag.seek(t);
      const BagDev* bp = tbl + cur_bag.bag;
      ld_N = static_cast<uint32_t>(bp->N);
      ld_rowoff = bp->row_off;
      ld_row = static_cast<uint32_t>(t - bp->tile_off) * kTileM + r0;
      ld_full = ld_row - r0 + kTileM <= ld_N;
      xrow = reinterpret_cast<const float4*>(bp->X + static_cast<long long>(ld_row) * D) + seg;
    };
    auto load_chunk = [&](int kc, float4* dst) {
      if (ld_full) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = ldg_stream(xrow + static_cast<long long>(i) * (8ll * D) + kc * (kChunkK / 4));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4* src = xrow + static_cast<long long>(i) * (8ll * D) + kc * (kChunkK / 4);   // 32 rows apart
          dst[i] = (ld_row + 32 * i < ld_N) ? ldg_stream(src) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    float4 cur[4], nxt[4];
    if (tile < tile_end) { open_tile(tile); load_chunk(0, cur); }
    float sc[4][CT];
    // sWi rows >= C are zero-filled (CT rows are staged), so the class loop needs no bound check
    auto process = [&](const float4* now, int kc) {     // convert + publish chunk kc held in `now`
      float4 wk[CT];
      if (do_scores) {
#pragma unroll
        for (int k = 0; k < CT; ++k) wk[k] = lds128(swi_u32 + static_cast<uint32_t>(k * D + kc * kChunkK + seg * 4) * 4u);
      }
      mbar_wait(bar(A_EMPTY + stage), phase ^ 1, 128);
      const uint32_t hi_tile = a_ring_u32 + stage * kChunkBytes + off0;   // shared-space addresses: STS with
      const uint32_t lo_tile = hi_tile + kTileBytes;                       // immediate offsets
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = now[i];
        if (do_scores) {
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            float s = sc[i][k];
            s = fmaf(x.x, wk[k].x, s); s = fmaf(x.y, wk[k].y, s); s = fmaf(x.z, wk[k].z, s); s = fmaf(x.w, wk[k].w, s);
            sc[i][k] = s;
          }
        }
        // x = hi + lo, both bf16: hi = RN(x); lo = RN(x - hi); unpacking a bf16 pair is one shift + one mask
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
        const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(u01 << 16), x.y - __uint_as_float(u01 & 0xffff0000u));
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(u23 << 16), x.w - __uint_as_float(u23 & 0xffff0000u));
        sts64(hi_tile + i * 4096, u01, u23);
        sts64(lo_tile + i * 4096, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      // No proxy fence here: fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR would
      // wait for this thread's prefetched global loads (a full HBM latency per chunk).  The stores are released to
      // the fence warp (mbarrier arrive = release), which has no loads in flight, fences, and publishes A_FULL.
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(A_WRITTEN + stage));
      if (++stage == kAStages) { stage = 0; phase ^= 1; }
    };
    while (tile < tile_end) {
      const int my_bag = a.bag0 + cur_bag.bag;   // this tile's bag (the load state may move on below)
      const uint32_t t_N = ld_N, t_row = ld_row;
      const long long t_rowoff = ld_rowoff;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k) sc[i][k] = 0.f;
      const int next_tile = tile + gridDim.x;
      // two chunks per iteration: the buffers swap roles, no register copies (nchunks is even: D % 128 == 0)
#pragma unroll
      for (int kc = 0; kc < nchunks; kc += 2) {
        load_chunk(kc + 1, nxt);
        process(cur, kc);
        if (kc + 2 < nchunks) load_chunk(kc + 2, cur);
        else if (next_tile < tile_end) { open_tile(next_tile); load_chunk(0, cur); }
        process(nxt, kc + 1);
      }
      // instance scores of this tile: reduce over the 16 threads (seg) that share a row; the per-class
      // arg-max key of the tile goes straight to the bag's key slot (one atomicMax per warp and class)
      if (do_scores) {
        unsigned long long best[CT];
#pragma unroll
        for (int k = 0; k < CT; ++k) best[k] = 0ull;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t n = t_row + 32 * i;
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            float v = sc[i][k];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (seg == 0 && k < C && n < t_N) {
              v += a.bi[k];
              a.classes[(t_rowoff + n) * C + k] = v;
              const unsigned long long key = pack_key(v, n);
              best[k] = key > best[k] ? key : best[k];
            }
          }
        }
#pragma unroll
        for (int k = 0; k < CT; ++k) {
          const unsigned long long b = warp_max_u64(best[k]);
          if (lane == 0 && k < C && b) atomicMax(a.keys + static_cast<size_t>(my_bag) * kMaxC + k, b);
        }
      }
      tile = next_tile;
    }
  } else if (warp == kWarpTma + 1) {
    // ====== proxy-fence warp: generic-proxy A tiles -> async proxy, then hand the stage to the MMA issuer ======
    reg_dec<kRegsCtl>();
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = a.tile0 + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc) {
          mbar_wait(bar(A_WRITTEN + stage), phase);
          fence_proxy_async();
          mbar_arrive(bar(A_FULL + stage));
          if (++stage == kAStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kWarpMma && warp != kWarpMma && warp != kWarpTma) {
    reg_dec<kRegsCtl>();     // idle warp of the control warpgroup
  } else if (warp == kWarpTma) {
    // =============================== W1 image producer (bulk copies) ==========================
    reg_dec<kRegsCtl>();
    if (lane == 0) {
      // chunk order == the MMA issuer's consumption order: tile it: W1 chunks 0..n-1, then the two W2 chunks of
      // layer 2 of tile it-1; after the last tile its own two W2 chunks
      uint32_t stage = 0, phase = 0;
      auto push = [&](const uint8_t* src) {
        mbar_wait(bar(W_EMPTY + stage), phase ^ 1, 256);
        mbar_expect_tx(bar(W_FULL + stage), kChunkBytes);
        bulk_g2s(smem_u32(smem + kOffWRing + stage * kChunkBytes), src, kChunkBytes, bar(W_FULL + stage));
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      for (int tile = a.tile0 + blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
        for (int kc = 0; kc < nchunks; ++kc) push(a.w1img + static_cast<size_t>(kc) * kChunkBytes);
        if (it > 0) { push(a.w2img); push(a.w2img + kChunkBytes); }
      }
      if (it > 0) { push(a.w2img); push(a.w2img + kChunkBytes); }
    }
  } else if (warp == kWarpMma) {
    // =============================== MMA issuer ================================================
    reg_dec<kRegsCtl>();
    if (lane == 0) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0;
      auto issue_l2 = [&](int j) {   // layer 2 of the j-th local tile: Qacc = A2(tmem) * W2^T, W2 from the W ring
        DSMIL_TRACE(1, 3, j);
        mbar_wait(bar(A2_FULL), j & 1);
        DSMIL_TRACE(1, 4, j);
        mbar_wait(bar(Q_EMPTY), (j & 1) ^ 1);
        DSMIL_TRACE(1, 5, j);
        tc_fence_after();
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          mbar_wait(bar(W_FULL + ws), wph);
          tc_fence_after();
          const uint32_t wb = desc_lo(smem_u32(smem + kOffWRing + ws * kChunkBytes));
#pragma unroll
          for (int k4 = 0; k4 < ((a.mode & 8) ? 0 : 4); ++k4) {
            const int ks = c2 * 4 + k4;
            mma_ts2(tm_q, tm_a2hi + ks * 8, wb + k4 * 2, ks > 0);
            mma_ts2(tm_q, tm_a2lo + ks * 8, wb + k4 * 2, 1);
            mma_ts2(tm_q, tm_a2hi + ks * 8, wb + (kTileBytes >> 4) + k4 * 2, 1);
          }
          tc_commit(bar(W_EMPTY + ws));
          if (++ws == kWStages) { ws = 0; wph ^= 1; }
        }
        tc_commit(bar(Q_FULL));
        tc_commit(bar(A2_EMPTY));
      };
      int it = 0;
      for (int tile = a.tile0 + blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(bar(H1_EMPTY + b), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < nchunks; ++kc) {
          mbar_wait(bar(A_FULL + as), aph);
          if (it == 2) DSMIL_TRACE(1, 0, kc);
          mbar_wait(bar(W_FULL + ws), wph);
          if (it == 2) DSMIL_TRACE(1, 1, kc);
          tc_fence_after();
          const uint32_t ab = desc_lo(smem_u32(smem + kOffARing + as * kChunkBytes));
          const uint32_t wb = desc_lo(smem_u32(smem + kOffWRing + ws * kChunkBytes));
          const uint32_t dacc = b ? tm_h1_1 : tm_h1_0;
#pragma unroll
          for (int ks = 0; ks < ((a.mode & 8) ? 0 : 4); ++ks) {   // +32 B per K-step == +2 in descriptor units; lo tile is +1024 units
            mma_ss2(dacc, ab + ks * 2, wb + ks * 2, (kc | ks) != 0);
            mma_ss2(dacc, ab + (kTileBytes >> 4) + ks * 2, wb + ks * 2, 1);
            mma_ss2(dacc, ab + ks * 2, wb + (kTileBytes >> 4) + ks * 2, 1);
          }
          tc_commit(bar(A_EMPTY + as));
          tc_commit(bar(W_EMPTY + ws));
          if (it == 2) DSMIL_TRACE(1, 2, kc);
          if (++as == kAStages) { as = 0; aph ^= 1; }
          if (++ws == kWStages) { ws = 0; wph ^= 1; }
        }
        tc_commit(bar(H1_FULL + b));
        if (it > 0) issue_l2(it - 1);
      }
      if (it > 0) issue_l2(it - 1);
    }
  } else {
    // ====== epilogue warps: TMEM lane quadrant = warp & 3, column half = warp >> 2 =============
    // (epilogue warps stay at the launch register count, kRegsEpi)
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row_in_tile = (warp & 3) * 32 + lane;
    const int col0 = (warp >> 2) * 64;
    auto q_epilogue = [&](int j, long long grow, bool live, int qtile) {   // grow: packed row of this thread
      mbar_wait(bar(Q_FULL), j & 1, 256);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + 64; c0 += 16) {
        uint32_t v[16];
        DSMIL_TMEM_LD16(tm_q + lane_sel + c0, v);
        tmem_wait_ld();
        if (c0 == col0 + 48) { tc_fence_before(); mbar_arrive(bar(Q_EMPTY)); }
        if (a.q_blocked) {
          // tile-blocked, column-major: lanes (= rows) are contiguous -> every store is one 128-byte line
          float* dst = a.Q + static_cast<size_t>(qtile) * (kTileM * kQ) + static_cast<size_t>(c0) * kTileM + row_in_tile;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(&s_b2[c0 + 4 * q]);
            f2 t0 = add2(f2{__uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1])}, f2{bb.x, bb.y});
            f2 t1 = add2(f2{__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])}, f2{bb.z, bb.w});
            if (a.q_blocked == 1) { t0 = fast_tanh2(t0); t1 = fast_tanh2(t1); }
            dst[(4 * q + 0) * kTileM] = t0.x;
            dst[(4 * q + 1) * kTileM] = t0.y;
            dst[(4 * q + 2) * kTileM] = t1.x;
            dst[(4 * q + 3) * kTileM] = t1.y;
          }
        } else if (live && !(a.mode & 16)) {
          float4* dst = reinterpret_cast<float4*>((a.mode & 1) ? a.Q + (threadIdx.x & 255) * 4 : a.Q + grow * kQ + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(&s_b2[c0 + 4 * q]);
            // same tanh formulation as the tile-blocked (inference) store: train and eval give the same Q bits
            const f2 t0 = fast_tanh2(add2(f2{__uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1])}, f2{bb.x, bb.y}));
            const f2 t1 = fast_tanh2(add2(f2{__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])}, f2{bb.z, bb.w}));
            dst[q] = make_float4(t0.x, t0.y, t1.x, t1.y);
          }
        }
      }
    };
    int it = 0, prev_tile = 0;
    long long prev_grow = 0;
    bool prev_live = false;
    TileCursor cur_bag(tbl, 0, a.nb);
    for (int tile = a.tile0 + blockIdx.x; tile < tile_end; tile += gridDim.x, ++it) {
      cur_bag.seek(tile);
      const BagDev bg = tbl[cur_bag.bag];
      const long long n = static_cast<long long>(tile - bg.tile_off) * kTileM + row_in_tile;
      const bool live = n < bg.N;
      const long long grow = bg.row_off + n;
      const int b = it & 1;
      mbar_wait(bar(H1_FULL + b), (it >> 1) & 1, 512);
      if (tid == 0) DSMIL_TRACE(2, 0, it);
      mbar_wait(bar(A2_EMPTY), (it & 1) ^ 1, 128);     // layer 2 of the previous tile has consumed A2
      if (tid == 0) DSMIL_TRACE(2, 3, it);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + 64; c0 += 16) {
        uint32_t v[16];
        DSMIL_TMEM_LD16((b ? tm_h1_1 : tm_h1_0) + lane_sel + c0, v);
        tmem_wait_ld();
        uint32_t hi[8], lo[8];
        if (a.mode & 16) {
#pragma unroll
          for (int q = 0; q < 8; ++q) { hi[q] = v[2 * q]; lo[q] = v[2 * q + 1]; }
        } else
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 bb = *reinterpret_cast<const float2*>(&s_b1[c0 + 2 * q]);
          const f2 z = add2(f2{__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])}, f2{bb.x, bb.y});
          const float h0 = fmaxf(z.x, 0.f), h1 = fmaxf(z.y, 0.f);
          v[2 * q] = __float_as_uint(h0);
          v[2 * q + 1] = __float_as_uint(h1);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
          const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh);
          const f2 res = add2(f2{h0, h1}, f2{-__uint_as_float(hu << 16), -__uint_as_float(hu & 0xffff0000u)});
          const __nv_bfloat162 ll = __floats2bfloat162_rn(res.x, res.y);
          hi[q] = hu;
          lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        DSMIL_TMEM_ST8(tm_a2hi + lane_sel + (c0 >> 1), hi);
        DSMIL_TMEM_ST8(tm_a2lo + lane_sel + (c0 >> 1), lo);
        if (a.H1 != nullptr && live) {
          float4* dst = reinterpret_cast<float4*>(a.H1 + grow * kQ + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                 __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(H1_EMPTY + b));
      mbar_arrive(bar(A2_FULL));
      if (tid == 0) DSMIL_TRACE(2, 1, it);
      if (it > 0) q_epilogue(it - 1, prev_grow, prev_live, prev_tile);
      if (tid == 0) DSMIL_TRACE(2, 2, it);
      prev_grow = grow;
      prev_live = live;
      prev_tile = tile;
    }
    if (it > 0) q_epilogue(it - 1, prev_grow, prev_live, prev_tile);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// one kernel for both images
__global__ void __launch_bounds__(256)
k_prep_wimg2(const float* __restrict__ W1, int D, const float* __restrict__ W2, uint8_t* __restrict__ img1,
             uint8_t* __restrict__ img2) {
  const int t1 = 128 * D, total = t1 + 128 * kQ;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool first = i < t1;
    const int K = first ? D : kQ;
    const int j = first ? i : i - t1;
    const int n = j / K, k = j % K;
    const float w = first ? W1[j] : W2[j];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    uint8_t* chunk = (first ? img1 : img2) + static_cast<size_t>(k / kChunkK) * kChunkBytes;
    const uint32_t off = swz_off(n, k % kChunkK);
    *reinterpret_cast<__nv_bfloat16*>(chunk + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(chunk + kTileBytes + off) = lo;
  }
}

inline size_t qmlp_smem_bytes(int C, int D) {   // Wi is staged with C rounded up to 1, 2, 4, 8 rows
  const int ct = C <= 1 ? 1 : (C <= 2 ? 2 : (C <= 4 ? 4 : 8));
  return kSmemFixed + sizeof(float) * ct * D + 1024;
}
inline size_t wimg_bytes(int D) { return static_cast<size_t>(D / kChunkK) * kChunkBytes + 2 * kChunkBytes; }
inline bool qmlp_supported(const dsmil_params_t* p) {
  return p->nonlinear && p->D % (2 * kChunkK) == 0 && qmlp_smem_bytes(p->C, p->D) <= 232448;
}

inline int launch_prep_wimg(const dsmil_params_t* p, uint8_t* wimg, cudaStream_t st) {
  uint8_t* w2img = wimg + static_cast<size_t>(p->D / kChunkK) * kChunkBytes;
  k_prep_wimg2<<<80, 256, 0, st>>>(p->W1, p->D, p->W2, wimg, w2img);
  DSMIL_LAUNCH_OK("k_prep_wimg2");
  return 0;
}

// scores + arg-max keys + Q (+H1) for the tiles [tile0, tile0+ntiles) of bags [bag0, bag0+nb).
// wimg must already hold the images (launch_prep_wimg).
inline int launch_qmlp(const dsmil_params_t* p, const BagDev* bags_dev, int bag0, int nb, int tile0, int ntiles,
                       float* classes, unsigned long long* keys, float* Q, float* H1, const uint8_t* wimg,
                       int num_sms, cudaStream_t st, int q_blocked = 0) {
  const int D = p->D, C = p->C;
  const uint8_t* w2img = wimg + static_cast<size_t>(D / kChunkK) * kChunkBytes;
  QmlpArgs a{bags_dev, bag0, nb, tile0, ntiles, D, C, p->Wi, p->bi, p->b1, p->b2, wimg, w2img, classes, keys, Q, H1,
             g_trace_buf, q_blocked, debug_mode()};
  const size_t smem = qmlp_smem_bytes(C, D);
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  auto go = [&](auto kern) -> int {
    // (set on every launch: the instantiations share one function-pointer TYPE, so a cached flag here would be
    //  shared between them -- found by the D=1024 / C=1 shape tests; the call costs ~1 us of host time)
    DSMIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    prof_begin(PROF_FUSED, st);
    kern<<<grid, kThreads, smem, st>>>(a);
    prof_end(PROF_FUSED, st);
    DSMIL_LAUNCH_OK("k_qmlp_sm100");
    return 0;
  };
  if (D == 512) {
    if (C == 1) return go(k_qmlp_sm100<1, 512>);
    if (C == 2) return go(k_qmlp_sm100<2, 512>);
    if (C <= 4) return go(k_qmlp_sm100<4, 512>);
    return go(k_qmlp_sm100<8, 512>);
  }
  if (C == 1) return go(k_qmlp_sm100<1, 0>);
  if (C == 2) return go(k_qmlp_sm100<2, 0>);
  if (C <= 4) return go(k_qmlp_sm100<4, 0>);
  return go(k_qmlp_sm100<8, 0>);
}

}  // namespace sm100
}  // namespace dsmil
