// CTA-pair phase 1 for D <= 512 (every shipped configuration): scores + arg-max keys + Q-MLP (dsmil.py:11, :49, :52).
//
// Why a second phase-1 kernel: profiles/r1_k1_budget.md + profiles/r2_chainbench.txt -- k_qmlp_sm100 is held back by the
// units its converter chain shares with the tensor pipe: SS-mode MMAs re-read both operands from shared memory for each
// of the three split products (64 wavefronts per MMA), the W1/W2 images are re-streamed from L2 for every 128-row tile
// (as many bytes as X itself) and X is written to shared memory a second time as bf16.  This kernel removes all three:
//
//   * two CTAs of one TPC form a pair (cluster of 2, tcgen05 cta_group::2, M = 256: 128 rows per CTA).  Each CTA
//     keeps HALF of the weight images (64 of the 128 output features of W1 and W2, bf16 hi/lo) RESIDENT in shared memory
//     for the whole kernel: no W streaming, and the B-operand fetch per CTA is halved;
//   * the A operand lives in TENSOR MEMORY: X arrives by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of
//     [128 rows x 32 floats] = 16 KB, out-of-bag rows zero-filled by the TMA unit), converter warps (lane = row) read
//     a box with conflict-free LDS.128, add it to the fp32 instance scores, split x = hi + lo (two bf16) and tcgen05.st
//     both halves into a TMEM operand stage -- no bf16 copy of X in shared memory, no A-operand wavefronts;
//   * layer 2 as before from TMEM (H1 -> bf16 hi/lo -> tcgen05.st), its accumulator ALIASES the drained layer-1
//     accumulator of the same tile, which frees the TMEM columns the X operand stages need.
//
//   warps: 0-7 epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2) | 8-23 converters in 4 groups of 4
//          (one warp per lane quadrant; group g takes the boxes g, g+4, ... of the CTA's box sequence and owns TMEM
//          operand stage g: four boxes are in conversion at any time, which is what hides the per-box chain
//          TMA -> LDS -> split -> tcgen05.st -> cross-CTA arrive -> MMA -> commit; measured variants in
//          profiles/r2_bench_history.md) | 24 MMA issuer (leader CTA only) | 25 TMA producer |
//          26 TMEM allocation + weight load | 27 idle
//   smem : W1 half (D/64 x 16 KB) | W2 half (2 x 16 KB) | X staging 3 slots x 16 KB | Wi | score partials
//   TMEM : acc0 128 | acc1 128 (layer-1 accumulator, then layer-2 accumulator of the same tile) | A2 hi 64 | A2 lo 64 |
//          X operand stages 4 x (hi 16 + lo 16): stage = box % 4 = group
//   barriers: a staging slot (box % 3) serves the four groups in turn, so its full/empty barriers are indexed by
//          box % 12 (fixed group, fixed slot): every mbarrier has ONE waiting party that sees every phase in order
//          (a barrier per slot would let a group wait for use u+1 before use u has completed)
//   cross-CTA: converters / epilogue warps of the second CTA arrive REMOTELY (mapa) on the leader's barriers; the MMA
//          issuer answers with multicast tcgen05.commit on the barrier of the same offset in both CTAs.
// Mechanisms first validated in isolation by tools/probe_pair.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "fwd_sm100.cuh"

namespace dsmil {
namespace pair {

using sm100::BagDev;
using sm100::f2;
using sm100::add2;
using sm100::fast_tanh2;
using sm100::smem_u32;
using sm100::mbar_init;
using sm100::mbar_arrive;
using sm100::mbar_expect_tx;
using sm100::bulk_g2s;
using sm100::tc_fence_before;
using sm100::tc_fence_after;
using sm100::tmem_wait_ld;
using sm100::tmem_wait_st;
using sm100::lds128;
using sm100::swz_off;
using sm100::TileCursor;

constexpr int kTileM = 128;
constexpr int kBoxK = 32;                         // floats per TMA box row = 128 B = one swizzle row
constexpr int kBoxBytes = kTileM * kBoxK * 4;     // 16 KB
constexpr int kGroups = 4;                        // converter groups (box n belongs to group n % 4)
constexpr int kSlots = 3;                         // X staging slots (box n lands in slot n % 3)
constexpr int kXBars = 12;                        // full/empty barrier index = n % 12 (lcm of the two)
constexpr int kTStages = 4;                       // TMEM operand stages (box n is stored to stage n % 4 = its group)
constexpr int kWChunk = 2 * 64 * 128;             // per 64-k chunk and CTA: hi tile [64 features x 128 B] + lo tile
constexpr int kEpiWarps = 8, kConvWarps = 4 * kGroups;
constexpr int kWarpConv0 = kEpiWarps, kWarpMma = kEpiWarps + kConvWarps, kWarpProd = kWarpMma + 1, kWarpAux = kWarpMma + 2;
constexpr int kThreads = 32 * (kEpiWarps + kConvWarps + 4);     // 896
constexpr int kSmemBags = 96;
constexpr int kMaxD = 512;

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the pair.  Default (.release.cta) semantics, as
// CUTLASS' ClusterBarrier::arrive(cta_id): what the arrival publishes is TMEM content, ordered by tcgen05.wait::st +
// tcgen05.fence::before_thread_sync.  A `.release.cluster` arrive costs ~1.4 us per box here (measured,
// profiles/r2_ptrace_pair.md) -- the peer CTA then paces the pair.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar, uint32_t my_rank) {
  if (my_rank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0);
}
// CL: the arrivals come (partly) from the peer CTA -> acquire at cluster scope
template <bool CL>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = 0) {
  uint32_t done = 0, spins = 0;
  while (true) {
    (void)CL;   // (arrivals from the peer CTA need no cluster-scope acquire either: see mbar_arrive_remote)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (sleep_ns) __nanosleep(sleep_ns);
    if (++spins > (1u << 26)) __trap();          // a protocol bug must not hang the GPU
  }
}
// (L2 cache hint: evict-last -- the attend kernel re-reads X, walking the tiles backwards from the most recent one)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}
// L2 prefetch of a box (no shared memory, no barrier): issued one tile ahead so that the staging loads, which can only
// be one box per converter group in flight, pay L2 latency instead of HBM latency
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of CONTIGUOUS bytes (a 128-row tile of X is one contiguous 128 * D * 4-byte range): sequential DRAM pages
// instead of the 128 scattered 128-byte rows a box prefetch touches
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrive::one on `bar` in BOTH CTAs when the MMAs retire
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
// kind::f16, D = f32, A = B = bf16, K-major, N = 128, M = 256 (128 rows per CTA of the pair)
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %3, 0;\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %5, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "r"(blo), "r"(acc), "r"(sm100::kDescHi), "r"(kIdesc2) : "memory");
}
#define DSMIL_TMEM_ST16(taddr, v)                                                                            \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),   \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")

struct PairArgs {
  const BagDev* bags;       // [nb]
  const CUtensorMap* tmaps; // [nb]: X of bag b as a 2-D fp32 tensor {D, N}, box {32, 128}, SWIZZLE_128B
  int nb, ntiles;           // the launch covers every tile of every bag in the table
  int D, C;
  const float* Wi;
  const float* bi;
  const float* b1;
  const float* b2;
  const uint8_t* wimg;      // pair image (k_prep_wimg_pair): [rank][W1 D/64 chunks | W2 2 chunks] x 16 KB
  float* classes;           // packed [sumN, C] or NULL (scores given by the caller)
  unsigned long long* keys; // [nb][kMaxC]
  float* Q;                 // tile-blocked [tile][128 col][128 row]
  long long* dbg;           // optional timeline of CTA 0 (clock64 stamps; tools/ptrace.py), NULL = off
  int flags;                // experiments (DSMIL_B200_PAIR_FLAGS): 1 no L2 prefetch, 4 producer polls without back-off
};
// trace slots [role 8][event 8][index 128], stamps = %globaltimer (ns; comparable across the two CTAs):
// role 0 converter warp 0 of CTA 0, 1 MMA issuer, 2 epilogue warp 0 of CTA 0, 3 producer of CTA 0,
// 4 converter warp 0 of CTA 1, 5 converter warp 15 of CTA 0, 6 converter warp 15 of CTA 1, 7 epilogue warp 0 of CTA 1
__device__ __forceinline__ long long pair_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// second trace area at +8192: [event 3: XS_FULL seen, XT_EMPTY seen, XT_FULL arrived][box 40..47][cta 2][converter warp 16]
#define PAIR_TRACE_ALL(ev, n, cw)                                                                   \
  do {                                                                                              \
    if (a.dbg != nullptr && blockIdx.x < 2 && (n) >= 40u && (n) < 48u && (threadIdx.x & 31) == 0)    \
      a.dbg[8192 + (((ev) * 8 + ((n) - 40u)) * 2 + blockIdx.x) * 16 + (cw)] = pair_now();           \
  } while (0)
#define PAIR_TRACE(role, ev, idx)                                                                   \
  do {                                                                                              \
    if (a.dbg != nullptr && blockIdx.x < 2 && static_cast<uint32_t>(idx) < 128u)                    \
      a.dbg[((role) * 8 + (ev)) * 128 + (idx)] = pair_now();                                        \
  } while (0)

__host__ __device__ inline size_t pair_wimg_rank_bytes(int D) { return static_cast<size_t>(D / 64 + 2) * kWChunk; }
inline size_t pair_smem_bytes(int C, int D) {
  const int ct = C <= 1 ? 1 : (C <= 2 ? 2 : 4);
  return pair_wimg_rank_bytes(D) + kSlots * kBoxBytes + sizeof(float) * ct * D + sizeof(float) * 2 * 4 * kTileM * ct + 1024;
}
inline bool pair_supported(const dsmil_params_t* p) {
  return p->nonlinear && !p->passing_v && p->D % 128 == 0 && p->D >= 128 && p->D <= kMaxD && p->C <= 4 &&
         pair_smem_bytes(p->C, p->D) + 5632 <= 232448;
}

// CT = classes rounded up to 1/2/4; NB = boxes per tile (D / 32) or 0 = run-time
template <int CT, int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_fwd_pair(const PairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  enum { XS_FULL = 0, XS_EMPTY = XS_FULL + kXBars, XT_FULL = XS_EMPTY + kXBars, XT_EMPTY = XT_FULL + kTStages,
         H1_FULL = XT_EMPTY + kTStages, Q_FULL = H1_FULL + 2, ACC_EMPTY = Q_FULL + 2, A2_FULL = ACC_EMPTY + 2, A2_EMPTY,
         SC_DONE, W_BAR = SC_DONE + 2, W_READY, NBARS };
  __shared__ __align__(8) uint64_t bars[NBARS];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_b1[kQ], s_b2[kQ];
  __shared__ BagDev s_bags[kSmemBags];
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cta_rank();
  const int D = a.D, C = a.C;
  const int nchunks = D >> 6;
  const int nboxes = NB ? NB : (D >> 5);
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
  const int nsuper = (a.ntiles + 1) >> 1;
  const int niter = cl < nsuper ? (nsuper - cl + ncl - 1) / ncl : 0;          // 256-row super-tiles of this pair
  const size_t wbytes = pair_wimg_rank_bytes(D);
  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + static_cast<size_t>(nchunks) * kWChunk;
  uint8_t* sX = smem + wbytes;
  float* sWi = reinterpret_cast<float*>(sX + kSlots * kBoxBytes);
  float* sSc = sWi + CT * D;                       // [tile parity 2][partial 4 = group x k-half][128 rows][CT]
  const bool do_scores = a.classes != nullptr;

  if (do_scores)
    for (int i = tid; i < CT * D; i += kThreads) sWi[i] = (i < C * D) ? a.Wi[i] : 0.f;
  if (tid < kQ) { s_b1[tid] = a.b1[tid]; s_b2[tid] = a.b2[tid]; }
  const bool tbl_in_smem = a.nb <= kSmemBags;
  if (tbl_in_smem)
    for (int i = tid; i < a.nb; i += kThreads) s_bags[i] = a.bags[i];
  const BagDev* tbl = tbl_in_smem ? s_bags : a.bags;
  if (tid == 0) {
    for (int s = 0; s < kXBars; ++s) { mbar_init(bar(XS_FULL + s), 1); mbar_init(bar(XS_EMPTY + s), 4); }
    for (int s = 0; s < kTStages; ++s) { mbar_init(bar(XT_FULL + s), 8); mbar_init(bar(XT_EMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(H1_FULL + b), 1); mbar_init(bar(Q_FULL + b), 1); mbar_init(bar(ACC_EMPTY + b), 2 * kEpiWarps);
      mbar_init(bar(SC_DONE + b), kConvWarps);
    }
    mbar_init(bar(A2_FULL), 2 * kEpiWarps); mbar_init(bar(A2_EMPTY), 1);
    mbar_init(bar(W_BAR), 1); mbar_init(bar(W_READY), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpAux) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // both CTAs' barriers initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  const uint32_t tm_a2hi = tmem + 256, tm_a2lo = tmem + 320, tm_x = tmem + 384;

  if (warp == kWarpAux) {
    // ===== resident weight half of this CTA (bulk copies), then report "weights resident" to the leader =====
    if (lane == 0) {
      mbar_expect_tx(bar(W_BAR), static_cast<uint32_t>(wbytes));
      const uint8_t* src = a.wimg + static_cast<size_t>(rank) * wbytes;
      for (int c = 0; c < nchunks + 2; ++c)
        bulk_g2s(smem_u32(smem + static_cast<size_t>(c) * kWChunk), src + static_cast<size_t>(c) * kWChunk, kWChunk, bar(W_BAR));
      mbar_wait<false>(bar(W_BAR), 0, 64);
      mbar_arrive_leader(bar(W_READY), rank);
    }
  } else if (warp == kWarpProd) {
    // ===== TMA producer: the boxes of this CTA's tiles in order; box n -> slot n % 3, barriers n % 6.  One L2 prefetch
    // per real load, one tile ahead (the few boxes in flight then pay L2 latency, not HBM latency) =====
    if (lane == 0) {
      TileCursor cur(tbl, 0, a.nb), cur_pf(tbl, 0, a.nb);
      const uint64_t pol_keep = sm100::l2_policy_evict_last();
      uint32_t n = 0;
      for (int j = 0; j < niter; ++j) {
        const int tile = 2 * (cl + j * ncl) + static_cast<int>(rank);
        if (tile >= a.ntiles) break;               // odd tile count: the second CTA of the last pair has no tile
        cur.seek(tile);
        const BagDev* bp = tbl + cur.bag;
        const CUtensorMap* tm = a.tmaps + cur.bag;
        const int row0 = (tile - bp->tile_off) * kTileM;
        const int ptile = 2 * (cl + (j + 1) * ncl) + static_cast<int>(rank);
        const bool pf = j + 1 < niter && ptile < a.ntiles;
        const uint8_t* pf_ptr = nullptr;             // next tile of this CTA as a contiguous byte range
        uint32_t pf_bytes = 0;
        if (pf) {
          cur_pf.seek(ptile);
          const BagDev* pb = tbl + cur_pf.bag;
          const long long prow0 = static_cast<long long>(ptile - pb->tile_off) * kTileM;
          const long long prows = (pb->N - prow0) < kTileM ? (pb->N - prow0) : kTileM;
          pf_ptr = reinterpret_cast<const uint8_t*>(pb->X + prow0 * D);
          pf_bytes = static_cast<uint32_t>(prows * D * 4);
        }
        const uint32_t pf_step = static_cast<uint32_t>(kTileM * D * 4) / nboxes;     // 16 KB at D = 512
        for (int kb = 0; kb < nboxes; ++kb, ++n) {
          if (n >= kSlots) {                         // slot last used by box n-3: wait until its group has read it
            const uint32_t pn = n - kSlots;
            mbar_wait<false>(bar(XS_EMPTY + pn % kXBars), (pn / kXBars) & 1, (a.flags & 4) ? 0 : 32);
          }
          if (blockIdx.x == 0) PAIR_TRACE(3, 0, n);
          mbar_expect_tx(bar(XS_FULL + n % kXBars), kBoxBytes);
          tma_load_2d(smem_u32(sX + (n % kSlots) * kBoxBytes), tm, kb * kBoxK, row0, bar(XS_FULL + n % kXBars), pol_keep);
          if (pf && !(a.flags & 1) && kb * pf_step < pf_bytes) {
            const uint32_t off = kb * pf_step;
            bulk_prefetch_l2(pf_ptr + off, (pf_bytes - off) < pf_step ? (pf_bytes - off) : pf_step);
          }
          if (j == 0 && kb == kSlots - 1 && !(a.flags & 1)) {          // the first tile itself: start the rest of it towards L2 now
            const long long rws = (bp->N - row0) < kTileM ? (bp->N - row0) : kTileM;
            const uint8_t* t0 = reinterpret_cast<const uint8_t*>(bp->X + static_cast<long long>(row0) * D);
            const uint32_t tb = static_cast<uint32_t>(rws * D * 4);
            for (uint32_t off = 0; off < tb; off += 32768) bulk_prefetch_l2(t0 + off, (tb - off) < 32768u ? (tb - off) : 32768u);
          }
        }
      }
    }
  } else if (warp >= kWarpConv0 && warp < kWarpMma) {
    // ===== converters: lane = row of the quadrant; warp (g, q): group g takes boxes g, g+4, ...; TMEM stage g =====
    const int cw = warp - kWarpConv0, q = cw & 3, g = cw >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sx_u32 = smem_u32(sX) + row * 128;
    const uint32_t swi_u32 = smem_u32(sWi);
    const uint32_t total = static_cast<uint32_t>(niter) * nboxes;
    const uint32_t tstage = tm_x + lane_sel + g * 32;
    float sc[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) sc[k] = 0.f;
    int j = 0, kb = g;                              // (tile iteration, box in tile) of box n; nboxes % 4 == 0
    bool valid = 2 * cl + static_cast<int>(rank) < a.ntiles;
    const int trole = (lane == 0 && blockIdx.x < 2 && (cw == 0 || cw == 15)) ? (cw == 0 ? (blockIdx.x ? 4 : 0) : (blockIdx.x ? 6 : 5)) : -1;
    uint32_t m = 0;
    for (uint32_t n = g; n < total; n += kGroups, ++m) {
      const uint32_t fb = n % kXBars;
      if (valid) {
        if (trole >= 0) PAIR_TRACE(trole, 0, m);
        mbar_wait<false>(bar(XS_FULL + fb), (n / kXBars) & 1);
        if (trole >= 0) PAIR_TRACE(trole, 1, m);
        PAIR_TRACE_ALL(0, n, cw);
      }
      const uint32_t rowb = sx_u32 + (n % kSlots) * kBoxBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {                 // two halves of 16 k: 8 + 8 TMEM columns each
        uint32_t hi[8], lo[8];
        if (valid) {
          float4 x[4];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii)              // the four loads are issued back to back, then used
            x[ii] = lds128(rowb + ((((4 * h + ii) ^ (row & 7)) & 7) << 4));        // floats 4i .. 4i+3, i = 4h+ii
          if (do_scores) {
#pragma unroll
            for (int k = 0; k < CT; ++k) {
              float4 w[4];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii)
                w[ii] = lds128(swi_u32 + static_cast<uint32_t>(k * D + kb * kBoxK + 16 * h + 4 * ii) * 4u);   // broadcast
              float sv = sc[k];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                sv = fmaf(x[ii].x, w[ii].x, sv); sv = fmaf(x[ii].y, w[ii].y, sv);
                sv = fmaf(x[ii].z, w[ii].z, sv); sv = fmaf(x[ii].w, w[ii].w, sv);
              }
              sc[k] = sv;
            }
          }
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const float4 xv = x[ii];
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(xv.x, xv.y), h23 = __floats2bfloat162_rn(xv.z, xv.w);
            const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(xv.x - __uint_as_float(u01 << 16), xv.y - __uint_as_float(u01 & 0xffff0000u));
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(xv.z - __uint_as_float(u23 << 16), xv.w - __uint_as_float(u23 & 0xffff0000u));
            hi[2 * ii] = u01; hi[2 * ii + 1] = u23;
            lo[2 * ii] = *reinterpret_cast<const uint32_t*>(&l01); lo[2 * ii + 1] = *reinterpret_cast<const uint32_t*>(&l23);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { hi[i] = 0u; lo[i] = 0u; }
        }
        if (h == 0) {
          if (trole >= 0) PAIR_TRACE(trole, 2, m);
          mbar_wait<false>(bar(XT_EMPTY + g), (m & 1) ^ 1, 32);           // operand stage released by the MMAs of box n-4
          if (trole >= 0) PAIR_TRACE(trole, 3, m);
          PAIR_TRACE_ALL(1, n, cw);
          tc_fence_after();
        } else if (valid) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(XS_EMPTY + fb));                  // staging slot read by this warp
        }
        DSMIL_TMEM_ST8(tstage + h * 8, hi);
        DSMIL_TMEM_ST8(tstage + 16 + h * 8, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar(XT_FULL + g), rank);
      if (trole >= 0) PAIR_TRACE(trole, 4, m);
      PAIR_TRACE_ALL(2, n, cw);
      // next box of this group; on leaving a tile hand the score partial of group g to the epilogue
      const int kb_next = kb + kGroups;
      if (kb_next >= nboxes) {
        if (do_scores) {
          float* dst = sSc + ((static_cast<size_t>(j & 1) * 4 + g) * kTileM + row) * CT;
#pragma unroll
          for (int k = 0; k < CT; ++k) { dst[k] = sc[k]; sc[k] = 0.f; }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SC_DONE + (j & 1)));
        }
        kb = kb_next - nboxes;
        ++j;
        valid = 2 * (cl + j * ncl) + static_cast<int>(rank) < a.ntiles;
      } else {
        kb = kb_next;
      }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (leader CTA): layer 1 box by box, layer 2 of the previous tile in the middle of a tile =====
    if (rank == 0 && lane == 0) {
      mbar_wait<true>(bar(W_READY), 0, 64);
      tc_fence_after();
      const uint32_t w1d = sm100::desc_lo(smem_u32(sW1)), w2d = sm100::desc_lo(smem_u32(sW2));
      auto issue_l2 = [&](int jj) {
        PAIR_TRACE(1, 2, jj);
        mbar_wait<true>(bar(A2_FULL), jj & 1);
        PAIR_TRACE(1, 3, jj);
        tc_fence_after();
        const uint32_t dacc = tmem + (jj & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t bhi = w2d + (((ks >> 2) * kWChunk + (ks & 3) * 32) >> 4), blo = bhi + ((64 * 128) >> 4);
          mma2_ts(dacc, tm_a2hi + ks * 8, bhi, ks > 0);
          mma2_ts(dacc, tm_a2lo + ks * 8, bhi, 1);
          mma2_ts(dacc, tm_a2hi + ks * 8, blo, 1);
        }
        tc_commit_pair(bar(Q_FULL + (jj & 1)));
        tc_commit_pair(bar(A2_EMPTY));
      };
      uint32_t n = 0;
      for (int j = 0; j < niter; ++j) {
        const int b = j & 1;
        mbar_wait<true>(bar(ACC_EMPTY + b), ((j >> 1) & 1) ^ 1);
        PAIR_TRACE(1, 4, j);
        tc_fence_after();
        const uint32_t dacc = tmem + b * 128;
        for (int kb = 0; kb < nboxes; ++kb, ++n) {
          const uint32_t g = n % kTStages, u = n / kTStages;      // TMEM operand stage of box n
          mbar_wait<true>(bar(XT_FULL + g), u & 1);
          PAIR_TRACE(1, 0, n);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int kabs = kb * kBoxK + ks * 16;
            const uint32_t bhi = w1d + (((kabs >> 6) * kWChunk + (kabs & 63) * 2) >> 4), blo = bhi + ((64 * 128) >> 4);
            const uint32_t ahi = tm_x + g * 32 + ks * 8, alo = ahi + 16;
            mma2_ts(dacc, ahi, bhi, (kb | ks) != 0);
            mma2_ts(dacc, alo, bhi, 1);
            mma2_ts(dacc, ahi, blo, 1);
          }
          tc_commit_pair(bar(XT_EMPTY + g));
          PAIR_TRACE(1, 1, n);
          if (j > 0 && kb == (nboxes >> 1) - 1) issue_l2(j - 1);
        }
        tc_commit_pair(bar(H1_FULL + b));
      }
      if (niter > 0) issue_l2(niter - 1);
    }
  } else if (warp < kEpiWarps) {
    // ===== epilogue: scores / keys, H1 -> A2 (TMEM), Q = tanh(.) -> global (tile-blocked) =====
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row_in_tile = (warp & 3) * 32 + lane;
    const int half = warp >> 2, col0 = half * 64;
    TileCursor cur(tbl, 0, a.nb);
    for (int j = 0; j < niter; ++j) {
      const int b = j & 1, ub = (j >> 1) & 1;
      const int tile = 2 * (cl + j * ncl) + static_cast<int>(rank);
      const bool valid = tile < a.ntiles;
      BagDev bg = tbl[0];
      long long nrow = 0;
      bool live = false;
      if (valid) {
        cur.seek(tile);
        bg = tbl[cur.bag];
        nrow = static_cast<long long>(tile - bg.tile_off) * kTileM + row_in_tile;
        live = nrow < bg.N;
      }
      // ---- instance scores of this tile: four partials (boxes kb = g mod 4) summed in a fixed order + bias ----
      if (do_scores && valid) {
        mbar_wait<false>(bar(SC_DONE + b), ub, 64);
        if (half == 0) {
          unsigned long long best[CT];
          const float* p0 = sSc + (static_cast<size_t>(b) * 4 * kTileM + row_in_tile) * CT;
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            best[k] = 0ull;
            if (k < C && live) {
              const float v = (((p0[k] + p0[kTileM * CT + k]) + p0[2 * kTileM * CT + k]) + p0[3 * kTileM * CT + k]) + a.bi[k];
              a.classes[(bg.row_off + nrow) * C + k] = v;
              best[k] = pack_key(v, static_cast<uint32_t>(nrow));
            }
          }
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            const unsigned long long bb = warp_max_u64(best[k]);
            if (lane == 0 && k < C && bb) atomicMax(a.keys + static_cast<size_t>(cur.bag) * kMaxC + k, bb);
          }
        }
      }
      // ---- H1 = relu(acc + b1) -> bf16 hi/lo -> A operand of layer 2 in TMEM ----
      if (tid == 0) PAIR_TRACE(blockIdx.x ? 7 : 2, 4, j);
      mbar_wait<false>(bar(H1_FULL + b), ub, 128);
      if (tid == 0) PAIR_TRACE(blockIdx.x ? 7 : 2, 0, j);
      mbar_wait<false>(bar(A2_EMPTY), (j & 1) ^ 1, 64);        // layer 2 of the previous tile has consumed A2
      tc_fence_after();
      const uint32_t tacc = tmem + b * 128;
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + 64; c0 += 16) {
        uint32_t v[16];
        DSMIL_TMEM_LD16(tacc + lane_sel + c0, v);
        tmem_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const float2 bb = *reinterpret_cast<const float2*>(&s_b1[c0 + 2 * qq]);
          const f2 z = add2(f2{__uint_as_float(v[2 * qq]), __uint_as_float(v[2 * qq + 1])}, f2{bb.x, bb.y});
          const float h0 = fmaxf(z.x, 0.f), h1 = fmaxf(z.y, 0.f);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
          const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh);
          const f2 res = add2(f2{h0, h1}, f2{-__uint_as_float(hu << 16), -__uint_as_float(hu & 0xffff0000u)});
          const __nv_bfloat162 ll = __floats2bfloat162_rn(res.x, res.y);
          hi[qq] = hu;
          lo[qq] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        DSMIL_TMEM_ST8(tm_a2hi + lane_sel + (c0 >> 1), hi);
        DSMIL_TMEM_ST8(tm_a2lo + lane_sel + (c0 >> 1), lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar(A2_FULL), rank);
      if (tid == 0) PAIR_TRACE(blockIdx.x ? 7 : 2, 1, j);
      // ---- Q = tanh(acc + b2): the layer-2 result arrives in the SAME accumulator in the middle of the next tile ----
      mbar_wait<false>(bar(Q_FULL + b), ub, 128);
      if (tid == 0) PAIR_TRACE(blockIdx.x ? 7 : 2, 2, j);
      tc_fence_after();
      float* qdst = a.Q + static_cast<size_t>(valid ? tile : 0) * (kTileM * kQ) + row_in_tile;
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + 64; c0 += 16) {
        uint32_t v[16];
        DSMIL_TMEM_LD16(tacc + lane_sel + c0, v);
        tmem_wait_ld();
        if (c0 == col0 + 48) {                               // accumulator drained by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(bar(ACC_EMPTY + b), rank);
        }
        if (valid) {
          float* dst = qdst + static_cast<size_t>(c0) * kTileM;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float4 bb = *reinterpret_cast<const float4*>(&s_b2[c0 + 4 * qq]);
            const f2 t0 = fast_tanh2(add2(f2{__uint_as_float(v[4 * qq + 0]), __uint_as_float(v[4 * qq + 1])}, f2{bb.x, bb.y}));
            const f2 t1 = fast_tanh2(add2(f2{__uint_as_float(v[4 * qq + 2]), __uint_as_float(v[4 * qq + 3])}, f2{bb.z, bb.w}));
            dst[(4 * qq + 0) * kTileM] = t0.x;
            dst[(4 * qq + 1) * kTileM] = t0.y;
            dst[(4 * qq + 2) * kTileM] = t1.x;
            dst[(4 * qq + 3) * kTileM] = t1.y;
          }
        }
      }
      if (tid == 0) PAIR_TRACE(blockIdx.x ? 7 : 2, 3, j);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // the peer's barriers / TMEM / weights stay alive until both are done
  if (warp == kWarpAux) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Weight images for the pair kernel: per CTA rank r the 64 output features [64r, 64r+64) of W1 (D/64 chunks) and W2
// (2 chunks); per chunk a hi tile and a lo tile [64 x 64] bf16, K-major, SWIZZLE_128B -- byte for byte the smem operand.
__global__ void __launch_bounds__(256)
k_prep_wimg_pair(const float* __restrict__ W1, int D, const float* __restrict__ W2, uint8_t* __restrict__ img) {
  const int t1 = 128 * D, total = t1 + 128 * kQ;
  const size_t rank_bytes = static_cast<size_t>(D / 64 + 2) * kWChunk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool first = i < t1;
    const int K = first ? D : kQ;
    const int jj = first ? i : i - t1;
    const int n = jj / K, k = jj % K;
    const float w = first ? W1[jj] : W2[jj];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    uint8_t* chunk = img + static_cast<size_t>(n >> 6) * rank_bytes +
                     static_cast<size_t>((first ? 0 : D / 64) + (k >> 6)) * kWChunk;
    const uint32_t off = swz_off(n & 63, k & 63);
    *reinterpret_cast<__nv_bfloat16*>(chunk + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(chunk + 64 * 128 + off) = lo;
  }
}

inline int launch_prep_wimg_pair(const dsmil_params_t* p, uint8_t* img, cudaStream_t st) {
  k_prep_wimg_pair<<<80, 256, 0, st>>>(p->W1, p->D, p->W2, img);
  DSMIL_LAUNCH_OK("k_prep_wimg_pair");
  return 0;
}

// Host: tensor map of one bag (X [N, D] fp32 row-major), box {32 floats, 128 rows}, 128-byte swizzle, OOB rows -> 0.
inline int encode_bag_tmap(CUtensorMap* tm, const float* X, long long N, int D) {
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(D) * sizeof(float)};
  const cuuint32_t box[2] = {kBoxK, kTileM};
  const cuuint32_t estr[2] = {1, 1};
  // resolved through the runtime (no link-time dependency on libcuda: the library must load on a machine without a
  // driver, where only the host-side logic is exercised)
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from the installed driver");
      return DSMIL_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%lld, %d] bag", static_cast<int>(r), N, D);
    return DSMIL_ERR_CUDA;
  }
  return 0;
}

// scores + arg-max keys + tile-blocked Q for every tile of every bag of the table (inference path).
inline int launch_fwd_pair(const dsmil_params_t* p, const BagDev* bags_dev, const CUtensorMap* tmaps_dev, int nb, int ntiles,
                           float* classes, unsigned long long* keys, float* Q, const uint8_t* wimg, int num_sms,
                           cudaStream_t st) {
  const int D = p->D, C = p->C;
  static int flags = -1;
  if (flags < 0) { const char* e = getenv("DSMIL_B200_PAIR_FLAGS"); flags = e ? atoi(e) : 0; }
  PairArgs a{bags_dev, tmaps_dev, nb, ntiles, D, C, p->Wi, p->bi, p->b1, p->b2, wimg, classes, keys, Q, sm100::g_trace_buf, flags};
  const size_t smem = pair_smem_bytes(C, D);
  const int nsuper = (ntiles + 1) / 2;
  const int pairs = std::max(1, std::min(nsuper, num_sms / 2));
  auto go = [&](auto kern) -> int {
    DSMIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    prof_begin(PROF_FUSED, st);
    kern<<<2 * pairs, kThreads, smem, st>>>(a);
    prof_end(PROF_FUSED, st);
    DSMIL_LAUNCH_OK("k_fwd_pair");
    return 0;
  };
  if (D == 512) {
    if (C == 1) return go(k_fwd_pair<1, 16>);
    if (C == 2) return go(k_fwd_pair<2, 16>);
    return go(k_fwd_pair<4, 16>);
  }
  if (C == 1) return go(k_fwd_pair<1, 0>);
  if (C == 2) return go(k_fwd_pair<2, 0>);
  return go(k_fwd_pair<4, 0>);
}

}  // namespace pair
}  // namespace dsmil
