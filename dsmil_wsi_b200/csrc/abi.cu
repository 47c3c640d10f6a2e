// extern "C" entry points of libdsmil_b200.so (see include/dsmil_b200.h for the contract and the
// reference spans each call replaces).  Host orchestration only; kernels live in *_kernels.cuh
// and fwd_sm100.cuh.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "common.cuh"
#include "gemm_generic.cuh"
#include "fwd_kernels.cuh"
#include "bwd_kernels.cuh"
#include "fwd_sm100.cuh"
#include "fwd_batched.cuh"
#include "fwd_pair.cuh"
#include "embed_kernels.cuh"
#include "jpeg_kernels.cuh"

namespace dsmil {

static thread_local char g_err[512] = "";
__device__ unsigned long long scratch_keys[kMaxC];  // sink for k_scores' arg-max by-product
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return DSMIL_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

namespace sm100 { long long* g_trace_buf = nullptr; }

// train_tcga.py:78-83 dropout_patches: `feats[random_indices]` -- a row gather.  One warp per output row.
__global__ void __launch_bounds__(256)
k_gather_rows(const float* __restrict__ X, int D, const long long* __restrict__ idx, long long M,
              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long stride = static_cast<long long>(gridDim.x) * 8;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (long long m = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); m < M; m += stride) {
    const float* src = X + idx[m] * D;
    float* dst = out + m * D;
    if (vec) {
      for (int j = lane; j < (D >> 2); j += 32)
        reinterpret_cast<float4*>(dst)[j] = __ldg(reinterpret_cast<const float4*>(src) + j);
    } else {
      for (int j = lane; j < D; j += 32) dst[j] = __ldg(src + j);
    }
  }
}



// compute_feats.py:19-46 (PIL -> VF.to_tensor): uint8 HWC -> float32 CHW, value / 255 (an IEEE division, as
// torchvision's `img.div(255)`), done on the device so that patches cross PCIe as bytes (4x less H2D traffic).
__global__ void __launch_bounds__(256)
k_u8hwc_to_f32chw(const uint8_t* __restrict__ in, long long B, int H, int W, int Cc, float* __restrict__ out) {
  const long long plane = static_cast<long long>(H) * W;
  const long long total = B * plane;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i / plane, px = i % plane;
    const uint8_t* src = in + i * Cc;
    float* dst = out + b * Cc * plane + px;
    for (int c = 0; c < Cc; ++c) dst[c * plane] = __fdiv_rn(static_cast<float>(src[c]), 255.f);
  }
}

// ---- live kernel timing ------------------------------------------------------------------
bool g_prof_on = false;
struct ProfRec { int tag; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;       // recorded pairs since the last read
static std::vector<ProfRec> g_prof_pool;  // recycled events
static int g_prof_open[PROF_NTAGS];
void prof_begin_impl(int tag, cudaStream_t st) {
  if (g_prof.size() >= 16384) { g_prof_open[tag] = -1; return; }
  ProfRec r;
  if (!g_prof_pool.empty()) { r = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else { if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { g_prof_open[tag] = -1; return; } }
  r.tag = tag;
  cudaEventRecord(r.a, st);
  g_prof_open[tag] = static_cast<int>(g_prof.size());
  g_prof.push_back(r);
}
void prof_end_impl(int tag, cudaStream_t st) {
  const int i = g_prof_open[tag];
  if (i >= 0 && i < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[i].b, st);
}

static int check_params(const dsmil_params_t* p, bool need_scores = false) {
  DSMIL_REQUIRE(p != nullptr, "params is NULL");
  DSMIL_REQUIRE(!need_scores || (p->Wi && p->bi), "NULL instance-classifier weights");
  DSMIL_REQUIRE(p->D >= 1 && p->D <= DSMIL_MAX_D, "feature size D=%d outside [1,%d]", p->D, DSMIL_MAX_D);
  DSMIL_REQUIRE(p->C >= 1 && p->C <= DSMIL_MAX_C, "output classes C=%d outside [1,%d]", p->C, DSMIL_MAX_C);
  DSMIL_REQUIRE(p->W1 && p->b1 && p->Wf && p->bf, "NULL weight pointer");
  DSMIL_REQUIRE(!p->nonlinear || (p->W2 && p->b2), "nonlinear q needs W2/b2");
  DSMIL_REQUIRE(!p->passing_v || (p->Wv && p->bv), "passing_v needs Wv/bv");
  return 0;
}

static inline int attend_ctas(int64_t N) {
  const int64_t tiles = (N + kAttendRows - 1) / kAttendRows;
  return static_cast<int>(tiles < 296 ? (tiles < 1 ? 1 : tiles) : 296);
}

struct FwdWs {
  unsigned long long* keys;
  float *Q, *H1, *V, *cand, *qmax, *recs, *rec;
  int64_t* crit;
  uint8_t* wimg;
  size_t bytes;
};
static FwdWs carve_fwd(const dsmil_params_t* p, int64_t N, void* ws, size_t cap, bool* ok) {
  Carver c(ws, cap);
  FwdWs w;
  const int64_t n = N > 0 ? N : 1;
  w.keys = c.take<unsigned long long>(kMaxC);
  w.Q = c.take<float>(n * kQ);
  w.H1 = p->nonlinear ? c.take<float>(n * kQ) : nullptr;
  w.V = p->passing_v ? c.take<float>(n * p->D) : nullptr;
  w.cand = c.take<float>(cand_floats(p->C));
  w.qmax = c.take<float>(static_cast<size_t>(p->C) * kQ);
  w.crit = c.take<int64_t>(kMaxC);
  w.recs = c.take<float>(static_cast<size_t>(attend_ctas(N)) * rec_floats(p->C, p->D));
  w.rec = c.take<float>(rec_floats(p->C, p->D));
  w.wimg = sm100::qmlp_supported(p) ? c.take<uint8_t>(sm100::wimg_bytes(p->D) + 1024 + 256) : nullptr;
  w.bytes = c.off;
  *ok = c.ok();
  return w;
}

// ---- phase 1: scores + arg-max key + Q-MLP (+V) + candidate record --------------------------

static int launch_scores(const dsmil_params_t* p, const float* X, int64_t N, float* classes,
                         unsigned long long* keys, cudaStream_t st) {
  const int C = p->C, D = p->D;
  const size_t smem = sizeof(float) * C * D;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const int mode = !vec ? 0 : (D % 64 == 0 ? 2 : 1);
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N, mode == 2 ? 16 : 8), 148 * 8));
  prof_begin(PROF_SCORES, st);
  if (mode == 2) {
    if (smem > 48 * 1024) DSMIL_CUDA_OK(cudaFuncSetAttribute(k_scores<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_scores<2><<<grid, 256, smem, st>>>(X, N, D, p->Wi, p->bi, C, classes, keys);
  } else if (mode == 1) {
    if (smem > 48 * 1024) DSMIL_CUDA_OK(cudaFuncSetAttribute(k_scores<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_scores<1><<<grid, 256, smem, st>>>(X, N, D, p->Wi, p->bi, C, classes, keys);
  } else {
    if (smem > 48 * 1024) DSMIL_CUDA_OK(cudaFuncSetAttribute(k_scores<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_scores<0><<<grid, 256, smem, st>>>(X, N, D, p->Wi, p->bi, C, classes, keys);
  }
  prof_end(PROF_SCORES, st);
  DSMIL_LAUNCH_OK("k_scores");
  return 0;
}

static int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
static bool use_sm100(const dsmil_params_t* p) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("DSMIL_B200_GENERIC"); disabled = (e && e[0] == '1') ? 1 : 0; }
  return !disabled && sm100::qmlp_supported(p);
}

// Under stream capture (CUDA-graph serving loops) the pageable host->device copies of the bag table cannot be
// recorded; the captured call then reuses the table that the preceding EAGER call with the same arguments wrote into
// the same workspace (dsmil_wsi_b200.sharded.ShardedBagsGraph does exactly that: warm-up run, then capture).
static bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

// tile-blocked Q holds the pre-activation, tanh applied by its readers (DSMIL_B200_QPRE=0: r1 behaviour, tanh in phase 1)
static int blocked_q_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("DSMIL_B200_QPRE"); mode = (e && e[0] == '0') ? 1 : 2; }
  return mode;
}
static bool use_pair(const dsmil_params_t* p) {
  // CTA-pair phase 1 (fwd_pair.cuh): parity-green, but not yet faster than k_qmlp_sm100 (profiles/r2_bench_history.md),
  // so it is opt-in: DSMIL_B200_PAIR=1
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DSMIL_B200_PAIR"); enabled = (e && e[0] == '1') ? 1 : 0; }
  return enabled && use_sm100(p) && pair::pair_supported(p);
}

static int phase1_impl(const dsmil_params_t* p, const float* X, const float* xv, const float* classes_in,
                       int64_t N, int64_t row_offset, float* classes, float* Q, float* H1, float* V,
                       float* cand, unsigned long long* keys, uint8_t* wimg, cudaStream_t st) {
  const int C = p->C, D = p->D;
  DSMIL_CUDA_OK(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * kMaxC, st));
  if (N > 0 && use_sm100(p) && wimg && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    // tensor-core path: scores + arg-max + Q-MLP in one persistent kernel
    int rc;
    if (classes_in) {
      if (classes && classes != classes_in)
        DSMIL_CUDA_OK(cudaMemcpyAsync(classes, classes_in, sizeof(float) * N * C, cudaMemcpyDeviceToDevice, st));
      const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N, 256), 296));
      k_argmax<<<grid, 256, 0, st>>>(classes_in, N, C, keys);
      DSMIL_LAUNCH_OK("k_argmax");
    }
    uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wimg) + 1023) & ~uintptr_t(1023));
    sm100::BagDev* tbl = reinterpret_cast<sm100::BagDev*>(img + sm100::wimg_bytes(D));
    sm100::BagDev one{X, N, 0, 0, 0, 1, 0};
    DSMIL_CUDA_OK(cudaMemcpyAsync(tbl, &one, sizeof(one), cudaMemcpyHostToDevice, st));
    if ((rc = sm100::launch_prep_wimg(p, img, st))) return rc;
    const int ntiles = static_cast<int>((N + sm100::kTileM - 1) / sm100::kTileM);
    if ((rc = sm100::launch_qmlp(p, tbl, 0, 1, 0, ntiles, classes_in ? nullptr : classes, keys, Q, H1, img, num_sms(), st)))
      return rc;
    if (p->passing_v) {
      if ((rc = launch_linear<ACT_RELU, false>(xv ? xv : X, N, D, p->Wv, p->bv, D, V, nullptr, 0, st))) return rc;
    }
  } else if (N > 0) {
    if (classes_in) {
      if (classes && classes != classes_in)
        DSMIL_CUDA_OK(cudaMemcpyAsync(classes, classes_in, sizeof(float) * N * C, cudaMemcpyDeviceToDevice, st));
      const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N, 256), 296));
      k_argmax<<<grid, 256, 0, st>>>(classes_in, N, C, keys);
      DSMIL_LAUNCH_OK("k_argmax");
    } else {
      int rcs = launch_scores(p, X, N, classes, keys, st);
      if (rcs) return rcs;
    }
    int rc;
    prof_begin(PROF_QMLP, st);
    if (p->nonlinear) {
      if ((rc = launch_linear<ACT_RELU, false>(X, N, D, p->W1, p->b1, kQ, H1, nullptr, 0, st))) return rc;
      if ((rc = launch_linear<ACT_TANH, false>(H1, N, kQ, p->W2, p->b2, kQ, Q, nullptr, 0, st))) return rc;
    } else {
      if ((rc = launch_linear<ACT_NONE, false>(X, N, D, p->W1, p->b1, kQ, Q, nullptr, 0, st))) return rc;
    }
    prof_end(PROF_QMLP, st);
    if (p->passing_v) {
      if ((rc = launch_linear<ACT_RELU, false>(xv ? xv : X, N, D, p->Wv, p->bv, D, V, nullptr, 0, st))) return rc;
    }
  }
  const float* cls = classes_in ? classes_in : classes;
  k_gather_cand<<<C, kQ, 0, st>>>(keys, cls, Q, N, C, row_offset, cand);
  DSMIL_LAUNCH_OK("k_gather_cand");
  return 0;
}

__global__ void k_empty_rec(float* rec, int C, int Dv) {
  const size_t n = rec_floats(C, Dv);
  for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    rec[i] = (i < static_cast<size_t>(C)) ? -INFINITY : 0.f;
}

template <int CT>
static int launch_attend_j(int J, int grid, cudaStream_t st, const float* V, int Dv, const float* Q, int64_t N,
                           const float* qmax, int C, float* A, float* recs) {
  switch (J) {
    case 1: k_attend<CT, 1><<<grid, 256, 0, st>>>(V, Dv, Q, N, qmax, C, A, recs); break;
    case 2: k_attend<CT, 2><<<grid, 256, 0, st>>>(V, Dv, Q, N, qmax, C, A, recs); break;
    case 4: k_attend<CT, 4><<<grid, 256, 0, st>>>(V, Dv, Q, N, qmax, C, A, recs); break;
    case 8: k_attend<CT, 8><<<grid, 256, 0, st>>>(V, Dv, Q, N, qmax, C, A, recs); break;
    default: k_attend<CT, 16><<<grid, 256, 0, st>>>(V, Dv, Q, N, qmax, C, A, recs); break;
  }
  DSMIL_LAUNCH_OK("k_attend");
  return 0;
}

// ---- phase 2: logits -> A (unnormalised), device-level (m, s, Bp) record ------------------------
static int phase2_impl(const dsmil_params_t* p, const float* V, const float* Q, int64_t N, const float* qmax,
                       float* A, float* rec, float* recs, cudaStream_t st) {
  const int C = p->C, Dv = p->D;
  if (N <= 0) {
    k_empty_rec<<<4, 256, 0, st>>>(rec, C, Dv);
    DSMIL_LAUNCH_OK("k_empty_rec");
    return 0;
  }
  const int grid = attend_ctas(N);
  int j = ceil_div(Dv, 256), J = 1;
  while (J < j) J <<= 1;
  int rc;
  prof_begin(PROF_ATTEND, st);
  if (C == 1) rc = launch_attend_j<1>(J, grid, st, V, Dv, Q, N, qmax, C, A, recs);
  else if (C == 2) rc = launch_attend_j<2>(J, grid, st, V, Dv, Q, N, qmax, C, A, recs);
  else if (C <= 4) rc = launch_attend_j<4>(J, grid, st, V, Dv, Q, N, qmax, C, A, recs);
  else rc = launch_attend_j<8>(J, grid, st, V, Dv, Q, N, qmax, C, A, recs);
  prof_end(PROF_ATTEND, st);
  if (rc) return rc;
  dim3 g2(C, ceil_div(Dv, 256));
  k_combine_rec<<<g2, 256, 0, st>>>(recs, grid, C, Dv, rec);
  DSMIL_LAUNCH_OK("k_combine_rec");
  return 0;
}

static int phase3_impl(const dsmil_params_t* p, int64_t N, const float* rec, float* A, float* B, float* pred,
                       cudaStream_t st) {
  const int grid = static_cast<int>(std::min<int64_t>(std::max<int64_t>(ceil_div(N * p->C, 256), 1), 296));
  prof_begin(PROF_FINAL, st);
  k_finalize<<<grid, 256, 0, st>>>(rec, N, p->C, p->D, p->Wf, p->bf, A, B, pred);
  prof_end(PROF_FINAL, st);
  DSMIL_LAUNCH_OK("k_finalize");
  return 0;
}

static int forward_bags_impl(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int nb,
                             const float* classes_in, float* classes, float* pred, float* A, float* B,
                             int64_t* crit, float* save_Q, float* save_H1, void* ws, size_t ws_bytes,
                             cudaStream_t st);

static int forward_impl(const dsmil_params_t* p, const float* X, const float* xv, const float* classes_in,
                        int64_t N, float* classes, float* pred, float* A, float* B, int64_t* crit_idx,
                        float* save_Q, float* save_H1, float* save_V, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = check_params(p, classes_in == nullptr);
  if (rc) return rc;
  DSMIL_REQUIRE(N >= 0 && N < 0xffffffffll, "N=%lld out of range", (long long)N);
  if (N == 0) {
    set_error("empty bag (N == 0): the reference raises IndexError at dsmil.py:53");
    return DSMIL_ERR_EMPTY;
  }
  DSMIL_REQUIRE(X && pred && A && B && (classes || classes_in), "NULL tensor pointer");
  if (use_sm100(p) && sm100::batched_supported(p) && (reinterpret_cast<uintptr_t>(X) & 15) == 0)
    return forward_bags_impl(p, &X, &N, 1, classes_in, classes, pred, A, B, crit_idx, save_Q, save_H1, ws, ws_bytes, st);
  bool ok;
  FwdWs w = carve_fwd(p, N, ws, ws_bytes, &ok);
  if (!ws || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, ws_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  float* Q = save_Q ? save_Q : w.Q;
  float* H1 = p->nonlinear ? (save_H1 ? save_H1 : w.H1) : nullptr;
  float* V = p->passing_v ? (save_V ? save_V : w.V) : nullptr;
  if ((rc = phase1_impl(p, X, xv, classes_in, N, 0, classes, Q, H1, V, w.cand, w.keys, w.wimg, st))) return rc;
  int64_t* crit = crit_idx ? crit_idx : w.crit;
  k_merge_cand<<<p->C, kQ, 0, st>>>(w.cand, 1, p->C, w.qmax, crit);
  DSMIL_LAUNCH_OK("k_merge_cand");
  const float* Vv = p->passing_v ? V : X;
  if ((rc = phase2_impl(p, Vv, Q, N, w.qmax, A, w.rec, w.recs, st))) return rc;
  return phase3_impl(p, N, w.rec, A, B, pred, st);
}


// ---- batched forward: a stream of bags in a handful of launches (tensor-core path only) -----------
struct BagsWs {
  sm100::BagDev* table;
  CUtensorMap* tmaps;       // one per bag (pair kernel: X of the bag as a 2-D TMA tensor)
  unsigned long long* keys;
  float* Q;
  uint8_t* wimg;
  float* recs;
  float* pred_part;
  unsigned int* counters;
  size_t bytes;
};
static inline int recs_for_bag(int64_t N) {
  const int64_t t = (N + sm100::kAttRows - 1) / sm100::kAttRows;
  return static_cast<int>(t < sm100::kMaxRecPerBag ? (t < 1 ? 1 : t) : sm100::kMaxRecPerBag);
}
static BagsWs carve_bags(const dsmil_params_t* p, const int64_t* Ns, int nb, bool need_Q, void* ws, size_t cap,
                         bool* ok) {
  Carver c(ws, cap);
  BagsWs w;
  int64_t total = 0, nrec = 0;
  for (int b = 0; b < nb; ++b) { total += Ns[b]; nrec += recs_for_bag(Ns[b]); }
  w.table = c.take<sm100::BagDev>(nb);
  w.tmaps = c.take<CUtensorMap>(nb);
  // keys and the finalize arrival counters are zeroed together (one memset): keep them adjacent
  w.keys = c.take<unsigned long long>(static_cast<size_t>(nb) * (kMaxC + 1));
  w.counters = reinterpret_cast<unsigned int*>(w.keys ? w.keys + static_cast<size_t>(nb) * kMaxC : nullptr);
  w.pred_part = c.take<float>(static_cast<size_t>(nb) * sm100::kFinSlices * kMaxC);
  {   // tile-blocked Q: one 128x128 block per 128-row tile
    int64_t tiles = 0;
    for (int b = 0; b < nb; ++b) tiles += (Ns[b] + sm100::kTileM - 1) / sm100::kTileM;
    w.Q = need_Q ? c.take<float>(static_cast<size_t>(tiles) * sm100::kTileM * kQ) : nullptr;
  }
  (void)total;
  w.wimg = c.take<uint8_t>(sm100::wimg_bytes(p->D) + 1024);
  w.recs = c.take<float>(static_cast<size_t>(nrec) * rec_floats(p->C, p->D));
  w.bytes = c.off;
  *ok = c.ok();
  return w;
}
static size_t l2_budget_bytes() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("DSMIL_B200_L2_MB");
    // Default: no sub-batching.  At the kernels' current speed a second HBM read of X (3 us per 10k-row bag)
    // costs less than the wave quantisation of small launches; set e.g. DSMIL_B200_L2_MB=72 to keep each
    // sub-batch L2-resident between phase 1 and phase 2 instead.
    const long mb = e ? atol(e) : (1l << 20);
    v = static_cast<size_t>(mb > 0 ? mb : (1l << 20)) << 20;
  }
  return v;
}

static int forward_bags_impl(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int nb,
                             const float* classes_in, float* classes, float* pred, float* A, float* B,
                             int64_t* crit, float* save_Q, float* save_H1, void* ws, size_t ws_bytes,
                             cudaStream_t st) {
  const int C = p->C, D = p->D;
  bool ok;
  BagsWs w = carve_bags(p, Ns, nb, save_Q == nullptr, ws, ws_bytes, &ok);
  if (!ws || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, ws_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  std::vector<sm100::BagDev> tbl(nb);
  long long row = 0;
  int tile = 0, rec = 0;
  // inference (Q stays in the tile-blocked workspace) on D <= 512, opt-in: the CTA-pair phase-1 kernel (fwd_pair.cuh)
  const bool pair_path = save_Q == nullptr && save_H1 == nullptr && use_pair(p);
  for (int b = 0; b < nb; ++b) {
    DSMIL_REQUIRE(Ns[b] >= 1 && Ns[b] < 0xffffffffll && Xs[b], "bag %d: empty or NULL", b);
    DSMIL_REQUIRE((reinterpret_cast<uintptr_t>(Xs[b]) & 15) == 0, "bag %d: features must be 16-byte aligned", b);
    const int nrec = recs_for_bag(Ns[b]);
    tbl[b] = sm100::BagDev{Xs[b], Ns[b], row, tile, rec, nrec, 0};
    row += Ns[b];
    tile += static_cast<int>((Ns[b] + sm100::kTileM - 1) / sm100::kTileM);
    rec += nrec;
  }
  DSMIL_CUDA_OK(cudaMemcpyAsync(w.table, tbl.data(), sizeof(sm100::BagDev) * nb, cudaMemcpyHostToDevice, st));
  DSMIL_CUDA_OK(cudaMemsetAsync(w.keys, 0, sizeof(unsigned long long) * (kMaxC + 1) * nb, st));
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w.wimg) + 1023) & ~uintptr_t(1023));
  int rc;
  if (pair_path) {
    std::vector<CUtensorMap> maps(nb);
    for (int b = 0; b < nb; ++b)
      if ((rc = pair::encode_bag_tmap(&maps[b], Xs[b], Ns[b], D))) return rc;
    DSMIL_CUDA_OK(cudaMemcpyAsync(w.tmaps, maps.data(), sizeof(CUtensorMap) * nb, cudaMemcpyHostToDevice, st));
    if ((rc = pair::launch_prep_wimg_pair(p, img, st))) return rc;
  } else {
    if ((rc = sm100::launch_prep_wimg(p, img, st))) return rc;
  }
  float* Q = save_Q ? save_Q : w.Q;
  if (classes_in) {   // bag form: arg-max of the given scores (single bag only)
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(Ns[0], 256), 296));
    k_argmax<<<grid, 256, 0, st>>>(classes_in, Ns[0], C, w.keys);
    DSMIL_LAUNCH_OK("k_argmax");
  }
  if (pair_path) {
    if ((rc = pair::launch_fwd_pair(p, w.table, w.tmaps, nb, tile, classes_in ? nullptr : classes, w.keys, Q, img,
                                    num_sms(), st)))
      return rc;
    sm100::AttendArgs aa{w.table, 0, nb, 0, D, C, Q, 1, w.keys, A, w.recs, nullptr};
    if ((rc = sm100::launch_attend_b(aa, rec, st))) return rc;
    sm100::FinalizeArgs fa{w.table, 0, D, C, w.recs, w.keys, p->Wf, p->bf, A, B, pred,
                           reinterpret_cast<long long*>(crit), w.pred_part, w.counters, nullptr, 0, 0};
    return sm100::launch_finalize_b(fa, nb, st);
  }
  // sub-batches sized so that a sub-batch's features (+Q) are still in L2 when the attend pass re-reads them
  const size_t budget = l2_budget_bytes();
  int b0 = 0;
  while (b0 < nb) {
    int b1 = b0;
    size_t bytes = 0;
    while (b1 < nb) {
      const size_t add = static_cast<size_t>(Ns[b1]) * (D + kQ) * sizeof(float);
      if (b1 > b0 && bytes + add > budget) break;
      bytes += add;
      ++b1;
    }
    const int t0 = tbl[b0].tile_off;
    const int t1 = (b1 < nb) ? tbl[b1].tile_off : tile;
    const int r0 = tbl[b0].rec_off;
    const int r1 = (b1 < nb) ? tbl[b1].rec_off : rec;
    const int q_blocked = save_Q ? 0 : blocked_q_mode();   // training keeps Q (after tanh) row-major for the backward kernels
    if ((rc = sm100::launch_qmlp(p, w.table, b0, b1 - b0, t0, t1 - t0, classes_in ? nullptr : classes, w.keys, Q,
                                 save_H1, img, num_sms(), st, q_blocked)))
      return rc;
    sm100::AttendArgs aa{w.table, b0, b1 - b0, r0, D, C, Q, q_blocked, w.keys, A, w.recs, nullptr};
    if ((rc = sm100::launch_attend_b(aa, r1 - r0, st))) return rc;
    sm100::FinalizeArgs fa{w.table, b0, D, C, w.recs, w.keys, p->Wf, p->bf, A, B, pred,
                           reinterpret_cast<long long*>(crit), w.pred_part, w.counters, nullptr, 0, 0};
    if ((rc = sm100::launch_finalize_b(fa, b1 - b0, st))) return rc;
    b0 = b1;
  }
  return 0;
}


// ---- row-sharded BATCH of bags (one call per phase for all bags; two all-gathers per step) --------------
struct ShardBagsWs {
  BagsWs base;
  long long* row_offsets;   // [nb] device copy
  float* qmax;              // [nb][C][128]
  size_t bytes;
};
static ShardBagsWs carve_shard_bags(const dsmil_params_t* p, const int64_t* Ns, int nb, void* ws, size_t cap, bool* ok) {
  ShardBagsWs s;
  s.base = carve_bags(p, Ns, nb, true, ws, cap, ok);
  Carver c(ws, cap);
  c.off = s.base.bytes;
  s.row_offsets = c.take<long long>(nb);
  s.qmax = c.take<float>(static_cast<size_t>(nb) * p->C * kQ);
  s.bytes = c.off;
  *ok = c.ok();
  return s;
}
static int build_table(const float* const* Xs, const int64_t* Ns, int nb, std::vector<sm100::BagDev>& tbl, int* tiles,
                       int* recs) {
  long long row = 0;
  int tile = 0, rec = 0;
  tbl.resize(nb);
  for (int b = 0; b < nb; ++b) {
    DSMIL_REQUIRE(Ns[b] >= 1 && Ns[b] < 0xffffffffll && Xs[b], "bag %d: empty or NULL (sharded batches need >= 1 row per rank)", b);
    DSMIL_REQUIRE((reinterpret_cast<uintptr_t>(Xs[b]) & 15) == 0, "bag %d: features must be 16-byte aligned", b);
    const int nrec = recs_for_bag(Ns[b]);
    tbl[b] = sm100::BagDev{Xs[b], Ns[b], row, tile, rec, nrec, 0};
    row += Ns[b];
    tile += static_cast<int>((Ns[b] + sm100::kTileM - 1) / sm100::kTileM);
    rec += nrec;
  }
  *tiles = tile;
  *recs = rec;
  return 0;
}

}  // namespace dsmil

using namespace dsmil;

extern "C" {

int dsmil_abi_version(void) { return DSMIL_ABI_VERSION; }
const char* dsmil_last_error(void) { return g_err; }
uint64_t dsmil_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int dsmil_forward_path(const dsmil_params_t* p, int64_t N) {
  (void)N;
  return (p && p->C >= 1 && p->C <= DSMIL_MAX_C && p->D >= 1 && p->D <= DSMIL_MAX_D && use_sm100(p)) ? 2 : 1;
}

/* Debug: CTA-0 timeline of the tensor-core kernel (clock64 stamps).  buf = device int64[3*8*64] or NULL. */
int dsmil_debug_set_trace(void* buf) {
  sm100::g_trace_buf = static_cast<long long*>(buf);
  return 0;
}

int dsmil_gather_rows(const float* X, int64_t N, int32_t D, const int64_t* idx, int64_t M, float* out, void* stream) {
  DSMIL_REQUIRE(N >= 0 && M >= 0 && D >= 1 && (M == 0 || (X && idx && out)), "bad arguments");
  if (M == 0) return 0;
  const int grid = static_cast<int>(std::min<int64_t>((M + 7) / 8, 148 * 8));
  k_gather_rows<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, D, reinterpret_cast<const long long*>(idx), M, out);
  DSMIL_LAUNCH_OK("k_gather_rows");
  return 0;
}

int dsmil_patches_u8_to_f32(const uint8_t* in, int64_t B, int32_t H, int32_t W, int32_t Cc, float* out, void* stream) {
  DSMIL_REQUIRE(B >= 0 && H >= 1 && W >= 1 && Cc >= 1 && Cc <= 4 && (B == 0 || (in && out)), "bad arguments");
  if (B == 0) return 0;
  const long long total = static_cast<long long>(B) * H * W;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  k_u8hwc_to_f32chw<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, B, H, W, Cc, out);
  DSMIL_LAUNCH_OK("k_u8hwc_to_f32chw");
  return 0;
}

int dsmil_instnorm_act(const float* x, const float* residual, float* y, int64_t planes, int32_t HW, float eps,
                       int32_t relu, void* stream) {
  DSMIL_REQUIRE(planes >= 0 && HW >= 1 && HW <= kInPlaneMax && eps >= 0.f && (planes == 0 || (x && y)),
                "bad arguments (HW must be in [1, %d])", kInPlaneMax);
  DSMIL_REQUIRE(planes < (1ll << 31), "too many planes");
  if (planes == 0) return 0;
  return launch_instnorm(x, residual, y, planes, HW, eps, relu, static_cast<cudaStream_t>(stream));
}

int dsmil_instnorm_act_nhwc(const float* x, const float* residual, float* y, int64_t N, int32_t HW, int32_t C, float eps,
                            int32_t relu, void* stream) {
  DSMIL_REQUIRE(N >= 0 && HW >= 1 && C >= 32 && C % 32 == 0 && eps >= 0.f && (N == 0 || (x && y)),
                "bad arguments (C must be a multiple of 32)");
  DSMIL_REQUIRE(N * (C / 32) < (1ll << 31), "too many (sample, channel group) slabs");
  if (N == 0) return 0;
  return launch_instnorm_nhwc(x, residual, y, N, HW, C, eps, relu, static_cast<cudaStream_t>(stream));
}

static JpegBatch carve_jpeg(void* ws, size_t cap, int n, int H, int W, int64_t blob_bytes, size_t* need) {
  Carver cv(ws, cap);
  JpegBatch a{};
  a.n = n; a.H = H; a.W = W;
  a.plane_elems = jpeg_plane_elems(H, W);
  a.unstuffed = cv.take<uint8_t>(static_cast<size_t>(blob_bytes) + 64);
  a.coef = cv.take<int16_t>(static_cast<size_t>(3) * a.plane_elems * n);
  a.planes = cv.take<uint8_t>(static_cast<size_t>(3) * a.plane_elems * n);
  *need = cv.off;
  return a;
}

int32_t dsmil_jpeg_header_bytes_dev(void) { return static_cast<int32_t>(sizeof(dsmil_jpeg_header)); }

int64_t dsmil_jpeg_workspace_bytes(int32_t n, int32_t H, int32_t W, int64_t blob_bytes) {
  if (n < 0 || H < 1 || W < 1 || H > 65535 || W > 65535 || blob_bytes < 0) return -1;
  size_t need = 0;
  carve_jpeg(nullptr, 0, n, H, W, blob_bytes, &need);
  return static_cast<int64_t>(need);
}

int dsmil_jpeg_decode_batch(const uint8_t* blob, int64_t blob_bytes, const void* headers, int32_t n, int32_t H, int32_t W,
                            uint8_t* out_u8, float* out_f32, int32_t f32_channels_last, int32_t* status, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  DSMIL_REQUIRE(n >= 0 && H >= 1 && W >= 1 && H <= 65535 && W <= 65535 && blob_bytes >= 0 &&
                (f32_channels_last == 0 || f32_channels_last == 1), "bad arguments");
  if (n == 0) return 0;
  DSMIL_REQUIRE(blob && headers && status && workspace && (out_u8 || out_f32), "null pointer");
  DSMIL_REQUIRE((reinterpret_cast<uintptr_t>(headers) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
                (out_f32 == nullptr || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) &&
                (out_u8 == nullptr || (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0),
                "headers must be 16-byte, workspace 256-byte, out_f32 16-byte, out_u8 4-byte aligned");
  size_t need = 0;
  JpegBatch a = carve_jpeg(workspace, static_cast<size_t>(workspace_bytes), n, H, W, blob_bytes, &need);
  if (static_cast<int64_t>(need) > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %lld", need, static_cast<long long>(workspace_bytes));
    return DSMIL_ERR_WORKSPACE;
  }
  a.blob = blob;
  a.hdr = static_cast<const dsmil_jpeg_header*>(headers);
  a.out_u8 = out_u8;
  a.out_f32 = out_f32;
  a.f32_hwc = f32_channels_last;
  a.status = status;
  return launch_jpeg_decode(a, static_cast<cudaStream_t>(stream));
}

int dsmil_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
int dsmil_profile_read(double* ms_per_tag, uint64_t* launches_per_tag) {
  DSMIL_REQUIRE(ms_per_tag && launches_per_tag, "NULL output");
  for (int t = 0; t < PROF_NTAGS; ++t) { ms_per_tag[t] = 0.0; launches_per_tag[t] = 0; }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_per_tag[r.tag] += ms;
      launches_per_tag[r.tag] += 1;
    }
    g_prof_pool.push_back(r);
  }
  g_prof.clear();
  cudaGetLastError();
  return 0;
}

size_t dsmil_cand_floats(int32_t C) { return cand_floats(C); }
size_t dsmil_rec_floats(int32_t C, int32_t Dv) { return rec_floats(C, Dv); }

size_t dsmil_forward_workspace_bytes(const dsmil_params_t* p, int64_t N) {
  if (!p || p->C < 1 || p->C > DSMIL_MAX_C || p->D < 1 || p->D > DSMIL_MAX_D || N < 0) return 0;
  bool ok;
  size_t a = carve_fwd(p, N, nullptr, 0, &ok).bytes;
  if (sm100::batched_supported(p) && N > 0) a = std::max(a, carve_bags(p, &N, 1, true, nullptr, 0, &ok).bytes);
  return a;
}

size_t dsmil_forward_bags_workspace_bytes(const dsmil_params_t* p, const int64_t* Ns, int32_t nb) {
  if (!p || !Ns || nb < 1 || p->C < 1 || p->C > DSMIL_MAX_C || p->D < 1 || p->D > DSMIL_MAX_D) return 0;
  bool ok;
  // dsmil_forward_bags may take either route (tensor-core batch, or the per-bag generic loop when
  // DSMIL_B200_GENERIC=1 / a bag is not 16-byte aligned): the workspace must cover both layouts.
  int64_t mx = 0;
  for (int b = 0; b < nb; ++b) mx = std::max<int64_t>(mx, Ns[b]);
  size_t bytes = carve_fwd(p, mx, nullptr, 0, &ok).bytes;
  if (sm100::batched_supported(p)) bytes = std::max(bytes, carve_bags(p, Ns, nb, true, nullptr, 0, &ok).bytes);
  return bytes;
}

int dsmil_forward_bags(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int32_t nb,
                       float* classes, float* pred, float* A, float* B, int64_t* crit_idx, void* workspace,
                       size_t workspace_bytes, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  DSMIL_REQUIRE(Xs && Ns && nb >= 1 && classes && pred && A && B, "NULL pointer or nb < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {   // one contract for both routes below: the size dsmil_forward_bags_workspace_bytes reports
    const size_t need = dsmil_forward_bags_workspace_bytes(p, Ns, nb);
    if (!workspace || workspace_bytes < need) {
      set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : size_t(0));
      return DSMIL_ERR_WORKSPACE;
    }
  }
  bool aligned = true;
  for (int b = 0; b < nb; ++b) aligned = aligned && Xs[b] && (reinterpret_cast<uintptr_t>(Xs[b]) & 15) == 0 && Ns[b] >= 1;
  if (use_sm100(p) && sm100::batched_supported(p) && aligned)
    return forward_bags_impl(p, Xs, Ns, nb, nullptr, classes, pred, A, B, crit_idx, nullptr, nullptr, workspace,
                             workspace_bytes, st);
  // shapes the tensor-core kernels do not take: same packed outputs, one bag at a time
  int64_t row = 0;
  for (int b = 0; b < nb; ++b) {
    rc = forward_impl(p, Xs[b], nullptr, nullptr, Ns[b], classes + row * p->C, pred + static_cast<size_t>(b) * p->C,
                      A + row * p->C, B + static_cast<size_t>(b) * p->C * p->D,
                      crit_idx ? crit_idx + static_cast<size_t>(b) * p->C : nullptr, nullptr, nullptr, nullptr,
                      workspace, workspace_bytes, st);
    if (rc) return rc;
    row += Ns[b];
  }
  return 0;
}
size_t dsmil_shard_workspace_bytes(const dsmil_params_t* p, int64_t N_local) {
  return dsmil_forward_workspace_bytes(p, N_local);
}

int dsmil_forward(const dsmil_params_t* p, const float* X, const float* x_for_v, int64_t N, float* classes,
                  float* pred, float* A, float* B, int64_t* crit_idx, float* save_Q, float* save_H1,
                  float* save_V, void* workspace, size_t workspace_bytes, void* stream) {
  DSMIL_REQUIRE(classes != nullptr, "classes is NULL");
  return forward_impl(p, X, x_for_v, nullptr, N, classes, pred, A, B, crit_idx, save_Q, save_H1, save_V,
                      workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dsmil_bag_forward(const dsmil_params_t* p, const float* X, const float* x_for_v, const float* classes_in,
                      int64_t N, float* pred, float* A, float* B, int64_t* crit_idx, float* save_Q,
                      float* save_H1, float* save_V, void* workspace, size_t workspace_bytes, void* stream) {
  DSMIL_REQUIRE(classes_in != nullptr, "classes_in is NULL");
  return forward_impl(p, X, x_for_v, classes_in, N, nullptr, pred, A, B, crit_idx, save_Q, save_H1, save_V,
                      workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dsmil_instance_scores(const dsmil_params_t* p, const float* X, int64_t N, float* classes, void* stream) {
  DSMIL_REQUIRE(p && p->Wi && p->bi && p->C >= 1 && p->C <= DSMIL_MAX_C && p->D >= 1 && p->D <= DSMIL_MAX_D,
                "bad params");
  DSMIL_REQUIRE(N >= 0 && (N == 0 || (X && classes)), "NULL tensor pointer");
  if (N == 0) return 0;
  // The arg-max by-product goes to a scratch key slot that is simply ignored here.
  unsigned long long* keys;
  DSMIL_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&keys), scratch_keys));
  return launch_scores(p, X, N, classes, keys, static_cast<cudaStream_t>(stream));
}

// ---- sharded phases -----------------------------------------------------------------------
int dsmil_shard_phase1(const dsmil_params_t* p, const float* X, const float* x_for_v, const float* classes_in,
                       int64_t N_local, int64_t row_offset, float* classes, float* Q, float* H1, float* V,
                       float* cand_rec, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p, classes_in == nullptr);
  if (rc) return rc;
  DSMIL_REQUIRE(N_local >= 0 && N_local < 0xffffffffll, "N_local out of range");
  DSMIL_REQUIRE(cand_rec && (N_local == 0 || (X && Q && (classes || classes_in))), "NULL tensor pointer");
  DSMIL_REQUIRE(N_local == 0 || !p->passing_v || V, "passing_v needs a V buffer");
  bool ok;
  FwdWs w = carve_fwd(p, N_local, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  // the tensor-core path keeps H1 in TMEM; the generic path (also taken for an unaligned X) needs a buffer
  const bool tc = use_sm100(p) && w.wimg && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  float* h1 = p->nonlinear ? (H1 ? H1 : (tc ? nullptr : w.H1)) : nullptr;
  return phase1_impl(p, X, x_for_v, classes_in, N_local, row_offset, classes, Q, h1, V, cand_rec, w.keys, w.wimg,
                     static_cast<cudaStream_t>(stream));
}

int dsmil_shard_merge_candidates(int32_t C, const float* cand_recs, int32_t G, float* q_max, int64_t* crit_idx,
                                 void* stream) {
  DSMIL_REQUIRE(C >= 1 && C <= DSMIL_MAX_C && G >= 1 && cand_recs && q_max && crit_idx, "bad arguments");
  k_merge_cand<<<C, kQ, 0, static_cast<cudaStream_t>(stream)>>>(cand_recs, G, C, q_max, crit_idx);
  DSMIL_LAUNCH_OK("k_merge_cand");
  return 0;
}

int dsmil_shard_phase2(const dsmil_params_t* p, const float* Xv, const float* Q, int64_t N_local,
                       const float* q_max, float* A_logits, float* rec, void* workspace, size_t workspace_bytes,
                       void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(rec && q_max && (N_local == 0 || (Xv && Q && A_logits)), "NULL tensor pointer");
  bool ok;
  FwdWs w = carve_fwd(p, N_local, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  return phase2_impl(p, Xv, Q, N_local, q_max, A_logits, rec, w.recs, static_cast<cudaStream_t>(stream));
}

int dsmil_shard_merge_partials(int32_t C, int32_t Dv, const float* recs, int32_t G, float* rec_out, void* stream) {
  DSMIL_REQUIRE(C >= 1 && C <= DSMIL_MAX_C && Dv >= 1 && G >= 1 && G <= kMaxRecs && recs && rec_out, "bad arguments");
  dim3 g2(C, ceil_div(Dv, 256));
  k_combine_rec<<<g2, 256, 0, static_cast<cudaStream_t>(stream)>>>(recs, G, C, Dv, rec_out);
  DSMIL_LAUNCH_OK("k_combine_rec");
  return 0;
}

int dsmil_shard_phase3(const dsmil_params_t* p, int64_t N_local, const float* rec_global, float* A, float* B,
                       float* pred, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(rec_global && B && pred && (N_local == 0 || A), "NULL tensor pointer");
  return phase3_impl(p, N_local, rec_global, A, B, pred, static_cast<cudaStream_t>(stream));
}

// ---- backward -----------------------------------------------------------------------------
struct BwdWs {
  float *dB, *dA, *tpart, *dqm, *dz2, *dz1, *tnpart, *cspart, *dzv, *tmp;
  size_t bytes;
};
static BwdWs carve_bwd(const dsmil_params_t* p, int64_t N, int need_gX, void* ws, size_t cap, bool* ok) {
  Carver c(ws, cap);
  BwdWs w;
  const int C = p->C, D = p->D;
  const int64_t n = N > 0 ? N : 1;
  w.dB = c.take<float>(static_cast<size_t>(C) * D);
  w.dA = c.take<float>(n * C);
  w.tpart = c.take<float>(296 * kMaxC);
  w.dqm = c.take<float>(static_cast<size_t>(C) * kQ);
  w.dz2 = c.take<float>(n * kQ);
  w.dz1 = p->nonlinear ? c.take<float>(n * kQ) : nullptr;
  size_t tn = tn_partial_floats(kQ, D, N);
  tn = std::max(tn, tn_partial_floats(kQ, kQ, N));
  tn = std::max(tn, tn_partial_floats(C, kQ, N));
  tn = std::max(tn, tn_partial_floats(C, D, N));
  if (p->passing_v) tn = std::max(tn, tn_partial_floats(D, D, N));
  w.tnpart = c.take<float>(tn);
  w.cspart = c.take<float>(static_cast<size_t>(296) * std::max(D, kQ));
  w.dzv = p->passing_v ? c.take<float>(n * D) : nullptr;
  w.tmp = (p->passing_v && need_gX) ? c.take<float>(n * D) : nullptr;
  w.bytes = c.off;
  *ok = c.ok();
  return w;
}

size_t dsmil_backward_workspace_bytes(const dsmil_params_t* p, int64_t N, int need_gX) {
  if (!p || p->C < 1 || p->C > DSMIL_MAX_C || p->D < 1 || p->D > DSMIL_MAX_D || N < 0) return 0;
  bool ok;
  return carve_bwd(p, N, need_gX, nullptr, 0, &ok).bytes;
}

int dsmil_backward(const dsmil_params_t* p, const float* X, const float* x_for_v, int64_t N, const float* Q,
                   const float* H1, const float* V, const float* A, const float* B, const int64_t* crit_idx,
                   const float* d_classes, const float* d_pred, const float* d_A, const float* d_B,
                   const dsmil_grads_t* g, const float* v_mask, void* workspace, size_t workspace_bytes,
                   void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(N >= 1 && X && Q && A && B && crit_idx && g, "NULL tensor pointer or N < 1");
  DSMIL_REQUIRE(!p->nonlinear || H1, "nonlinear q backward needs saved H1");
  DSMIL_REQUIRE(!p->passing_v || V, "passing_v backward needs saved V");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = p->C, D = p->D;
  bool ok;
  BwdWs w = carve_bwd(p, N, g->gX != nullptr, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  const float* Vv = p->passing_v ? V : X;
  const float* Xv = x_for_v ? x_for_v : X;
  const int gs = static_cast<int>(std::min<int64_t>(ceil_div(N, 256), 296));

  // bag classifier (dsmil.py:59-61) and B
  k_bwd_bag<<<ceil_div(static_cast<int64_t>(C) * D, 256), 256, 0, st>>>(p->Wf, B, d_pred, d_B, C, D, w.dB, g->gWf,
                                                                        g->gbf);
  DSMIL_LAUNCH_OK("k_bwd_bag");
  // instance classifier (dsmil.py:11): only rows with non-zero upstream grad contribute
  if (g->gWi) {
    if (d_classes) { if ((rc = launch_gemm_tn(d_classes, C, X, D, N, w.tnpart, g->gWi, st))) return rc; }
    else DSMIL_CUDA_OK(cudaMemsetAsync(g->gWi, 0, sizeof(float) * C * D, st));
  }
  if (g->gbi) {
    if (d_classes) { if ((rc = launch_colsum(d_classes, C, N, w.cspart, g->gbi, st))) return rc; }
    else DSMIL_CUDA_OK(cudaMemsetAsync(g->gbi, 0, sizeof(float) * C, st));
  }
  // dA = V dB^T (+ upstream), softmax-over-instances backward (dsmil.py:56-57)
  {
    const size_t smem = sizeof(float) * C * D;
    if (smem > 48 * 1024)
      DSMIL_CUDA_OK(cudaFuncSetAttribute(k_rowdot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N, 8), 148 * 8));
    k_rowdot<<<grid, 256, smem, st>>>(Vv, N, D, w.dB, C, d_A, w.dA);
    DSMIL_LAUNCH_OK("k_rowdot");
  }
  k_bwd_t_partial<<<gs, 256, 0, st>>>(A, w.dA, N, C, w.tpart);
  DSMIL_LAUNCH_OK("k_bwd_t_partial");
  k_bwd_dL<<<gs, 256, 0, st>>>(A, w.dA, N, C, w.tpart, gs);
  DSMIL_LAUNCH_OK("k_bwd_dL");
  const float* dL = w.dA;
  // dq_max = dL^T Q  (dsmil.py:55), then dQ rows (+ the critical rows' share, dsmil.py:53-54)
  if ((rc = launch_gemm_tn(dL, C, Q, kQ, N, w.tnpart, w.dqm, st))) return rc;
  {
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N * kQ, 256), 148 * 8));
    k_bwd_dq<<<grid, 256, 0, st>>>(dL, Q, w.dqm, crit_idx, N, C, p->nonlinear, w.dz2);
    DSMIL_LAUNCH_OK("k_bwd_dq");
  }
  const float* dz1 = w.dz2;
  if (p->nonlinear) {
    if (g->gW2 && (rc = launch_gemm_tn(w.dz2, kQ, H1, kQ, N, w.tnpart, g->gW2, st))) return rc;
    if (g->gb2 && (rc = launch_colsum(w.dz2, kQ, N, w.cspart, g->gb2, st))) return rc;
    if ((rc = launch_linear<ACT_MASK_POS, true>(w.dz2, N, kQ, p->W2, nullptr, kQ, w.dz1, H1, 0, st))) return rc;
    dz1 = w.dz1;
  }
  if (g->gW1 && (rc = launch_gemm_tn(dz1, kQ, X, D, N, w.tnpart, g->gW1, st))) return rc;
  if (g->gb1 && (rc = launch_colsum(dz1, kQ, N, w.cspart, g->gb1, st))) return rc;

  const int ge = static_cast<int>(std::min<int64_t>(ceil_div(N * D, 256), 148 * 8));
  if (p->passing_v) {
    k_bwd_dzv<<<ge, 256, 0, st>>>(A, w.dB, V, N, C, D, w.dzv);
    DSMIL_LAUNCH_OK("k_bwd_dzv");
    if (g->gWv && (rc = launch_gemm_tn(w.dzv, D, Xv, D, N, w.tnpart, g->gWv, st))) return rc;
    if (g->gbv && (rc = launch_colsum(w.dzv, D, N, w.cspart, g->gbv, st))) return rc;
  }
  if (g->gX) {
    if ((rc = launch_linear<ACT_NONE, true>(dz1, N, kQ, p->W1, nullptr, D, g->gX, nullptr, 0, st))) return rc;
    k_bwd_dx_extra<<<ge, 256, 0, st>>>(d_classes, p->Wi, p->passing_v ? nullptr : A, w.dB, N, C, D, 1, g->gX);
    DSMIL_LAUNCH_OK("k_bwd_dx_extra");
    if (p->passing_v) {
      if ((rc = launch_linear<ACT_NONE, true>(w.dzv, N, D, p->Wv, nullptr, D, w.tmp, nullptr, 0, st))) return rc;
      k_axpy_mask<<<ge, 256, 0, st>>>(w.tmp, v_mask, N * D, g->gX);
      DSMIL_LAUNCH_OK("k_axpy_mask");
    }
  }
  return 0;
}

// ---- row-sharded backward: dsmil_backward cut at its two cross-row sums (+ the caller's grad all-reduce) ----
static int shard_bwd_ws(const dsmil_params_t* p, int64_t N, void* ws, size_t cap, BwdWs* w) {
  bool ok;
  *w = carve_bwd(p, N, 0, ws, cap, &ok);
  if (!ws || !ok) {
    set_error("workspace too small: need %zu bytes, got %zu", w->bytes, cap);
    return DSMIL_ERR_WORKSPACE;
  }
  return 0;
}

int dsmil_shard_backward_phase1(const dsmil_params_t* p, const float* X, int64_t N, const float* A, const float* B,
                                const float* d_classes, const float* d_pred, float* dA, float* t_local, float* gWi,
                                float* gbi, float* gWf, float* gbf, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(!p->passing_v, "sharded backward supports the identity v only");
  DSMIL_REQUIRE(N >= 0 && B && t_local && (N == 0 || (X && A && dA)), "NULL tensor pointer or N < 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = p->C, D = p->D;
  BwdWs w;
  if ((rc = shard_bwd_ws(p, N, workspace, workspace_bytes, &w))) return rc;
  k_bwd_bag<<<ceil_div(static_cast<int64_t>(C) * D, 256), 256, 0, st>>>(p->Wf, B, d_pred, nullptr, C, D, w.dB, gWf, gbf);
  DSMIL_LAUNCH_OK("k_bwd_bag");
  if (gWi) {
    if (d_classes && N > 0) { if ((rc = launch_gemm_tn(d_classes, C, X, D, N, w.tnpart, gWi, st))) return rc; }
    else DSMIL_CUDA_OK(cudaMemsetAsync(gWi, 0, sizeof(float) * C * D, st));
  }
  if (gbi) {
    if (d_classes && N > 0) { if ((rc = launch_colsum(d_classes, C, N, w.cspart, gbi, st))) return rc; }
    else DSMIL_CUDA_OK(cudaMemsetAsync(gbi, 0, sizeof(float) * C, st));
  }
  if (N == 0) {
    DSMIL_CUDA_OK(cudaMemsetAsync(t_local, 0, sizeof(float) * C, st));
    return 0;
  }
  const size_t smem = sizeof(float) * C * D;
  if (smem > 48 * 1024)
    DSMIL_CUDA_OK(cudaFuncSetAttribute(k_rowdot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N, 8), 148 * 8));
  k_rowdot<<<grid, 256, smem, st>>>(X, N, D, w.dB, C, nullptr, dA);
  DSMIL_LAUNCH_OK("k_rowdot");
  const int gs = static_cast<int>(std::min<int64_t>(ceil_div(N, 256), 296));
  k_bwd_t_partial<<<gs, 256, 0, st>>>(A, dA, N, C, w.tpart);
  DSMIL_LAUNCH_OK("k_bwd_t_partial");
  k_sum_partials<<<1, 256, 0, st>>>(w.tpart, gs, C, t_local);
  DSMIL_LAUNCH_OK("k_sum_partials");
  return 0;
}

int dsmil_shard_backward_phase2(const dsmil_params_t* p, int64_t N, const float* A, float* dA, const float* t_global,
                                const float* Q, float* dqm_local, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(N >= 0 && t_global && dqm_local && (N == 0 || (A && dA && Q)), "NULL tensor pointer or N < 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = p->C;
  BwdWs w;
  if ((rc = shard_bwd_ws(p, N, workspace, workspace_bytes, &w))) return rc;
  if (N > 0) {
    const int gs = static_cast<int>(std::min<int64_t>(ceil_div(N, 256), 296));
    k_bwd_dL<<<gs, 256, 0, st>>>(A, dA, N, C, t_global, 1);
    DSMIL_LAUNCH_OK("k_bwd_dL");
  }
  return launch_gemm_tn(dA, C, Q, kQ, N, w.tnpart, dqm_local, st);   // zeros when N == 0
}

int dsmil_shard_backward_phase3(const dsmil_params_t* p, const float* X, int64_t N, int64_t row_offset, const float* Q,
                                const float* H1, const float* dL, const float* dqm_global, const float* q_max,
                                const int64_t* crit_idx, float* gW1, float* gb1, float* gW2, float* gb2,
                                void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p);
  if (rc) return rc;
  DSMIL_REQUIRE(N >= 0 && dqm_global && q_max && crit_idx && (N == 0 || (X && Q && dL)), "NULL tensor pointer or N < 0");
  DSMIL_REQUIRE(!p->nonlinear || N == 0 || H1, "nonlinear q backward needs saved H1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = p->D, C = p->C;
  BwdWs w;
  if ((rc = shard_bwd_ws(p, N, workspace, workspace_bytes, &w))) return rc;
  const float* dz1 = w.dz2;
  if (N > 0) {
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(N * kQ, 256), 148 * 8));
    k_bwd_dq_shard<<<grid, 256, 0, st>>>(dL, Q, q_max, dqm_global, crit_idx, N, row_offset, C, p->nonlinear, w.dz2);
    DSMIL_LAUNCH_OK("k_bwd_dq_shard");
  }
  if (p->nonlinear) {
    if (gW2 && (rc = launch_gemm_tn(w.dz2, kQ, H1, kQ, N, w.tnpart, gW2, st))) return rc;
    if (gb2 && (rc = launch_colsum(w.dz2, kQ, N, w.cspart, gb2, st))) return rc;
    if (N > 0 && (rc = launch_linear<ACT_MASK_POS, true>(w.dz2, N, kQ, p->W2, nullptr, kQ, w.dz1, H1, 0, st))) return rc;
    dz1 = w.dz1;
  }
  if (gW1 && (rc = launch_gemm_tn(dz1, kQ, X, D, N, w.tnpart, gW1, st))) return rc;
  if (gb1 && (rc = launch_colsum(dz1, kQ, N, w.cspart, gb1, st))) return rc;
  return 0;
}

int dsmil_instance_scores_backward(const dsmil_params_t* p, const float* X, int64_t N, const float* d_classes,
                                   float* gWi, float* gbi, float* gX, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  DSMIL_REQUIRE(p && p->Wi && p->C >= 1 && p->C <= DSMIL_MAX_C && p->D >= 1 && p->D <= DSMIL_MAX_D, "bad params");
  DSMIL_REQUIRE(N >= 1 && X && d_classes, "NULL tensor pointer or N < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = p->C, D = p->D;
  Carver c(workspace, workspace_bytes);
  float* tnpart = c.take<float>(tn_partial_floats(C, D, N));
  float* cspart = c.take<float>(static_cast<size_t>(296) * kMaxC);
  if (!workspace || !c.ok()) {
    set_error("workspace too small: need %zu bytes, got %zu", c.off, workspace_bytes);
    return DSMIL_ERR_WORKSPACE;
  }
  int rc;
  if (gWi && (rc = launch_gemm_tn(d_classes, C, X, D, N, tnpart, gWi, st))) return rc;
  if (gbi && (rc = launch_colsum(d_classes, C, N, cspart, gbi, st))) return rc;
  if (gX) {
    const int ge = static_cast<int>(std::min<int64_t>(ceil_div(N * D, 256), 148 * 8));
    k_bwd_dx_extra<<<ge, 256, 0, st>>>(d_classes, p->Wi, nullptr, nullptr, N, C, D, 0, gX);
    DSMIL_LAUNCH_OK("k_bwd_dx_extra");
  }
  return 0;
}

// ---- sharded batch ABI ---------------------------------------------------------------------------
int dsmil_shard_bags_supported(const dsmil_params_t* p) {
  return (p && p->C >= 1 && p->C <= DSMIL_MAX_C && p->D >= 1 && p->D <= DSMIL_MAX_D && use_sm100(p) &&
          sm100::batched_supported(p)) ? 1 : 0;
}
size_t dsmil_shard_bags_workspace_bytes(const dsmil_params_t* p, const int64_t* Ns, int32_t nb) {
  if (!dsmil_shard_bags_supported(p) || !Ns || nb < 1) return 0;
  bool ok;
  return carve_shard_bags(p, Ns, nb, nullptr, 0, &ok).bytes;
}
int dsmil_shard_bags_phase1(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int32_t nb,
                            const int64_t* row_offsets, float* classes, float* cand_recs, void* workspace,
                            size_t workspace_bytes, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  DSMIL_REQUIRE(dsmil_shard_bags_supported(p), "shape not supported by the batched tensor-core path");
  DSMIL_REQUIRE(Xs && Ns && nb >= 1 && row_offsets && classes && cand_recs, "NULL pointer or nb < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool ok;
  ShardBagsWs w = carve_shard_bags(p, Ns, nb, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) { set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes); return DSMIL_ERR_WORKSPACE; }
  std::vector<sm100::BagDev> tbl;
  int tiles = 0, recs = 0;
  if ((rc = build_table(Xs, Ns, nb, tbl, &tiles, &recs))) return rc;
  if (!stream_is_capturing(st)) {
    DSMIL_CUDA_OK(cudaMemcpyAsync(w.base.table, tbl.data(), sizeof(sm100::BagDev) * nb, cudaMemcpyHostToDevice, st));
    DSMIL_CUDA_OK(cudaMemcpyAsync(w.row_offsets, row_offsets, sizeof(long long) * nb, cudaMemcpyHostToDevice, st));
  }
  DSMIL_CUDA_OK(cudaMemsetAsync(w.base.keys, 0, sizeof(unsigned long long) * (kMaxC + 1) * nb, st));
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w.base.wimg) + 1023) & ~uintptr_t(1023));
  // (a captured serving loop also reuses the weight images of the preceding eager call: re-capture after a weight update)
  if (!stream_is_capturing(st) && (rc = sm100::launch_prep_wimg(p, img, st))) return rc;
  if ((rc = sm100::launch_qmlp(p, w.base.table, 0, nb, 0, tiles, classes, w.base.keys, w.base.Q, nullptr, img, num_sms(), st,
                               blocked_q_mode())))
    return rc;
  sm100::k_gather_cand_b<<<dim3(p->C, nb), kQ, 0, st>>>(w.base.table, w.base.keys, classes, w.base.Q, blocked_q_mode(), w.row_offsets, p->C,
                                                        cand_recs);
  DSMIL_LAUNCH_OK("k_gather_cand_b");
  return 0;
}
int dsmil_shard_bags_phase2(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int32_t nb,
                            const float* cands_all, int32_t G, float* A, int64_t* crit_idx, float* recs_out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  DSMIL_REQUIRE(dsmil_shard_bags_supported(p) && Xs && Ns && nb >= 1 && cands_all && G >= 1 && A && crit_idx && recs_out,
                "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool ok;
  ShardBagsWs w = carve_shard_bags(p, Ns, nb, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) { set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes); return DSMIL_ERR_WORKSPACE; }
  std::vector<sm100::BagDev> tbl;
  int tiles = 0, recs = 0;
  if ((rc = build_table(Xs, Ns, nb, tbl, &tiles, &recs))) return rc;   // the device table was written by phase 1
  sm100::k_merge_cand_b<<<dim3(p->C, nb), kQ, 0, st>>>(cands_all, G, nb, p->C, w.qmax, reinterpret_cast<long long*>(crit_idx));
  DSMIL_LAUNCH_OK("k_merge_cand_b");
  sm100::AttendArgs aa{w.base.table, 0, nb, 0, p->D, p->C, w.base.Q, blocked_q_mode(), w.base.keys, A, w.base.recs, w.qmax};
  if ((rc = sm100::launch_attend_b(aa, recs, st))) return rc;
  sm100::FinalizeArgs fa{w.base.table, 0, p->D, p->C, w.base.recs, w.base.keys, p->Wf, p->bf, A, nullptr, nullptr, nullptr,
                         w.base.pred_part, w.base.counters, recs_out, 0, 0};
  return sm100::launch_finalize_b(fa, nb, st);
}
int dsmil_shard_bags_phase3(const dsmil_params_t* p, const float* const* Xs, const int64_t* Ns, int32_t nb,
                            const float* recs_all, int32_t G, float* A, float* B, float* pred, void* workspace,
                            size_t workspace_bytes, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  DSMIL_REQUIRE(dsmil_shard_bags_supported(p) && Xs && Ns && nb >= 1 && recs_all && G >= 1 && G <= sm100::kMaxRecPerBag && A && B && pred,
                "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool ok;
  ShardBagsWs w = carve_shard_bags(p, Ns, nb, workspace, workspace_bytes, &ok);
  if (!workspace || !ok) { set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes); return DSMIL_ERR_WORKSPACE; }
  sm100::FinalizeArgs fa{w.base.table, 0, p->D, p->C, recs_all, w.base.keys, p->Wf, p->bf, A, B, pred, nullptr,
                         w.base.pred_part, w.base.counters, nullptr, G, nb};
  return sm100::launch_finalize_b(fa, nb, st);
}

}  // extern "C"
