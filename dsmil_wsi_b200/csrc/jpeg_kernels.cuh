// Patch loader of the embedding loop on the device: the reference decodes every patch with PIL in
// DataLoader workers (compute_feats.py:26-29,55) and ships 77 MB of fp32 per 128-patch batch over PCIe
// (compute_feats.py:72).  Here the FILES cross PCIe (~2 MB per batch) and three kernels turn them into the
// [B, 3, H, W] fp32 tensor the backbone reads (== VF.to_tensor(Image.open(path)) bit for bit; jpeg_core.h holds the
// arithmetic, checked against PIL on the CPU as well):
//
//   k_jpeg_entropy   one WARP per patch.  All lanes: copy the entropy-coded segment with the stuffed zero bytes and
//                    RSTn markers removed (ballot + popc compaction), build the Huffman lookup tables in shared
//                    memory (one lane per table).  Lane 0: the serial part of JPEG -- Huffman decoding of every MCU
//                    (T.81 F.2.2) into quantised coefficients (int16, natural order).  128 patches = 128 independent
//                    warps, ~100 cycles per symbol; the kernel is latency-bound by construction (the bitstream has
//                    no entry points without restart markers) and is meant to run on a side stream under the
//                    backbone of the previous batch, where it occupies one warp slot on 128 SMs.
//   k_jpeg_idct      one THREAD per 8x8 block: dequantise + the 13-bit fixed-point LLM IDCT (JDCT_ISLOW) in
//                    registers, rows written as 8-byte stores into the component plane.
//   k_jpeg_color     one thread per 4 output pixels: fancy (triangle) chroma upsampling 4:2:0 / 4:2:2, JFIF
//                    YCbCr -> RGB, and BOTH output forms in one pass: uint8 HWC (what PIL returns) and/or
//                    float32 CHW / 255 (what VF.to_tensor returns), so the separate conversion pass disappears.
//
// HBM traffic per 224x224 4:2:0 patch: coefficients 150 KB written + read, planes 75 KB written + read, output
// 588 KB (fp32) -- about 1 MB, i.e. < 0.1 ms per 128-patch batch for the two data-parallel kernels.
#pragma once
#include "common.cuh"
#include <cstddef>
#include "jpeg_core.h"

namespace dsmil {

static_assert(sizeof(dsmil_jpeg_header) % 16 == 0 && offsetof(dsmil_jpeg_header, qt) % 16 == 0, "header layout");

__constant__ uint8_t c_jpeg_natural[64] = DSMIL_JPEG_NATURAL_ORDER;

struct JpegBatch {
  const uint8_t* blob;            // the files, back to back (device)
  const dsmil_jpeg_header* hdr;   // [n] parsed headers (device)
  int n, H, W;                    // every patch of the batch is H x W
  long long plane_elems;          // (round16 H) * (round16 W): capacity of one component's coefficient / sample plane
  uint8_t* unstuffed;             // blob_bytes + 64
  int16_t* coef;                  // [n][3][plane_elems], zeroed before k_jpeg_entropy
  uint8_t* planes;                // [n][3][plane_elems]
  uint8_t* out_u8;                // [n][H][W][3] or null
  float* out_f32;                 // [n][3][H][W] (f32_hwc == 0) or [n][H][W][3] (f32_hwc == 1: torch.channels_last) or null
  int f32_hwc;
  int32_t* status;                // [n]
};

__device__ __forceinline__ bool jpeg_usable(const dsmil_jpeg_header& h, const JpegBatch& a) {
  return h.status == DSMIL_JPEG_OK && h.width == a.W && h.height == a.H;
}

__global__ void __launch_bounds__(32)
k_jpeg_entropy(JpegBatch a) {
  __shared__ dsmil_jpeg_htab s_tab[8];
  __shared__ __align__(16) dsmil_jpeg_header s_hdr;
  __shared__ uint8_t s_nat[64];
  __shared__ int s_rc;
  const int img = blockIdx.x, lane = threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.hdr + img);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_hdr);
    for (int i = lane; i < static_cast<int>(sizeof(dsmil_jpeg_header) / 4); i += 32) dst[i] = __ldg(src + i);
    s_nat[lane] = c_jpeg_natural[lane];
    s_nat[lane + 32] = c_jpeg_natural[lane + 32];
    if (lane == 0) s_rc = DSMIL_JPEG_OK;
  }
  __syncwarp();
  if (!jpeg_usable(s_hdr, a)) {
    if (lane == 0) a.status[img] = s_hdr.status != DSMIL_JPEG_OK ? s_hdr.status : DSMIL_JPEG_UNSUPPORTED;
    return;
  }
  // ---- all lanes: unstuff the segment -------------------------------------------------------------------
  const uint8_t* s = a.blob + s_hdr.file_off + s_hdr.scan_off;
  uint8_t* un = a.unstuffed + ((s_hdr.file_off + s_hdr.scan_off + 15) & ~15ll);
  const int len = s_hdr.scan_len;
  int ulen = 0;
  uint32_t carry = 0;                                   // the byte before this group of 32
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    const uint32_t b = i < len ? s[i] : 0u;
    uint32_t prev = __shfl_up_sync(0xffffffffu, b, 1);
    uint32_t next = __shfl_down_sync(0xffffffffu, b, 1);
    if (lane == 0) prev = carry;
    if (lane == 31) next = i + 1 < len ? s[i + 1] : 0u;
    const bool rst = b >= 0xD0u && b <= 0xD7u, nrst = next >= 0xD0u && next <= 0xD7u;
    const bool keep = i < len && !(b == 0u && prev == 0xFFu) && !(b == 0xFFu && nrst) && !(rst && prev == 0xFFu);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) un[ulen + __popc(m & ((1u << lane) - 1u))] = static_cast<uint8_t>(b);
    ulen += __popc(m);
    carry = __shfl_sync(0xffffffffu, b, 31);
  }
  if (lane < 12) un[ulen + lane] = 0;                   // zero tail: the bit reader loads whole words
  // ---- lanes 0..7: one Huffman table each -----------------------------------------------------------------
  if (lane < 8 && ((s_hdr.h_present >> lane) & 1))
    if (dsmil_jpeg_build_htab(s_hdr.hbits[lane], s_hdr.hvals[lane], &s_tab[lane]) != DSMIL_JPEG_OK) s_rc = DSMIL_JPEG_CORRUPT;
  __syncwarp();
  // ---- lane 0: the serial part -----------------------------------------------------------------------------
  if (lane == 0) {
    int rc = s_rc;
    if (rc == DSMIL_JPEG_OK) {
      int16_t* coef[3];
      for (int c = 0; c < 3; ++c) coef[c] = a.coef + (static_cast<long long>(img) * 3 + c) * a.plane_elems;
      rc = dsmil_jpeg_decode_scan(&s_hdr, un, static_cast<uint32_t>(ulen), s_tab, s_nat, coef);
    }
    a.status[img] = rc;
  }
}

// thread per 8x8 block; grid.y = patch
__global__ void __launch_bounds__(128)
k_jpeg_idct(JpegBatch a) {
  const int img = blockIdx.y;
  const dsmil_jpeg_header& h = a.hdr[img];
  if (!jpeg_usable(h, a)) return;
  int b = blockIdx.x * 128 + threadIdx.x;
  int c = 0;
  for (; c < h.ncomp; ++c) {
    const int nb = dsmil_jpeg_comp_bw(&h, c) * dsmil_jpeg_comp_bh(&h, c);
    if (b < nb) break;
    b -= nb;
  }
  if (c >= h.ncomp) return;
  const int bw = dsmil_jpeg_comp_bw(&h, c);
  const long long base = (static_cast<long long>(img) * 3 + c) * a.plane_elems;
  int16_t cf[64];
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.coef + base + 64ll * b);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 v = __ldg(src + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cf[8 * i + 2 * j] = static_cast<int16_t>(w[j] & 0xFFFFu);
        cf[8 * i + 2 * j + 1] = static_cast<int16_t>(w[j] >> 16);
      }
    }
  }
  uint16_t qt[64];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(h.qt[h.comp[c].tq]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 v = __ldg(q4 + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        qt[8 * i + 2 * j] = static_cast<uint16_t>(w[j] & 0xFFFFu);
        qt[8 * i + 2 * j + 1] = static_cast<uint16_t>(w[j] >> 16);
      }
    }
  }
  uint8_t px[64];
  dsmil_jpeg_idct_block(cf, qt, px, 8);
  const int by = b / bw, bx = b - by * bw, stride = bw * 8;
  uint8_t* dst = a.planes + base + static_cast<long long>(by) * 8 * stride + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint2 v;
    v.x = px[8 * r] | (px[8 * r + 1] << 8) | (px[8 * r + 2] << 16) | (static_cast<uint32_t>(px[8 * r + 3]) << 24);
    v.y = px[8 * r + 4] | (px[8 * r + 5] << 8) | (px[8 * r + 6] << 16) | (static_cast<uint32_t>(px[8 * r + 7]) << 24);
    *reinterpret_cast<uint2*>(dst + r * stride) = v;
  }
}

// thread per 4 horizontally adjacent output pixels; grid.y = patch
__global__ void __launch_bounds__(256)
k_jpeg_color(JpegBatch a) {
  const int img = blockIdx.y;
  const dsmil_jpeg_header& h = a.hdr[img];
  if (!jpeg_usable(h, a)) return;
  const int W4 = (a.W + 3) >> 2;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= W4 * a.H) return;
  const int y = t / W4, x0 = (t - y * W4) * 4;
  const int nx = min(4, a.W - x0);
  const long long base = static_cast<long long>(img) * 3 * a.plane_elems;
  const uint8_t* py = a.planes + base;
  const int sy = dsmil_jpeg_comp_bw(&h, 0) * 8;
  uint8_t rgb[4][3];
  if (h.ncomp == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint8_t v = i < nx ? py[y * sy + x0 + i] : 0;
      rgb[i][0] = rgb[i][1] = rgb[i][2] = v;
    }
  } else {
    const uint8_t* pcb = a.planes + base + a.plane_elems;
    const uint8_t* pcr = a.planes + base + 2 * a.plane_elems;
    const int sc = dsmil_jpeg_comp_bw(&h, 1) * 8;
    const int dw = (h.width + h.hmax - 1) / h.hmax, dh = (h.height + h.vmax - 1) / h.vmax;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < nx) {
        const int Y = py[y * sy + x0 + i];
        const int cb = dsmil_jpeg_upsample(pcb, sc, dw, dh, h.hmax, h.vmax, y, x0 + i);
        const int cr = dsmil_jpeg_upsample(pcr, sc, dw, dh, h.hmax, h.vmax, y, x0 + i);
        dsmil_jpeg_ycc_to_rgb(Y, cb, cr, &rgb[i][0], &rgb[i][1], &rgb[i][2]);
      } else {
        rgb[i][0] = rgb[i][1] = rgb[i][2] = 0;
      }
    }
  }
  if (a.out_u8) {
    uint8_t* o = a.out_u8 + (static_cast<long long>(img) * a.H * a.W + static_cast<long long>(y) * a.W + x0) * 3;
    if (nx == 4 && (a.W & 3) == 0) {                    // 12 bytes, 4-byte aligned when W % 4 == 0
      uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
      o32[0] = rgb[0][0] | (rgb[0][1] << 8) | (rgb[0][2] << 16) | (static_cast<uint32_t>(rgb[1][0]) << 24);
      o32[1] = rgb[1][1] | (rgb[1][2] << 8) | (rgb[2][0] << 16) | (static_cast<uint32_t>(rgb[2][1]) << 24);
      o32[2] = rgb[2][2] | (rgb[3][0] << 8) | (rgb[3][1] << 16) | (static_cast<uint32_t>(rgb[3][2]) << 24);
    } else {
      for (int i = 0; i < nx; ++i) { o[3 * i] = rgb[i][0]; o[3 * i + 1] = rgb[i][1]; o[3 * i + 2] = rgb[i][2]; }
    }
  }
  if (a.out_f32 && a.f32_hwc) {
    float* o = a.out_f32 + (static_cast<long long>(img) * a.H * a.W + static_cast<long long>(y) * a.W + x0) * 3;
    float f[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) f[3 * i + ch] = __fdiv_rn(static_cast<float>(rgb[i][ch]), 255.f);
    if (nx == 4 && (a.W & 3) == 0) {                    // 48 bytes, 16-byte aligned when W % 4 == 0
      float4* o4 = reinterpret_cast<float4*>(o);
      o4[0] = make_float4(f[0], f[1], f[2], f[3]);
      o4[1] = make_float4(f[4], f[5], f[6], f[7]);
      o4[2] = make_float4(f[8], f[9], f[10], f[11]);
    } else {
      for (int i = 0; i < 3 * nx; ++i) o[i] = f[i];
    }
  } else if (a.out_f32) {
    const long long hw = static_cast<long long>(a.H) * a.W;
    float* o = a.out_f32 + static_cast<long long>(img) * 3 * hw + static_cast<long long>(y) * a.W + x0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (nx == 4 && (a.W & 3) == 0) {
        float4 v;
        v.x = __fdiv_rn(static_cast<float>(rgb[0][ch]), 255.f); v.y = __fdiv_rn(static_cast<float>(rgb[1][ch]), 255.f);
        v.z = __fdiv_rn(static_cast<float>(rgb[2][ch]), 255.f); v.w = __fdiv_rn(static_cast<float>(rgb[3][ch]), 255.f);
        *reinterpret_cast<float4*>(o + ch * hw) = v;
      } else {
        for (int i = 0; i < nx; ++i) o[ch * hw + i] = __fdiv_rn(static_cast<float>(rgb[i][ch]), 255.f);
      }
    }
  }
}

inline long long jpeg_plane_elems(int H, int W) {
  return static_cast<long long>((H + 15) & ~15) * ((W + 15) & ~15);
}

inline int launch_jpeg_decode(const JpegBatch& a, cudaStream_t st) {
  DSMIL_CUDA_OK(cudaMemsetAsync(a.coef, 0, sizeof(int16_t) * 3 * a.plane_elems * a.n, st));
  k_jpeg_entropy<<<a.n, 32, 0, st>>>(a);
  DSMIL_LAUNCH_OK("k_jpeg_entropy");
  const int max_blocks = static_cast<int>(3 * a.plane_elems / 64);
  k_jpeg_idct<<<dim3((max_blocks + 127) / 128, a.n), 128, 0, st>>>(a);
  DSMIL_LAUNCH_OK("k_jpeg_idct");
  const int groups = ((a.W + 3) / 4) * a.H;
  k_jpeg_color<<<dim3((groups + 255) / 256, a.n), 256, 0, st>>>(a);
  DSMIL_LAUNCH_OK("k_jpeg_color");
  return 0;
}

}  // namespace dsmil
