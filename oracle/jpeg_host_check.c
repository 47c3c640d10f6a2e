/* TEST INFRASTRUCTURE, not product code: a serial CPU build of the arithmetic in
 * dsmil_wsi_b200/csrc/jpeg_core.h (entropy decoding, IDCT, upsampling, colour), so that its bit-exactness against
 * PIL / libjpeg-turbo -- the decoder the reference calls at compute_feats.py:28 -- can be checked without a GPU
 * (tests/test_jpeg_host.py).  Only tests/ load it; the product decodes on the device (jpeg_kernels.cuh) and has no
 * CPU path.  The oracle for the JPEG loader is PIL itself. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../dsmil_wsi_b200/csrc/jpeg_core.h"

static const uint8_t k_natural[64] = DSMIL_JPEG_NATURAL_ORDER;

/* file -> RGB uint8 HWC (grey files: the sample replicated, as PIL's convert("RGB")).  `rgb` holds
 * height * width * 3 bytes of the size reported by jpegcheck_size.  Returns the status code. */
int jpegcheck_size(const uint8_t* file, int64_t len, int32_t* w, int32_t* h, int32_t* ncomp) {
  dsmil_jpeg_header hd;
  const int rc = dsmil_jpeg_parse_header(file, len, k_natural, &hd);
  if (rc == DSMIL_JPEG_OK) { *w = hd.width; *h = hd.height; *ncomp = hd.ncomp; }
  return rc;
}

int jpegcheck_decode(const uint8_t* file, int64_t len, uint8_t* rgb) {
  dsmil_jpeg_header hd;
  dsmil_jpeg_htab* tabs;
  int16_t* coef[3] = {0, 0, 0};
  uint8_t* plane[3] = {0, 0, 0};
  uint8_t* un;
  uint32_t ulen = 0;
  int rc = dsmil_jpeg_parse_header(file, len, k_natural, &hd), c, t;
  int64_t i;
  if (rc != DSMIL_JPEG_OK) return rc;
  /* unstuff: drop the 00 of FF 00 and RSTn markers */
  un = (uint8_t*)calloc((size_t)hd.scan_len + 16, 1);
  {
    const uint8_t* s = file + hd.scan_off;
    for (i = 0; i < hd.scan_len; ++i) {
      const int prev = i ? s[i - 1] : 0, next = i + 1 < hd.scan_len ? s[i + 1] : 0;
      if (s[i] == 0x00 && prev == 0xFF) continue;
      if (s[i] == 0xFF && next >= 0xD0 && next <= 0xD7) continue;
      if (s[i] >= 0xD0 && s[i] <= 0xD7 && prev == 0xFF) continue;
      un[ulen++] = s[i];
    }
  }
  tabs = (dsmil_jpeg_htab*)calloc(8, sizeof(dsmil_jpeg_htab));
  for (t = 0; t < 8; ++t)
    if ((hd.h_present >> t) & 1)
      if (dsmil_jpeg_build_htab(hd.hbits[t], hd.hvals[t], &tabs[t]) != DSMIL_JPEG_OK) rc = DSMIL_JPEG_CORRUPT;
  for (c = 0; c < hd.ncomp; ++c) {
    const size_t nb = (size_t)dsmil_jpeg_comp_bw(&hd, c) * dsmil_jpeg_comp_bh(&hd, c);
    coef[c] = (int16_t*)calloc(nb * 64, sizeof(int16_t));
    plane[c] = (uint8_t*)calloc(nb * 64, 1);
  }
  if (rc == DSMIL_JPEG_OK) rc = dsmil_jpeg_decode_scan(&hd, un, ulen, tabs, k_natural, coef);
  if (rc == DSMIL_JPEG_OK) {
    int x, y;
    for (c = 0; c < hd.ncomp; ++c) {
      const int bw = dsmil_jpeg_comp_bw(&hd, c), bh = dsmil_jpeg_comp_bh(&hd, c), stride = bw * 8;
      int bx, by;
      for (by = 0; by < bh; ++by)
        for (bx = 0; bx < bw; ++bx)
          dsmil_jpeg_idct_block(coef[c] + 64 * ((size_t)by * bw + bx), hd.qt[hd.comp[c].tq],
                                plane[c] + (size_t)by * 8 * stride + bx * 8, stride);
    }
    for (y = 0; y < hd.height; ++y)
      for (x = 0; x < hd.width; ++x) {
        uint8_t* o = rgb + 3 * ((size_t)y * hd.width + x);
        const int Y = plane[0][(size_t)y * dsmil_jpeg_comp_bw(&hd, 0) * 8 + x];
        if (hd.ncomp == 1) {
          o[0] = o[1] = o[2] = (uint8_t)Y;
        } else {
          const int dw = (hd.width + hd.hmax - 1) / hd.hmax, dh = (hd.height + hd.vmax - 1) / hd.vmax;
          const int cb = dsmil_jpeg_upsample(plane[1], dsmil_jpeg_comp_bw(&hd, 1) * 8, dw, dh, hd.hmax, hd.vmax, y, x);
          const int cr = dsmil_jpeg_upsample(plane[2], dsmil_jpeg_comp_bw(&hd, 2) * 8, dw, dh, hd.hmax, hd.vmax, y, x);
          dsmil_jpeg_ycc_to_rgb(Y, cb, cr, o, o + 1, o + 2);
        }
      }
  }
  for (c = 0; c < 3; ++c) { free(coef[c]); free(plane[c]); }
  free(tabs);
  free(un);
  return rc;
}
