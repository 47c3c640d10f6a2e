/* Host side of the GPU JPEG loader (reference compute_feats.py:26-29 `Image.open`): parses the marker
 * segments of every patch file of a batch into the fixed-size header the device kernels read
 * (csrc/jpeg_core.h: tables, geometry, where the entropy-coded segment lies) and says which files the device path
 * decodes (baseline / extended-sequential Huffman, 8-bit, grey or YCbCr 4:4:4 / 4:2:2 / 4:2:0, one interleaved
 * scan).  No pixel work happens here. */
#include <stddef.h>
#include <stdint.h>

#include "../csrc/jpeg_core.h"

static const uint8_t k_natural[64] = DSMIL_JPEG_NATURAL_ORDER;

int32_t dsmil_jpeg_header_bytes(void) { return (int32_t)sizeof(dsmil_jpeg_header); }

/* One file.  Returns DSMIL_JPEG_OK / _CORRUPT / _UNSUPPORTED (also stored in out->status). */
int32_t dsmil_jpeg_parse(const uint8_t* file, int64_t len, void* header) {
  dsmil_jpeg_header* out = (dsmil_jpeg_header*)header;
  int rc;
  if (!file || !out || len < 0) return -1;
  out->file_off = 0;
  rc = dsmil_jpeg_parse_header(file, len, k_natural, out);
  out->status = rc;
  return rc;
}

/* n files stored back to back in `blob`, file i = [offsets[i], offsets[i+1]).  Fills out[0..n) (file_off set) and
 * returns the number of files the device path can NOT take (0 = the whole batch is decodable on the device). */
int32_t dsmil_jpeg_parse_batch(const uint8_t* blob, const int64_t* offsets, int32_t n, void* headers) {
  dsmil_jpeg_header* out = (dsmil_jpeg_header*)headers;
  int32_t i, bad = 0;
  if (!blob || !offsets || !out || n < 0) return -1;
  for (i = 0; i < n; ++i) {
    const int64_t len = offsets[i + 1] - offsets[i];
    int rc = len >= 0 ? dsmil_jpeg_parse_header(blob + offsets[i], len, k_natural, &out[i]) : DSMIL_JPEG_CORRUPT;
    out[i].status = rc;
    out[i].file_off = offsets[i];
    if (rc != DSMIL_JPEG_OK) ++bad;
  }
  return bad;
}
