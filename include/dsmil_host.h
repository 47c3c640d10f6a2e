/* libdsmil_host.so -- host-only helpers around the DSMIL hot path (plain C, no CUDA).
 *
 * Bag feature CSV of the reference (writer compute_feats.py:80-82,123-125: DataFrame.to_csv(index=False,
 * float_format='%.4f'); reader train_tcga.py:24-26: pd.read_csv -> float32):
 *   header line "0,1,...,D-1", one line per instance, '\n' line ends, NaN as an empty field.
 * Error returns are negative: -1 bad argument, -2 I/O, -3 ragged row, -4 not a number, -5 buffer too small.
 * Implementation and exactness notes: dsmil_wsi_b200/csrc_host/bagcsv.c.  Python binding: dsmil_wsi_b200/_hostlib.py.
 */
#ifndef DSMIL_HOST_H_
#define DSMIL_HOST_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DSMIL_HOST_ABI_VERSION 1
int32_t dsmil_host_abi_version(void);

/* Text of a bag [N, D] (fp32, row-major) into out[cap]; returns the byte count.  cap >= 12*D + 49*N*D + 16 always
 * suffices.  Byte-identical to the pandas call above for every float32 value. */
int64_t dsmil_csv_format_bag(const float* x, int64_t N, int32_t D, char* out, int64_t cap);
/* The same text written to `path` (buffered).  Returns the bytes written. */
int64_t dsmil_csv_write_bag(const char* path, const float* x, int64_t N, int32_t D);
/* Shape of a CSV held in memory: D = fields of the header line, N = non-blank data lines.  Returns 0. */
int32_t dsmil_csv_shape(const char* buf, int64_t len, int64_t* N, int32_t* D);
/* Values of the data lines into out[N*D] (N, D from dsmil_csv_shape): float32(correctly rounded double of each
 * field); empty field = NaN.  On error *bad_line holds the 1-based data-line number.  Returns 0. */
int32_t dsmil_csv_parse_bag(const char* buf, int64_t len, float* out, int64_t N, int32_t D, int64_t* bad_line);

/* Host half of the device JPEG loader (reference: compute_feats.py:26-29 `Image.open`; device half:
 * dsmil_jpeg_decode_batch in include/dsmil_b200.h).  Parses the marker segments of JPEG files into fixed-size
 * header records (dsmil_jpeg_header_bytes() each; layout in dsmil_wsi_b200/csrc/jpeg_core.h: geometry, sampling,
 * quantisation and Huffman tables, where the entropy-coded segment lies).  No pixel work on the host.
 * Status of a file: 0 decodable on the device, -1 corrupt, -2 valid JPEG outside the device path (progressive,
 * arithmetic, 12-bit, CMYK / Adobe colour, other samplings, multi-scan).
 * dsmil_jpeg_parse_batch: n files back to back in `blob`, file i = [offsets[i], offsets[i+1]) (offsets has n+1
 * entries); fills headers[0..n) and returns how many files are NOT decodable on the device (< 0: bad argument). */
int32_t dsmil_jpeg_header_bytes(void);
int32_t dsmil_jpeg_parse(const uint8_t* file, int64_t len, void* header);
int32_t dsmil_jpeg_parse_batch(const uint8_t* blob, const int64_t* offsets, int32_t n, void* headers);

/* File reader of the loader: n files (NUL-terminated paths) straight into one caller buffer (normally pinned, the
 * source of the H2D copy).  dsmil_files_offsets: offsets[0..n] = prefix sums of the file sizes; dsmil_files_read:
 * file i -> blob[offsets[i], offsets[i+1]) with `threads` (1..16) reader threads.  Return 0, or -(i+1) when file i
 * cannot be stat'ed / read completely. */
int64_t dsmil_files_offsets(const char* const* paths, int32_t n, int64_t* offsets);
int64_t dsmil_files_read(const char* const* paths, int32_t n, const int64_t* offsets, uint8_t* blob, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* DSMIL_HOST_H_ */
