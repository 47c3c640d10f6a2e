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

/* ---- file reader of the loader: the patch files of a batch go straight from the page cache / disk into the (pinned)
 * blob the H2D copy reads from -- no per-file Python objects, no intermediate copies (the reference's workers do
 * `Image.open(path)` per file, compute_feats.py:26-28). ---------------------------------------------------------- */
#include <fcntl.h>
#include <pthread.h>
#include <sys/stat.h>
#include <unistd.h>

/* offsets[0..n] = prefix sums of the file sizes.  Returns 0, or -(i+1) if file i cannot be stat'ed. */
int64_t dsmil_files_offsets(const char* const* paths, int32_t n, int64_t* offsets) {
  int32_t i;
  if (!paths || !offsets || n < 0) return -1;
  offsets[0] = 0;
  for (i = 0; i < n; ++i) {
    struct stat sb;
    if (stat(paths[i], &sb) != 0 || !S_ISREG(sb.st_mode)) return -(int64_t)(i + 1);
    offsets[i + 1] = offsets[i] + (int64_t)sb.st_size;
  }
  return 0;
}

typedef struct {
  const char* const* paths;
  const int64_t* offsets;
  uint8_t* blob;
  int32_t n, first, step, failed;
} dsmil_read_job;

static void* dsmil_read_worker(void* arg) {
  dsmil_read_job* j = (dsmil_read_job*)arg;
  int32_t i;
  for (i = j->first; i < j->n; i += j->step) {
    const int64_t want = j->offsets[i + 1] - j->offsets[i];
    int64_t got = 0;
    const int fd = open(j->paths[i], O_RDONLY);
    if (fd < 0) { j->failed = i + 1; return NULL; }
    while (got < want) {
      const ssize_t r = read(fd, j->blob + j->offsets[i] + got, (size_t)(want - got));
      if (r <= 0) break;
      got += r;
    }
    close(fd);
    if (got != want) { j->failed = i + 1; return NULL; }
  }
  return NULL;
}

/* Reads file i into blob[offsets[i], offsets[i+1]) with `threads` reader threads (1..16).  Returns 0, or -(i+1) for
 * a file that could not be read completely (e.g. it changed size since dsmil_files_offsets). */
int64_t dsmil_files_read(const char* const* paths, int32_t n, const int64_t* offsets, uint8_t* blob, int32_t threads) {
  pthread_t tid[16];
  dsmil_read_job job[16];
  int32_t t, started = 0;
  int64_t rc = 0;
  if (!paths || !offsets || !blob || n < 0) return -1;
  if (threads < 1) threads = 1;
  if (threads > 16) threads = 16;
  if (threads > n) threads = n > 0 ? n : 1;
  for (t = 0; t < threads; ++t) {
    job[t].paths = paths; job[t].offsets = offsets; job[t].blob = blob;
    job[t].n = n; job[t].first = t; job[t].step = threads; job[t].failed = 0;
  }
  for (t = 1; t < threads; ++t) {
    if (pthread_create(&tid[t], NULL, dsmil_read_worker, &job[t]) != 0) break;
    started = t;
  }
  for (t = started + 1; t < threads; ++t) {       /* threads that could not start: their share runs here */
    dsmil_read_worker(&job[t]);
  }
  dsmil_read_worker(&job[0]);
  for (t = 1; t <= started; ++t) pthread_join(tid[t], NULL);
  for (t = 0; t < threads; ++t)
    if (job[t].failed && (rc == 0 || -(int64_t)job[t].failed > rc)) rc = -(int64_t)job[t].failed;
  return rc;
}
