/* Fuzz harness (test tooling) for dsmil_wsi_b200/csrc/jpeg_core.h: random byte mutations / truncations of a seed
 * JPEG through the CPU build of the parser + decoder, under -fsanitize=address,undefined.
 *   gcc -O1 -g -fsanitize=address,undefined -o /tmp/jpeg_fuzz tools/jpeg_fuzz.c oracle/jpeg_host_check.c && /tmp/jpeg_fuzz seed.jpg 100000
 * The same inline functions run on the device, where an out-of-bounds access would take the CUDA context down. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int jpegcheck_size(const uint8_t* file, int64_t len, int32_t* w, int32_t* h, int32_t* ncomp);
int jpegcheck_decode(const uint8_t* file, int64_t len, uint8_t* rgb);

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 32);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s seed.jpg iterations\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* seed = (uint8_t*)malloc((size_t)n);
  if (fread(seed, 1, (size_t)n, f) != (size_t)n) return 2;
  fclose(f);
  long iters = atol(argv[2]), ok = 0, corrupt = 0, unsup = 0, skipped = 0;
  for (long it = 0; it < iters; ++it) {
    long len = n;
    if (rnd() % 4 == 0) len = 1 + rnd() % n;                  /* truncation */
    uint8_t* buf = (uint8_t*)malloc((size_t)len);             /* exact size: ASAN sees any read past the end */
    memcpy(buf, seed, (size_t)len);
    int muts = rnd() % 6;
    int header_bias = rnd() % 2;
    for (int m = 0; m < muts; ++m) {
      long span = header_bias ? (len < 700 ? len : 700) : len;
      long pos = rnd() % span;
      switch (rnd() % 4) {
        case 0: buf[pos] = (uint8_t)rnd(); break;
        case 1: buf[pos] ^= (uint8_t)(1u << (rnd() % 8)); break;
        case 2: buf[pos] = 0xFF; break;
        default: buf[pos] = 0x00; break;
      }
    }
    int32_t w = 0, h = 0, nc = 0;
    int rc = jpegcheck_size(buf, len, &w, &h, &nc);
    if (rc == 0) {
      if ((int64_t)w * h > 4096 * 4096) { ++skipped; free(buf); continue; }
      uint8_t* rgb = (uint8_t*)malloc((size_t)w * h * 3);
      rc = jpegcheck_decode(buf, len, rgb);
      free(rgb);
    }
    if (rc == 0) ++ok; else if (rc == -1) ++corrupt; else ++unsup;
    free(buf);
  }
  free(seed);
  printf("iterations %ld: decoded %ld, corrupt %ld, unsupported %ld, skipped(huge) %ld\n", iters, ok, corrupt, unsup, skipped);
  return 0;
}
