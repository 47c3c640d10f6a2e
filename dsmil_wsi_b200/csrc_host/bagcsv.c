/* Native reader / writer of the reference's bag feature CSV (compute_feats.py:80-82 -> train_tcga.py:24-26):
 * header "0,1,...,D-1", then one row per instance, values printed with '%.4f', no index column.
 *
 * Host-only C (no CUDA): the reference spends seconds per 10 000 x 512 bag in DataFrame.to_csv / read_csv; this
 * is the same bytes and the same parsed values in tens of milliseconds.
 *
 * Exactness:
 *   writer  -- identical text to Python's '%.4f' % float(v) for every float32 v (what pandas' float_format
 *              applies): the decimal is rounded from the EXACT binary value, ties to even, "-0.0000" keeps its
 *              sign, NaN is the empty field pandas writes, inf is "inf".  Done in integer arithmetic
 *              (m * 10^4 / 2^s with the remainder inspected), snprintf only for |v| >= 2^39.
 *   reader  -- float32(correctly rounded double of the decimal string) for fields of up to 18 significant digits
 *              without exponent (K / 10^frac with K exact in a double: one IEEE division = the correctly rounded
 *              value, i.e. Python's float()); anything else goes through strtod.  Empty field = NaN.  For the '%.4f'
 *              wire format this is bit-identical to the reference's pd.read_csv route (pinned on the committed
 *              fixtures); for other fraction lengths pandas' own parser may differ from float() by one ulp.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DSMIL_CSV_ERR_ARG (-1)
#define DSMIL_CSV_ERR_IO (-2)
#define DSMIL_CSV_ERR_RAGGED (-3)
#define DSMIL_CSV_ERR_NUMBER (-4)
#define DSMIL_CSV_ERR_SPACE (-5)

static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

/* ---- writer ------------------------------------------------------------------------------- */

static inline char* put_u64(char* p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

/* one value, '%.4f'; returns the advanced pointer (at most 48 bytes are written) */
static inline char* put_f4(char* p, float f) {
  uint32_t bits;
  memcpy(&bits, &f, 4);
  const uint32_t sign = bits >> 31, ex = (bits >> 23) & 0xffu, man = bits & 0x7fffffu;
  if (ex == 0xffu) {
    if (man) return p;                        /* NaN: pandas writes an empty field (na_rep='') */
    if (sign) *p++ = '-';
    memcpy(p, "inf", 3);
    return p + 3;
  }
  uint64_t m;
  int e;
  if (ex == 0) { m = man; e = -149; } else { m = man | 0x800000u; e = (int)ex - 150; }
  if (e > 15) {                               /* |v| >= 2^39: integer, but too wide for the 64-bit path */
    return p + snprintf(p, 48, "%.4f", (double)f);
  }
  uint64_t K;                                 /* round-half-even(|v| * 10^4) */
  const uint64_t P = m * 10000u;              /* < 2^38 */
  if (e >= 0) {
    K = P << e;                               /* < 2^53 */
  } else {
    const int s = -e;
    if (s >= 64) {
      K = 0;
    } else {
      uint64_t q = P >> s;
      const uint64_t r = P & ((UINT64_C(1) << s) - 1), half = UINT64_C(1) << (s - 1);
      if (r > half || (r == half && (q & 1))) ++q;
      K = q;
    }
  }
  if (sign) *p++ = '-';
  p = put_u64(p, K / 10000u);
  uint32_t frac = (uint32_t)(K % 10000u);
  *p++ = '.';
  p[3] = (char)('0' + frac % 10); frac /= 10;
  p[2] = (char)('0' + frac % 10); frac /= 10;
  p[1] = (char)('0' + frac % 10); frac /= 10;
  p[0] = (char)('0' + frac);
  return p + 4;
}

static char* put_header(char* p, int32_t D) {
  for (int32_t j = 0; j < D; ++j) {
    if (j) *p++ = ',';
    p = put_u64(p, (uint64_t)j);
  }
  *p++ = '\n';
  return p;
}

/* Text of the whole bag into `out` (capacity `cap`); returns the byte count or DSMIL_CSV_ERR_SPACE.
 * A capacity of 12*D + 48*N*D + 16 always suffices. */
int64_t dsmil_csv_format_bag(const float* x, int64_t N, int32_t D, char* out, int64_t cap) {
  if (!x || !out || N < 0 || D < 1) return DSMIL_CSV_ERR_ARG;
  char* p = out;
  char* const end = out + cap;
  if (end - p < 12ll * D + 2) return DSMIL_CSV_ERR_SPACE;
  p = put_header(p, D);
  for (int64_t n = 0; n < N; ++n) {
    if (end - p < 49ll * D + 2) return DSMIL_CSV_ERR_SPACE;
    const float* row = x + n * D;
    for (int32_t j = 0; j < D; ++j) {
      if (j) *p++ = ',';
      p = put_f4(p, row[j]);
    }
    *p++ = '\n';
  }
  return p - out;
}

/* Same text straight to a file through a 1 MiB buffer.  Returns bytes written or a negative error. */
int64_t dsmil_csv_write_bag(const char* path, const float* x, int64_t N, int32_t D) {
  if (!path || !x || N < 0 || D < 1) return DSMIL_CSV_ERR_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return DSMIL_CSV_ERR_IO;
  const size_t cap = (size_t)1 << 20;
  const size_t row_max = (size_t)49 * (size_t)D + 2;
  const size_t buf_len = cap > 2 * row_max ? cap : 2 * row_max;
  char* buf = (char*)malloc(buf_len > (size_t)12 * D + 2 ? buf_len : (size_t)12 * D + 2);
  if (!buf) { fclose(f); return DSMIL_CSV_ERR_IO; }
  int64_t total = 0;
  char* p = put_header(buf, D);
  for (int64_t n = 0; n < N; ++n) {
    if ((size_t)(p - buf) + row_max > buf_len) {
      if (fwrite(buf, 1, (size_t)(p - buf), f) != (size_t)(p - buf)) { free(buf); fclose(f); return DSMIL_CSV_ERR_IO; }
      total += p - buf;
      p = buf;
    }
    const float* row = x + n * D;
    for (int32_t j = 0; j < D; ++j) {
      if (j) *p++ = ',';
      p = put_f4(p, row[j]);
    }
    *p++ = '\n';
  }
  if (fwrite(buf, 1, (size_t)(p - buf), f) != (size_t)(p - buf)) { free(buf); fclose(f); return DSMIL_CSV_ERR_IO; }
  total += p - buf;
  free(buf);
  if (fclose(f) != 0) return DSMIL_CSV_ERR_IO;
  return total;
}

/* ---- reader ------------------------------------------------------------------------------- */

/* Columns of the header line and number of non-blank data lines.  Returns 0 or a negative error. */
int32_t dsmil_csv_shape(const char* buf, int64_t len, int64_t* N, int32_t* D) {
  if (!buf || len < 0 || !N || !D) return DSMIL_CSV_ERR_ARG;
  const char* const end = buf + len;
  const char* nl = (const char*)memchr(buf, '\n', (size_t)len);
  const char* hend = nl ? nl : end;
  if (hend == buf || (hend == buf + 1 && buf[0] == '\r')) return DSMIL_CSV_ERR_ARG;   /* no header */
  int32_t cols = 1;
  for (const char* p = buf; p < hend; ++p) cols += (*p == ',');
  *D = cols;
  int64_t rows = 0;
  const char* p = nl ? nl + 1 : end;
  while (p < end) {
    const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* e = q ? q : end;
    if (e > p && e[-1] == '\r') --e;
    if (e > p) ++rows;                                                       /* blank lines are skipped */
    p = q ? q + 1 : end;
  }
  *N = rows;
  return 0;
}

static inline int parse_field(const char* s, const char* e, float* out) {
  if (s == e) { *out = NAN; return 0; }                                      /* empty field: NaN, as pandas */
  const char* p = s;
  int neg = 0;
  if (*p == '-') { neg = 1; ++p; } else if (*p == '+') { ++p; }
  uint64_t K = 0;
  int digits = 0, frac = 0, seen_dot = 0, any = 0;
  const char* q = p;
  for (; q < e; ++q) {
    const char c = *q;
    if (c >= '0' && c <= '9') {
      any = 1;
      if (digits < 18) { K = K * 10 + (uint64_t)(c - '0'); if (K || digits) ++digits; if (seen_dot) ++frac; }
      else break;                                                            /* too long for the exact path */
    } else if (c == '.' && !seen_dot) {
      seen_dot = 1;
    } else {
      break;
    }
  }
  if (q == e && any && K < (UINT64_C(1) << 53) && frac <= 22) {
    const double d = (double)K / kPow10[frac];
    *out = (float)(neg ? -d : d);
    return 0;
  }
  /* general path: exponents, inf/nan spellings, very long mantissas */
  char tmp[64];
  const size_t n = (size_t)(e - s);
  if (n >= sizeof(tmp)) return DSMIL_CSV_ERR_NUMBER;
  memcpy(tmp, s, n);
  tmp[n] = 0;
  char* endp = NULL;
  const double d = strtod(tmp, &endp);
  if (endp == tmp || *endp != 0) return DSMIL_CSV_ERR_NUMBER;
  *out = (float)d;
  return 0;
}

/* Parses the data lines into out[N*D] (row-major).  N and D must come from dsmil_csv_shape.  On a ragged row
 * or a bad number the 1-based data-line number is stored in *bad_line. */
int32_t dsmil_csv_parse_bag(const char* buf, int64_t len, float* out, int64_t N, int32_t D, int64_t* bad_line) {
  if (!buf || len < 0 || (!out && N > 0) || N < 0 || D < 1) return DSMIL_CSV_ERR_ARG;
  const char* const end = buf + len;
  const char* p = (const char*)memchr(buf, '\n', (size_t)len);              /* skip the header */
  p = p ? p + 1 : end;
  int64_t row = 0;
#define DSMIL_CSV_FAIL(code, line) do { if (bad_line) *bad_line = (line); return (code); } while (0)
  while (p < end) {
    if (*p == '\n') { ++p; continue; }                                       /* blank line */
    if (*p == '\r' && p + 1 < end && p[1] == '\n') { p += 2; continue; }
    if (*p == '\r' && p + 1 == end) break;
    if (row >= N) DSMIL_CSV_FAIL(DSMIL_CSV_ERR_RAGGED, row + 1);
    float* dst = out + row * D;
    int32_t col = 0;
    for (;;) {                                                               /* one field per iteration */
      if (col >= D) DSMIL_CSV_FAIL(DSMIL_CSV_ERR_RAGGED, row + 1);
      /* fast path: [sign] digits [. digits] with at most 18 significant digits, ended by , \r \n or EOF */
      const char* q = p;
      int neg = 0;
      if (q < end && (*q == '-' || *q == '+')) { neg = (*q == '-'); ++q; }
      uint64_t K = 0;
      int digits = 0, frac = 0, any = 0;
      while (q < end && (unsigned)(*q - '0') < 10u && digits < 18) {
        K = K * 10 + (uint64_t)(*q - '0');
        digits += (K != 0 || digits != 0);
        any = 1;
        ++q;
      }
      if (q < end && *q == '.') {
        ++q;
        while (q < end && (unsigned)(*q - '0') < 10u && digits < 18) {
          K = K * 10 + (uint64_t)(*q - '0');
          digits += (K != 0 || digits != 0);
          ++frac;
          any = 1;
          ++q;
        }
      }
      const int at_end = (q == end) || *q == ',' || *q == '\n' || *q == '\r';
      if (at_end && any && K < (UINT64_C(1) << 53) && frac <= 22) {
        const double d = (double)K / kPow10[frac];
        dst[col] = (float)(neg ? -d : d);
      } else {                                                               /* general path on the whole field */
        const char* e = p;
        while (e < end && *e != ',' && *e != '\n' && *e != '\r') ++e;
        const int rc = parse_field(p, e, dst + col);
        if (rc) DSMIL_CSV_FAIL(rc, row + 1);
        q = e;
      }
      ++col;
      p = q;
      if (p < end && *p == ',') { ++p; continue; }
      break;                                                                 /* end of line or of file */
    }
    if (p < end && *p == '\r') ++p;
    if (p < end) {
      if (*p != '\n') DSMIL_CSV_FAIL(DSMIL_CSV_ERR_NUMBER, row + 1);
      ++p;
    }
    if (col != D) DSMIL_CSV_FAIL(DSMIL_CSV_ERR_RAGGED, row + 1);
    ++row;
  }
#undef DSMIL_CSV_FAIL
  if (row != N) { if (bad_line) *bad_line = row; return DSMIL_CSV_ERR_RAGGED; }
  return 0;
}

int32_t dsmil_host_abi_version(void) { return 1; }
