/* Baseline JPEG decoding, the arithmetic of the patch loader of the embedding loop (reference
 * compute_feats.py:26-29: `Image.open(path)` -> `VF.to_tensor`, i.e. PIL's libjpeg(-turbo) decoder with its defaults:
 * JDCT_ISLOW inverse DCT, "fancy" (triangle) chroma upsampling, the 16-bit fixed-point YCbCr -> RGB tables).
 *
 * libjpeg itself is not part of /root/reference (it is PIL's third-party dependency; the image ships
 * libjpeg-turbo, jpeglib 6.2 API), so this file restates the PUBLISHED algorithms -- ITU-T T.81 Annex F for the
 * Huffman entropy coding, Loeffler-Ligtenberg-Moschytz for the 13-bit fixed-point 8x8 IDCT, the JFIF colour
 * transform -- with the exact rounding of the IJG implementation, and parity is pinned against PIL's own output on
 * the same files (tests/test_jpeg_host.py, tests/test_zz_jpeg_gpu.py: bit-exact).
 *
 * Plain C subset: the same inline functions compile as C11 (csrc_host/jpegparse.c, the header parser of the host
 * library), as C++ (oracle/jpeg_host_check.c, the CPU-side checker of this arithmetic) and as CUDA
 * (jpeg_kernels.cuh, where one warp decodes the entropy-coded segment of one patch and the IDCT / upsampling /
 * colour stages run data-parallel).
 */
#ifndef DSMIL_JPEG_CORE_H_
#define DSMIL_JPEG_CORE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define JPEG_FN __host__ __device__ __forceinline__
#else
#define JPEG_FN static inline
#endif

/* status codes of the parser / decoder */
#define DSMIL_JPEG_OK 0
#define DSMIL_JPEG_CORRUPT (-1)       /* truncated file, bad marker length, bad Huffman code ... */
#define DSMIL_JPEG_UNSUPPORTED (-2)   /* progressive / arithmetic / 12-bit / CMYK / Adobe-RGB / exotic sampling / multi-scan */

typedef struct {
  uint8_t id, h, v, tq, td, ta, pad0, pad1;
} dsmil_jpeg_comp;

/* One parsed file.  Everything the device needs besides the entropy-coded bytes themselves. */
typedef struct {
  int64_t file_off;          /* byte offset of the file inside the batch blob (filled by the batch parser) */
  int32_t width, height, ncomp;
  int32_t hmax, vmax;        /* sampling of component 0 (components 1, 2 are 1x1) */
  int32_t mcux, mcuy;        /* MCUs per row / column */
  int32_t restart_interval;  /* MCUs between RSTn markers, 0 = none */
  int32_t scan_off, scan_len;/* entropy-coded segment, relative to the start of the file, up to the closing marker */
  int32_t status;
  int32_t pad;
  dsmil_jpeg_comp comp[3];
  uint16_t qt[4][64];        /* quantisation tables in NATURAL (row-major) order */
  uint8_t hbits[8][16];      /* Huffman tables: index = class * 4 + id (class 0 = DC, 1 = AC); BITS list ... */
  uint8_t hvals[8][256];     /* ... and HUFFVAL list, as in the DHT segment */
  uint8_t qt_present, h_present, pad2[14];   /* sizeof == 2784, a multiple of 16: qt rows stay 16-byte aligned in arrays */
} dsmil_jpeg_header;

/* zig-zag position k -> natural (row-major) index; callers keep a copy where their memory model wants it */
#define DSMIL_JPEG_NATURAL_ORDER                                                                                   \
  {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, \
   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,  \
   54, 47, 55, 62, 63}

/* ------------------------------------------------------------------ header parser (T.81 Annex B) ------------ */

JPEG_FN int dsmil_jpeg_be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

/* Parses the marker segments of one file up to and including SOS and locates the end of the entropy-coded
 * segment.  `natural` = the zig-zag table above. */
JPEG_FN int dsmil_jpeg_parse_header(const uint8_t* f, int64_t len, const uint8_t* natural, dsmil_jpeg_header* h) {
  int64_t p = 2;
  int have_sof = 0, adobe = 0;
  int i, j;
  h->restart_interval = 0;
  h->qt_present = 0;
  h->h_present = 0;
  h->ncomp = 0;
  h->status = DSMIL_JPEG_CORRUPT;
  if (len < 4 || f[0] != 0xFF || f[1] != 0xD8) return DSMIL_JPEG_CORRUPT;
  for (;;) {
    int m, L;
    const uint8_t* s;
    if (p + 4 > len) return DSMIL_JPEG_CORRUPT;
    if (f[p] != 0xFF) return DSMIL_JPEG_CORRUPT;
    while (p < len && f[p] == 0xFF) ++p;             /* fill bytes before a marker */
    if (p >= len) return DSMIL_JPEG_CORRUPT;
    m = f[p++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;   /* stand-alone markers */
    if (m == 0xD9) return DSMIL_JPEG_CORRUPT;                           /* EOI before SOS */
    if (p + 2 > len) return DSMIL_JPEG_CORRUPT;
    L = dsmil_jpeg_be16(f + p);
    if (L < 2 || p + L > len) return DSMIL_JPEG_CORRUPT;
    s = f + p + 2;
    if (m == 0xC0 || m == 0xC1) {                    /* SOF0 baseline / SOF1 extended sequential, Huffman */
      if (L < 8 || s[0] != 8) return (h->status = DSMIL_JPEG_UNSUPPORTED);
      h->height = dsmil_jpeg_be16(s + 1);
      h->width = dsmil_jpeg_be16(s + 3);
      h->ncomp = s[5];
      if (h->height == 0 || h->width == 0) return (h->status = DSMIL_JPEG_UNSUPPORTED);   /* DNL-defined height */
      if (h->ncomp != 1 && h->ncomp != 3) return (h->status = DSMIL_JPEG_UNSUPPORTED);
      if (L < 8 + 3 * h->ncomp) return DSMIL_JPEG_CORRUPT;
      for (i = 0; i < h->ncomp; ++i) {
        h->comp[i].id = s[6 + 3 * i];
        h->comp[i].h = (uint8_t)(s[7 + 3 * i] >> 4);
        h->comp[i].v = (uint8_t)(s[7 + 3 * i] & 15);
        h->comp[i].tq = s[8 + 3 * i];
        if (h->comp[i].tq > 3 || h->comp[i].h == 0 || h->comp[i].v == 0) return DSMIL_JPEG_CORRUPT;
      }
      have_sof = 1;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return (h->status = DSMIL_JPEG_UNSUPPORTED);   /* progressive, lossless, arithmetic, hierarchical */
    } else if (m == 0xCC) {
      return (h->status = DSMIL_JPEG_UNSUPPORTED);   /* DAC */
    } else if (m == 0xDB) {                          /* DQT */
      int q = 0;
      while (q < L - 2) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        if (tq > 3 || pq > 1) return DSMIL_JPEG_CORRUPT;
        if (q + 1 + 64 * (pq + 1) > L - 2) return DSMIL_JPEG_CORRUPT;
        for (j = 0; j < 64; ++j)
          h->qt[tq][natural[j]] = (uint16_t)(pq ? dsmil_jpeg_be16(s + q + 1 + 2 * j) : s[q + 1 + j]);
        h->qt_present |= (uint8_t)(1 << tq);
        q += 1 + 64 * (pq + 1);
      }
    } else if (m == 0xC4) {                          /* DHT */
      int q = 0;
      while (q < L - 2) {
        const int tc = s[q] >> 4, th = s[q] & 15;
        int total = 0, t;
        if (tc > 1 || th > 3 || q + 17 > L - 2) return DSMIL_JPEG_CORRUPT;
        t = tc * 4 + th;
        for (j = 0; j < 16; ++j) {
          h->hbits[t][j] = s[q + 1 + j];
          total += s[q + 1 + j];
        }
        if (total > 256 || q + 17 + total > L - 2) return DSMIL_JPEG_CORRUPT;
        for (j = 0; j < total; ++j) h->hvals[t][j] = s[q + 17 + j];
        for (; j < 256; ++j) h->hvals[t][j] = 0;
        h->h_present |= (uint8_t)(1 << t);
        q += 17 + total;
      }
    } else if (m == 0xDD) {                          /* DRI */
      if (L != 4) return DSMIL_JPEG_CORRUPT;
      h->restart_interval = dsmil_jpeg_be16(s);
    } else if (m == 0xEE) {                          /* APP14: an Adobe marker changes the colour transform */
      if (L >= 14 && s[0] == 'A' && s[1] == 'd' && s[2] == 'o' && s[3] == 'b' && s[4] == 'e') adobe = 1;
    } else if (m == 0xDA) {                          /* SOS */
      int ns;
      int64_t e;
      if (!have_sof) return DSMIL_JPEG_CORRUPT;
      ns = s[0];
      if (ns != h->ncomp) return (h->status = DSMIL_JPEG_UNSUPPORTED);       /* non-interleaved multi-scan file */
      if (L != 6 + 2 * ns) return DSMIL_JPEG_CORRUPT;
      for (i = 0; i < ns; ++i) {
        if (s[1 + 2 * i] != h->comp[i].id) return (h->status = DSMIL_JPEG_UNSUPPORTED);
        h->comp[i].td = (uint8_t)(s[2 + 2 * i] >> 4);
        h->comp[i].ta = (uint8_t)(s[2 + 2 * i] & 15);
        if (h->comp[i].td > 3 || h->comp[i].ta > 3) return DSMIL_JPEG_CORRUPT;
        if (!((h->h_present >> h->comp[i].td) & 1) || !((h->h_present >> (4 + h->comp[i].ta)) & 1)) return DSMIL_JPEG_CORRUPT;
        if (!((h->qt_present >> h->comp[i].tq) & 1)) return DSMIL_JPEG_CORRUPT;
      }
      if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return (h->status = DSMIL_JPEG_UNSUPPORTED);
      if (adobe) return (h->status = DSMIL_JPEG_UNSUPPORTED);
      if (h->ncomp == 3) {
        if (h->comp[1].h != 1 || h->comp[1].v != 1 || h->comp[2].h != 1 || h->comp[2].v != 1) return (h->status = DSMIL_JPEG_UNSUPPORTED);
        if (!((h->comp[0].h == 1 && h->comp[0].v == 1) || (h->comp[0].h == 2 && h->comp[0].v == 1) ||
              (h->comp[0].h == 2 && h->comp[0].v == 2)))
          return (h->status = DSMIL_JPEG_UNSUPPORTED);
        if (h->comp[0].id == 'R' && h->comp[1].id == 'G' && h->comp[2].id == 'B') return (h->status = DSMIL_JPEG_UNSUPPORTED);
        h->hmax = h->comp[0].h;
        h->vmax = h->comp[0].v;
      } else {
        h->comp[0].h = h->comp[0].v = 1;             /* a single-component scan is never interleaved (A.2.2) */
        h->hmax = h->vmax = 1;
      }
      h->mcux = (h->width + 8 * h->hmax - 1) / (8 * h->hmax);
      h->mcuy = (h->height + 8 * h->vmax - 1) / (8 * h->vmax);
      p += L;
      h->scan_off = (int32_t)p;
      for (e = p; e + 1 < len; ++e)                  /* the segment ends at the first marker that is not RSTn */
        if (f[e] == 0xFF && f[e + 1] != 0x00 && !(f[e + 1] >= 0xD0 && f[e + 1] <= 0xD7) && f[e + 1] != 0xFF) break;
      if (e + 1 >= len) e = len;                     /* no EOI: libjpeg decodes what is there (and warns) */
      h->scan_len = (int32_t)(e - p);
      h->status = DSMIL_JPEG_OK;
      return DSMIL_JPEG_OK;
    }
    p += L;
  }
}

/* ------------------------------------------------------------------ Huffman tables (T.81 Annex C, F.2.2.3) --- */

#define DSMIL_JPEG_LOOK 9
typedef struct {
  uint16_t fast[1 << DSMIL_JPEG_LOOK];   /* next 9 bits -> (code length << 8) | symbol, 0 = longer than 9 bits */
  int32_t maxcode[18];                   /* largest code of each length (-1: none), [17] = sentinel */
  int32_t valoff[17];                    /* index of the first symbol of each length minus its first code */
  uint8_t vals[256];
} dsmil_jpeg_htab;

JPEG_FN int dsmil_jpeg_build_htab(const uint8_t* bits, const uint8_t* vals, dsmil_jpeg_htab* t) {
  int code = 0, k = 0, l, i;
  for (i = 0; i < (1 << DSMIL_JPEG_LOOK); ++i) t->fast[i] = 0;
  for (i = 0; i < 256; ++i) t->vals[i] = vals[i];
  for (l = 1; l <= 16; ++l) {
    const int n = bits[l - 1];
    t->valoff[l] = k - code;
    if (n) {
      if (code + n > (1 << l)) return DSMIL_JPEG_CORRUPT;
      if (l <= DSMIL_JPEG_LOOK) {
        for (i = 0; i < n; ++i) {
          const int first = (code + i) << (DSMIL_JPEG_LOOK - l), cnt = 1 << (DSMIL_JPEG_LOOK - l);
          int j;
          for (j = 0; j < cnt; ++j) t->fast[first + j] = (uint16_t)((l << 8) | vals[k + i]);
        }
      }
      code += n;
      k += n;
      t->maxcode[l] = code - 1;
    } else {
      t->maxcode[l] = -1;
    }
    code <<= 1;
  }
  t->maxcode[17] = 0x7FFFFFFF;
  t->maxcode[0] = -1;
  return DSMIL_JPEG_OK;
}

/* ------------------------------------------------------------------ bit reader over the UNSTUFFED segment ---- */
/* The segment has been copied with the stuffed zero of every FF 00 and every RSTn marker removed, starts 4-byte
 * aligned and is followed by >= 8 zero bytes; `acc` holds the next bits left-aligned. */
typedef struct {
  const uint8_t* p;
  uint32_t pos, len;      /* next byte to load (multiple of 4), number of valid bytes */
  uint64_t acc;
  int32_t nbits;
} dsmil_jpeg_bits;

JPEG_FN void dsmil_jpeg_bits_init(dsmil_jpeg_bits* b, const uint8_t* p, uint32_t len) {
  b->p = p; b->pos = 0; b->len = len; b->acc = 0; b->nbits = 0;
}

JPEG_FN void dsmil_jpeg_refill(dsmil_jpeg_bits* b) {      /* afterwards nbits >= 32 */
  if (b->nbits < 32) {
    uint32_t w = 0;
    if (b->pos < b->len + 4) {                             /* the zero padding makes the last partial word safe */
      const uint32_t raw = *(const uint32_t*)(b->p + b->pos);
#ifdef __CUDA_ARCH__
      w = __byte_perm(raw, 0, 0x0123);
#else
      w = (raw >> 24) | ((raw >> 8) & 0xFF00u) | ((raw << 8) & 0xFF0000u) | (raw << 24);
#endif
    }
    b->pos += 4;
    b->acc |= (uint64_t)w << (32 - b->nbits);
    b->nbits += 32;
  }
}

JPEG_FN int dsmil_jpeg_decode_sym(dsmil_jpeg_bits* b, const dsmil_jpeg_htab* t) {   /* needs nbits >= 16 */
  const uint32_t e = t->fast[b->acc >> (64 - DSMIL_JPEG_LOOK)];
  if (e) {
    b->acc <<= (e >> 8);
    b->nbits -= (int32_t)(e >> 8);
    return (int)(e & 255);
  } else {
    const int32_t c16 = (int32_t)(b->acc >> 48);
    int l = DSMIL_JPEG_LOOK + 1;
    while (l <= 16 && (c16 >> (16 - l)) > t->maxcode[l]) ++l;
    if (l > 16) return -1;
    b->acc <<= l;
    b->nbits -= l;
    return t->vals[((c16 >> (16 - l)) + t->valoff[l]) & 255];
  }
}

JPEG_FN int dsmil_jpeg_receive_extend(dsmil_jpeg_bits* b, int s) {   /* F.2.2.1 EXTEND(RECEIVE(s), s); 1 <= s <= 16 */
  const int v = (int)(b->acc >> (64 - s));
  b->acc <<= s;
  b->nbits -= s;
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

/* One 8x8 block: DC difference + run/size coded AC coefficients, written (quantised) into the zeroed `blk` in
 * natural order.  Returns 0 or DSMIL_JPEG_CORRUPT. */
JPEG_FN int dsmil_jpeg_decode_block(dsmil_jpeg_bits* b, const dsmil_jpeg_htab* dc, const dsmil_jpeg_htab* ac,
                                    const uint8_t* natural, int* pred, int16_t* blk) {
  int s, k;
  dsmil_jpeg_refill(b);
  s = dsmil_jpeg_decode_sym(b, dc);
  if (s < 0 || s > 16) return DSMIL_JPEG_CORRUPT;
  if (s) *pred = (int)((unsigned)*pred + (unsigned)dsmil_jpeg_receive_extend(b, s));
  blk[0] = (int16_t)*pred;
  for (k = 1; k < 64;) {
    int rs, r;
    dsmil_jpeg_refill(b);
    rs = dsmil_jpeg_decode_sym(b, ac);
    if (rs < 0) return DSMIL_JPEG_CORRUPT;
    r = rs >> 4;
    s = rs & 15;
    if (s) {
      k += r;
      if (k > 63) return DSMIL_JPEG_CORRUPT;
      blk[natural[k]] = (int16_t)dsmil_jpeg_receive_extend(b, s);
      ++k;
    } else if (r == 15) {
      k += 16;
    } else {
      break;                                               /* EOB */
    }
  }
  return DSMIL_JPEG_OK;
}

/* byte-align at a restart boundary: the encoder padded with 1-bits, the marker itself is already removed */
JPEG_FN void dsmil_jpeg_bits_align(dsmil_jpeg_bits* b) {
  const int drop = b->nbits & 7;
  b->acc <<= drop;
  b->nbits -= drop;
}

/* Where component c keeps its data.  Coefficients: blocks of 64 int16, block-row-major over the component's
 * (mcuy * v) x (mcux * h) blocks; planes: uint8, (mcuy * v * 8) rows of (mcux * h * 8) samples. */
JPEG_FN int dsmil_jpeg_comp_bw(const dsmil_jpeg_header* h, int c) { return h->mcux * h->comp[c].h; }
JPEG_FN int dsmil_jpeg_comp_bh(const dsmil_jpeg_header* h, int c) { return h->mcuy * h->comp[c].v; }

/* The whole entropy-coded segment of one file, serially (interleaved MCUs, T.81 A.2.3).  `coef[c]` = the zeroed
 * coefficient array of component c. */
JPEG_FN int dsmil_jpeg_decode_scan(const dsmil_jpeg_header* h, const uint8_t* unstuffed, uint32_t ulen,
                                   const dsmil_jpeg_htab* tabs /* [0..3] DC ids, [4..7] AC ids as used */,
                                   const uint8_t* natural, int16_t* const* coef) {
  dsmil_jpeg_bits b;
  int pred[3] = {0, 0, 0};
  int mx, my, c, bx, by, left = h->restart_interval;
  dsmil_jpeg_bits_init(&b, unstuffed, ulen);
  for (my = 0; my < h->mcuy; ++my) {
    for (mx = 0; mx < h->mcux; ++mx) {
      if (h->restart_interval) {
        if (left == 0) {
          dsmil_jpeg_bits_align(&b);
          pred[0] = pred[1] = pred[2] = 0;
          left = h->restart_interval;
        }
        --left;
      }
      for (c = 0; c < h->ncomp; ++c) {
        const int bw = dsmil_jpeg_comp_bw(h, c);
        const dsmil_jpeg_htab* dc = tabs + h->comp[c].td;
        const dsmil_jpeg_htab* ac = tabs + 4 + h->comp[c].ta;
        for (by = 0; by < h->comp[c].v; ++by)
          for (bx = 0; bx < h->comp[c].h; ++bx) {
            int16_t* blk = coef[c] + 64 * ((int64_t)(my * h->comp[c].v + by) * bw + (mx * h->comp[c].h + bx));
            if (dsmil_jpeg_decode_block(&b, dc, ac, natural, &pred[c], blk) != DSMIL_JPEG_OK) return DSMIL_JPEG_CORRUPT;
          }
      }
    }
  }
  return DSMIL_JPEG_OK;
}

/* ------------------------------------------------------------------ inverse DCT (JDCT_ISLOW) ---------------- */
/* Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits kept between the passes: the IJG "slow-but-
 * accurate integer" IDCT that libjpeg(-turbo) uses by default, 32-bit arithmetic as its SIMD builds. */
#define JPEG_CONST_BITS 13
#define JPEG_PASS1_BITS 2
#define JPEG_FIX_0_298631336 2446
#define JPEG_FIX_0_390180644 3196
#define JPEG_FIX_0_541196100 4433
#define JPEG_FIX_0_765366865 6270
#define JPEG_FIX_0_899976223 7373
#define JPEG_FIX_1_175875602 9633
#define JPEG_FIX_1_501321110 12299
#define JPEG_FIX_1_847759065 15137
#define JPEG_FIX_1_961570560 16069
#define JPEG_FIX_2_053119869 16819
#define JPEG_FIX_2_562915447 20995
#define JPEG_FIX_3_072711026 25172

/* 32-bit wrap-around arithmetic, spelled out so that it is defined behaviour for ANY coefficient values (damaged files
 * produce huge ones); valid files never wrap and libjpeg-turbo's SIMD IDCT wraps the same way. */
JPEG_FN int32_t jw_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
JPEG_FN int32_t jw_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
JPEG_FN int32_t jw_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
JPEG_FN int32_t jw_shl(int32_t a, int n) { return (int32_t)((uint32_t)a << n); }

JPEG_FN int32_t dsmil_jpeg_descale(int32_t x, int n) { return jw_add(x, 1 << (n - 1)) >> n; }

/* the IDCT's output stage: +128 and the 10-bit wrap-around range-limit table of the IJG code */
JPEG_FN uint8_t dsmil_jpeg_range_limit(int32_t x) {
  const int32_t i = x & 1023;
  return (uint8_t)(i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896)));
}

/* 1-D kernel on 8 values with stride-free arguments; `shift` = the descale of this pass */
#define DSMIL_JPEG_IDCT_1D(i0, i1, i2, i3, i4, i5, i6, i7, o0, o1, o2, o3, o4, o5, o6, o7, shift)                 \
  {                                                                                                                \
    int32_t z1, z2, z3, z4, z5, t0, t1, t2, t3, t10, t11, t12, t13;                                                \
    z2 = (i2); z3 = (i6);                                                                                          \
    z1 = jw_mul(jw_add(z2, z3), JPEG_FIX_0_541196100);                                                             \
    t2 = jw_add(z1, jw_mul(z3, -JPEG_FIX_1_847759065));                                                            \
    t3 = jw_add(z1, jw_mul(z2, JPEG_FIX_0_765366865));                                                             \
    z2 = (i0); z3 = (i4);                                                                                          \
    t0 = jw_shl(jw_add(z2, z3), JPEG_CONST_BITS);                                                                  \
    t1 = jw_shl(jw_sub(z2, z3), JPEG_CONST_BITS);                                                                  \
    t10 = jw_add(t0, t3); t13 = jw_sub(t0, t3); t11 = jw_add(t1, t2); t12 = jw_sub(t1, t2);                        \
    t0 = (i7); t1 = (i5); t2 = (i3); t3 = (i1);                                                                    \
    z1 = jw_add(t0, t3); z2 = jw_add(t1, t2); z3 = jw_add(t0, t2); z4 = jw_add(t1, t3);                            \
    z5 = jw_mul(jw_add(z3, z4), JPEG_FIX_1_175875602);                                                             \
    t0 = jw_mul(t0, JPEG_FIX_0_298631336); t1 = jw_mul(t1, JPEG_FIX_2_053119869);                                  \
    t2 = jw_mul(t2, JPEG_FIX_3_072711026); t3 = jw_mul(t3, JPEG_FIX_1_501321110);                                  \
    z1 = jw_mul(z1, -JPEG_FIX_0_899976223); z2 = jw_mul(z2, -JPEG_FIX_2_562915447);                                \
    z3 = jw_mul(z3, -JPEG_FIX_1_961570560); z4 = jw_mul(z4, -JPEG_FIX_0_390180644);                                \
    z3 = jw_add(z3, z5); z4 = jw_add(z4, z5);                                                                      \
    t0 = jw_add(t0, jw_add(z1, z3)); t1 = jw_add(t1, jw_add(z2, z4));                                              \
    t2 = jw_add(t2, jw_add(z2, z3)); t3 = jw_add(t3, jw_add(z1, z4));                                              \
    o0 = dsmil_jpeg_descale(jw_add(t10, t3), shift); o7 = dsmil_jpeg_descale(jw_sub(t10, t3), shift);              \
    o1 = dsmil_jpeg_descale(jw_add(t11, t2), shift); o6 = dsmil_jpeg_descale(jw_sub(t11, t2), shift);              \
    o2 = dsmil_jpeg_descale(jw_add(t12, t1), shift); o5 = dsmil_jpeg_descale(jw_sub(t12, t1), shift);              \
    o3 = dsmil_jpeg_descale(jw_add(t13, t0), shift); o4 = dsmil_jpeg_descale(jw_sub(t13, t0), shift);              \
  }

/* one block: dequantise, columns then rows, 8 rows of 8 samples written at out[r * stride + c] */
JPEG_FN void dsmil_jpeg_idct_block(const int16_t* coef, const uint16_t* qt, uint8_t* out, int stride) {
  int32_t ws[64];
  int c, r;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (c = 0; c < 8; ++c) {
    DSMIL_JPEG_IDCT_1D(coef[c] * (int32_t)qt[c], coef[8 + c] * (int32_t)qt[8 + c], coef[16 + c] * (int32_t)qt[16 + c],
                       coef[24 + c] * (int32_t)qt[24 + c], coef[32 + c] * (int32_t)qt[32 + c],
                       coef[40 + c] * (int32_t)qt[40 + c], coef[48 + c] * (int32_t)qt[48 + c],
                       coef[56 + c] * (int32_t)qt[56 + c], ws[c], ws[8 + c], ws[16 + c], ws[24 + c], ws[32 + c],
                       ws[40 + c], ws[48 + c], ws[56 + c], JPEG_CONST_BITS - JPEG_PASS1_BITS)
  }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (r = 0; r < 8; ++r) {
    int32_t o0, o1, o2, o3, o4, o5, o6, o7;
    DSMIL_JPEG_IDCT_1D(ws[8 * r], ws[8 * r + 1], ws[8 * r + 2], ws[8 * r + 3], ws[8 * r + 4], ws[8 * r + 5],
                       ws[8 * r + 6], ws[8 * r + 7], o0, o1, o2, o3, o4, o5, o6, o7,
                       JPEG_CONST_BITS + JPEG_PASS1_BITS + 3)
    out[r * stride + 0] = dsmil_jpeg_range_limit(o0); out[r * stride + 1] = dsmil_jpeg_range_limit(o1);
    out[r * stride + 2] = dsmil_jpeg_range_limit(o2); out[r * stride + 3] = dsmil_jpeg_range_limit(o3);
    out[r * stride + 4] = dsmil_jpeg_range_limit(o4); out[r * stride + 5] = dsmil_jpeg_range_limit(o5);
    out[r * stride + 6] = dsmil_jpeg_range_limit(o6); out[r * stride + 7] = dsmil_jpeg_range_limit(o7);
  }
}

/* ------------------------------------------------------------------ chroma upsampling + colour -------------- */
/* One chroma sample at full resolution, position (oy, ox), from the plane of a component subsampled by
 * (hmax, vmax) in {1x1, 2x1, 2x2}: libjpeg's "fancy" triangle filters with their alternating rounding; the rows
 * above the first / below the last real row and the columns left / right of the plane replicate the edge.
 * dw, dh = ceil(width / hmax), ceil(height / vmax) (the component's downsampled size); planes narrower than three
 * samples are replicated instead, as libjpeg does. */
JPEG_FN int dsmil_jpeg_upsample(const uint8_t* plane, int stride, int dw, int dh, int hmax, int vmax, int oy, int ox) {
  if (hmax == 1) return plane[oy * stride + ox];
  if (vmax == 1) {                                   /* h2v1 */
    const uint8_t* row = plane + oy * stride;
    const int col = ox >> 1, cur = row[col];
    if (dw <= 2) return cur;
    if (ox & 1) return col == dw - 1 ? cur : (cur * 3 + row[col + 1] + 2) >> 2;
    return col == 0 ? cur : (cur * 3 + row[col - 1] + 1) >> 2;
  } else {                                           /* h2v2 */
    const int inrow = oy >> 1, col = ox >> 1;
    int other = (oy & 1) ? inrow + 1 : inrow - 1;
    const uint8_t *r0, *r1;
    int cur;
    if (dw <= 2) return plane[inrow * stride + col];
    other = other < 0 ? 0 : (other > dh - 1 ? dh - 1 : other);
    r0 = plane + inrow * stride;
    r1 = plane + other * stride;
    cur = 3 * r0[col] + r1[col];
    if (ox & 1) return col == dw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + 3 * r0[col + 1] + r1[col + 1] + 7) >> 4;
    return col == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + 3 * r0[col - 1] + r1[col - 1] + 8) >> 4;
  }
}

JPEG_FN int dsmil_jpeg_clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

/* JFIF YCbCr -> RGB with the IJG 16-bit fixed-point tables: FIX(1.40200) = 91881, FIX(1.77200) = 116130,
 * FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768 */
JPEG_FN void dsmil_jpeg_ycc_to_rgb(int y, int cb, int cr, uint8_t* r, uint8_t* g, uint8_t* b) {
  const int u = cb - 128, v = cr - 128;
  *r = (uint8_t)dsmil_jpeg_clamp255(y + ((91881 * v + 32768) >> 16));
  *g = (uint8_t)dsmil_jpeg_clamp255(y + ((-22554 * u + 32768 - 46802 * v) >> 16));
  *b = (uint8_t)dsmil_jpeg_clamp255(y + ((116130 * u + 32768) >> 16));
}

#endif  /* DSMIL_JPEG_CORE_H_ */
