/* common.c -- CPU ORACLE (test infrastructure, not product code): RNG, input decoding,
 * painter's-algorithm raster, luma, INTER_AREA resize.  See tbo.h for the parity statement. */
#include "tbo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- ALE action ids, toybox/envs/atari/constants.py:16-35 */
int tbo_ale_action_to_input(int a) {
  static const int T[18] = {
    0, TBO_IN_BUTTON1, TBO_IN_UP, TBO_IN_RIGHT, TBO_IN_LEFT, TBO_IN_DOWN,
    TBO_IN_UP | TBO_IN_RIGHT, TBO_IN_UP | TBO_IN_LEFT, TBO_IN_DOWN | TBO_IN_RIGHT, TBO_IN_DOWN | TBO_IN_LEFT,
    TBO_IN_UP | TBO_IN_BUTTON1, TBO_IN_RIGHT | TBO_IN_BUTTON1, TBO_IN_LEFT | TBO_IN_BUTTON1, TBO_IN_DOWN | TBO_IN_BUTTON1,
    TBO_IN_UP | TBO_IN_RIGHT | TBO_IN_BUTTON1, TBO_IN_UP | TBO_IN_LEFT | TBO_IN_BUTTON1,
    TBO_IN_DOWN | TBO_IN_RIGHT | TBO_IN_BUTTON1, TBO_IN_DOWN | TBO_IN_LEFT | TBO_IN_BUTTON1 };
  if (a < 0 || a > 17) return -1;
  return T[a];
}

/* ---- xoroshiro128+ 55/14/36; KAT in SURVEY App. A.1, checked in tests/test_oracle_golden.py */
static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void tbo_rng_seed(tbo_rng *g, uint32_t seed) {
  /* [FIX] seed 13 gives the amidar/breakout default config rand, seed 17 (advanced 2 draws) the
   * space-invaders lineage: the constants are the classic unseeded-XorShift words. */
  g->s[0] = 0x193a6754a8a7d469ULL ^ (uint64_t)seed;
  g->s[1] = 0x97830e05113ba7bbULL;
}
uint64_t tbo_rng_next_u64(tbo_rng *g) {
  uint64_t s0 = g->s[0], s1 = g->s[1], r = s0 + s1;
  s1 ^= s0;
  g->s[0] = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
  g->s[1] = rotl64(s1, 36);
  return r;
}
uint32_t tbo_rng_next_u32(tbo_rng *g) { return (uint32_t)(tbo_rng_next_u64(g) >> 32); }
tbo_rng tbo_rng_child(tbo_rng *p) { tbo_rng c; c.s[0] = tbo_rng_next_u64(p); c.s[1] = tbo_rng_next_u64(p); return c; }
uint32_t tbo_rng_index(tbo_rng *g, uint32_t n) {
  /* [FIX] breakout fixture: child rng consumed exactly 2 draws choosing start index 2 of 4:
   * first draw rejected by the zone test, second accepted. */
  uint32_t zone;
  int lz = 0;
  if (n == 0) return 0;
  while (!((n << lz) & 0x80000000u)) lz++;
  zone = (n << lz) - 1u;
  for (;;) {
    uint64_t m = (uint64_t)tbo_rng_next_u32(g) * (uint64_t)n;
    if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
  }
}
double tbo_rng_f64(tbo_rng *g) { return (double)(tbo_rng_next_u64(g) >> 11) * (1.0 / 9007199254740992.0); }

uint32_t tbo_action_index(uint64_t seed, uint64_t env, uint64_t t, uint32_t n_legal) {
  uint64_t z = seed + env * 0x9E3779B97F4A7C15ULL + t * 0xD1B54A32D192ED03ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (uint32_t)(((z >> 32) * (uint64_t)n_legal) >> 32);
}

/* ---- raster */
uint8_t tbo_luma(tbo_color c) {
  double v = 0.299 * (double)c.r + 0.587 * (double)c.g + 0.114 * (double)c.b;
  return (uint8_t)v;
}
void tbo_clear(tbo_canvas *cv, tbo_color c) { tbo_rect(cv, c, 0, 0, cv->w, cv->h); }
void tbo_rect(tbo_canvas *cv, tbo_color c, int x, int y, int w, int h) {
  int x0 = x < 0 ? 0 : x, y0 = y < 0 ? 0 : y;
  int x1 = x + w > cv->w ? cv->w : x + w, y1 = y + h > cv->h ? cv->h : y + h;
  for (int yy = y0; yy < y1; yy++)
    for (int xx = x0; xx < x1; xx++) {
      uint8_t *p = cv->rgba + 4 * ((size_t)yy * cv->w + xx);
      p[0] = c.r; p[1] = c.g; p[2] = c.b; p[3] = c.a;
    }
}
void tbo_sprite1(tbo_canvas *cv, tbo_color c, int x, int y, int w, int h, const uint32_t *rows, int sx, int sy) {
  for (int r = 0; r < h; r++)
    for (int q = 0; q < w; q++)
      if ((rows[r] >> (w - 1 - q)) & 1u) tbo_rect(cv, c, x + q * sx, y + r * sy, sx, sy);
}
const uint32_t TBO_FONT3X5[10][5] = {
  {7, 5, 5, 5, 7}, {2, 6, 2, 2, 7}, {7, 1, 7, 4, 7}, {7, 1, 7, 1, 7}, {5, 5, 7, 1, 1},
  {7, 4, 7, 1, 7}, {7, 4, 7, 5, 7}, {7, 1, 1, 1, 1}, {7, 5, 7, 5, 7}, {7, 5, 7, 1, 7} };
void tbo_digits(tbo_canvas *cv, tbo_color c, int x_right, int y, int value, int sx, int sy) {
  int v = value < 0 ? 0 : value;
  int x = x_right - 3 * sx;
  do {
    tbo_sprite1(cv, c, x, y, 3, 5, TBO_FONT3X5[v % 10], sx, sy);
    v /= 10; x -= 4 * sx;
  } while (v > 0);
}
void tbo_rgba_to_gray(const uint8_t *rgba, int npix, uint8_t *gray) {
  for (int i = 0; i < npix; i++) { tbo_color c = { rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], 255 }; gray[i] = tbo_luma(c); }
}
void tbo_rgba_to_rgb(const uint8_t *rgba, int npix, uint8_t *rgb) {
  for (int i = 0; i < npix; i++) { rgb[3 * i] = rgba[4 * i]; rgb[3 * i + 1] = rgba[4 * i + 1]; rgb[3 * i + 2] = rgba[4 * i + 2]; }
}

/* ---- INTER_AREA (general, non-integer-scale path): weight table then separable accumulate in float,
 * exactly in the order OpenCV's ResizeArea_ does it; result rounded half-to-even and saturated. */
typedef struct { int si, di; float alpha; } area_tap;
static int area_tab(int ssize, int dsize, int cn, double scale, area_tap *tab) {
  int k = 0;
  for (int dx = 0; dx < dsize; dx++) {
    double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    double cell = scale < (double)ssize - fsx1 ? scale : (double)ssize - fsx1;
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    if (sx1 - fsx1 > 1e-3) { tab[k].di = dx * cn; tab[k].si = (sx1 - 1) * cn; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
    for (int sx = sx1; sx < sx2; sx++) { tab[k].di = dx * cn; tab[k].si = sx * cn; tab[k++].alpha = (float)(1.0 / cell); }
    if (fsx2 - sx2 > 1e-3) {
      double m = fsx2 - sx2; if (m > 1.0) m = 1.0; if (m > cell) m = cell;
      tab[k].di = dx * cn; tab[k].si = sx2 * cn; tab[k++].alpha = (float)(m / cell);
    }
  }
  return k;
}
static uint8_t sat_u8(float v) {
  long r = lrintf(v);        /* default rounding mode: half to even, as cvRound */
  return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}
void tbo_resize_area_u8(const uint8_t *src, int sw, int sh, int cn, uint8_t *dst, int dw, int dh) {
  double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
  double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
  area_tap *xtab = (area_tap *)malloc(sizeof(area_tap) * (size_t)sw * 2);
  area_tap *ytab = (area_tap *)malloc(sizeof(area_tap) * (size_t)sh * 2);
  int nx = area_tab(sw, dw, cn, scale_x, xtab), ny = area_tab(sh, dh, 1, scale_y, ytab);
  int row = dw * cn;
  float *buf = (float *)malloc(sizeof(float) * (size_t)row * 2), *sum = buf + row;
  int prev_dy = ytab[0].di;
  for (int i = 0; i < row; i++) sum[i] = 0.f;
  for (int j = 0; j < ny; j++) {
    float beta = ytab[j].alpha; int dy = ytab[j].di; const uint8_t *S = src + (size_t)ytab[j].si * sw * cn;
    for (int i = 0; i < row; i++) buf[i] = 0.f;
    for (int k = 0; k < nx; k++)
      for (int ch = 0; ch < cn; ch++) {
        float prod = (float)S[xtab[k].si + ch] * xtab[k].alpha;
        buf[xtab[k].di + ch] = buf[xtab[k].di + ch] + prod;
      }
    if (dy != prev_dy) {
      uint8_t *D = dst + (size_t)prev_dy * row;
      for (int i = 0; i < row; i++) { D[i] = sat_u8(sum[i]); sum[i] = beta * buf[i]; }
      prev_dy = dy;
    } else {
      for (int i = 0; i < row; i++) { float p = beta * buf[i]; sum[i] = sum[i] + p; }
    }
  }
  { uint8_t *D = dst + (size_t)prev_dy * row; for (int i = 0; i < row; i++) D[i] = sat_u8(sum[i]); }
  free(buf); free(xtab); free(ytab);
}

size_t tbo_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(tbo_brk_cfg); case 1: return sizeof(tbo_brk_state);
    case 2: return sizeof(tbo_si_cfg);  case 3: return sizeof(tbo_si_state);
    case 4: return sizeof(tbo_ami_cfg); case 5: return sizeof(tbo_ami_state);
  }
  return 0;
}
