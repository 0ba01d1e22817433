/* tbx_render_native.cuh -- native layouts (RGBA / packed RGB / gray) as BROADCAST + PATCH, two launches.
 *
 * A frame of Toybox.get_state() / get_rgb_frame() (toybox/envs/atari/base.py:109,164) is the env's base frame (static
 * scenery, or the look of a fresh game) everywhere except under the few draw-list entries that differ from it.  The
 * canvas kernel of tbx_render.cuh repaints a shared-memory canvas per env and streams it out, so every env costs a
 * restore / build / paint / fence / wait round per band -- half of all stall samples of the sprite-heavy games sit in
 * the wait for the bulk store to have read the canvas.  Here:
 *   1. base_fill_kernel: a CTA loads one band of a base frame into shared memory ONCE and hands it to the TMA engine
 *      for 32 envs in a row (cp.async.bulk shared -> global, no per-env instructions, nothing to wait for between
 *      them): the frames leave at the speed of HBM.
 *   2. native_patch_kernel (stream-ordered after 1.): one WARP per env builds the entries that differ from the base,
 *      marks the 32 x 8 pixel tiles they touch, and for every run of marked tiles copies the base window (L2) into a
 *      small scratch, paints the hit entries in draw order (painter's algorithm, identical result) and overwrites the
 *      window's bytes in the frame.  Work and traffic are proportional to what differs from the base.
 * Same list / sweep / rebuild machinery as tbx_render_area.cuh, in native pixel space and in the output pixel format.
 * (Measured and dropped: the patches of env chunk c on a second stream while chunk c+1 is broadcast -- no gain
 * with 2, 4 or 8 chunks, even with the broadcast CTAs padded to 104 KB of shared memory so that both grids fit an SM; and both steps fused per (8 envs, band) CTA -- 100+ registers, the patches wait for the stores to
 * LAND and every band rebuilds the entries: Amidar RGB 24.7 M frames/s against 35.3 M for the two launches.)
 */
#ifndef TBX_RENDER_NATIVE_CUH
#define TBX_RENDER_NATIVE_CUH
#include "tbx_render_area.cuh"

namespace tbxk {

#define TBX_FILL_ENVS 32   /* envs per CTA of the broadcast kernel */
#define TBX_NT_TW 32       /* tile: 32 x 8 pixels */
#define TBX_NT_TH 8
#define TBX_NT_MAX_RUN(PIX) ((PIX) == 1 ? 4 : 2) /* tiles per run: the scratch holds 128 x 8 gray or 64 x 8 colour pixels */
#define TBX_NT_LCAP 128

/* ---- 1. broadcast: frame rows [r0, r1) of every env <- its base frame */
template <int GAME, int PIX>
__global__ void __launch_bounds__(128) base_fill_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ typename Traits<GAME>::Cfg cfg_c) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H;
  extern __shared__ uint4 smem_raw[];
  uint8_t *band = reinterpret_cast<uint8_t *>(smem_raw);
  __shared__ int base_of[TBX_FILL_ENVS];
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * TBX_FILL_ENVS, ne = min(TBX_FILL_ENVS, a.n - e0);
  const int r0 = blockIdx.y * a.band_rows, r1 = min(H, r0 + a.band_rows);
  if (tid < ne) { /* base_id looks at the header words only */
    uint32_t hdr[TBX_HDR_WORDS];
#pragma unroll
    for (int w = 0; w < TBX_HDR_WORDS; w++) hdr[w] = a.planes[(size_t)w * a.n_pad + e0 + tid];
    base_of[tid] = T::base_id(hdr, cfg_c, tables);
  }
  __syncthreads();
  const int first = base_of[0];
  load_canvas<PIX, W>(band, first ? a.base[1] : a.base[0], r0, r1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t nbytes = (uint32_t)((r1 - r0) * W * PIX);
  if (tid == 0) {
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(band);
    for (int j = 0; j < ne; j++) {
      if (base_of[j] != first) continue;
      uint8_t *out = a.dst + (size_t)(e0 + j) * a.frame_bytes + (size_t)r0 * W * PIX;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(saddr), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  /* envs on the other base frame (custom brick tables): plain global -> global copies */
  for (int j = 0; j < ne; j++) {
    if (base_of[j] == first) continue;
    const uint4 *src = reinterpret_cast<const uint4 *>((base_of[j] ? a.base[1] : a.base[0]) + (size_t)r0 * W * PIX);
    uint4 *dst = reinterpret_cast<uint4 *>(a.dst + (size_t)(e0 + j) * a.frame_bytes + (size_t)r0 * W * PIX);
    for (int i = tid; i < (int)(nbytes >> 4); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* the band must outlive the engine's reads */
}

/* ---- 2. patches */
/* a warp paints one entry into the scratch window [wx0,wx1) x [wy0,wy1) (PIX bytes per pixel, `stride` bytes per row) */
template <int PIX>
__device__ __forceinline__ void paint_window(uint8_t *scr, int stride, int wx0, int wy0, int wx1, int wy1, const uint4 &q, const uint32_t *rec, int lane) {
  const int qx = (int16_t)(q.x & 0xffffu), qy = (int16_t)(q.x >> 16), qw = (int16_t)(q.y & 0xffffu), qh = (int16_t)(q.y >> 16);
  const int x0 = max(qx, wx0), x1 = min(qx + qw, wx1), y0 = max(qy, wy0), y1 = min(qy + qh, wy1);
  const int nw = x1 - x0, nh = y1 - y0;
  if (nw <= 0 || nh <= 0) return;
  const int lg = nw > 16 ? 5 : nw > 8 ? 4 : nw > 4 ? 3 : nw > 2 ? 2 : nw > 1 ? 1 : 0;
  const int cpl = 1 << lg, rpp = 32 >> lg, sub = lane >> lg, cx = lane & (cpl - 1);
  const uint32_t val = q.z;
  const int bw = (q.w >> 16) & 255;
  if (bw == 0) {
    for (int yy = sub; yy < nh; yy += rpp) {
      uint8_t *row = scr + (y0 + yy - wy0) * stride;
      for (int xb = cx; xb < nw; xb += cpl) put_pixel<PIX>(row, (size_t)(x0 + xb - wx0), val);
    }
  } else {
    const uint32_t off = q.w & 0xffffu;
    const bool state = (off & TBX_PRIM_STATE) != 0;
    const int o = state ? (int)(off & 0x7fffu) : (int)off;
    const int sx = (q.w >> 24) & 15, sy = q.w >> 28;
    const uint32_t ix = d_inv16[sx], iy = d_inv16[sy];
    for (int yy = sub; yy < nh; yy += rpp) {
      const int py = y0 + yy - qy;
      const int sy_i = sy == 1 ? py : (int)(((uint32_t)py * iy) >> 16);
      const uint32_t bits = state ? rec[o + sy_i] : __ldg(&d_bank[o + sy_i]);
      uint8_t *row = scr + (y0 + yy - wy0) * stride;
      for (int xb = cx; xb < nw; xb += cpl) {
        const int px = x0 + xb - qx;
        const int sx_i = sx == 1 ? px : (int)(((uint32_t)px * ix) >> 16);
        if ((bits >> (bw - 1 - sx_i)) & 1u) put_pixel<PIX>(row, (size_t)(x0 + xb - wx0), val);
      }
    }
  }
}
/* one lane fills its own small solid rectangle */
template <int PIX>
__device__ __forceinline__ void paint_window_lane(uint8_t *scr, int stride, int wx0, int wy0, int wx1, int wy1, const uint4 &q) {
  const int qx = (int16_t)(q.x & 0xffffu), qy = (int16_t)(q.x >> 16), qw = (int16_t)(q.y & 0xffffu), qh = (int16_t)(q.y >> 16);
  const int x0 = max(qx, wx0) - wx0, x1 = min(qx + qw, wx1) - wx0, y0 = max(qy, wy0) - wy0, y1 = min(qy + qh, wy1) - wy0;
  if (x0 >= x1 || y0 >= y1) return;
  for (int y = y0; y < y1; y++) {
    uint8_t *row = scr + y * stride;
    for (int x = x0; x < x1; x++) put_pixel<PIX>(row, (size_t)x, q.z);
  }
}

/* entry: x: x | y << 16; y: w | h << 16; z: colour (RGBA, or the gray byte for PIX 1); w: sprite off | bw << 16 | scale << 24
 * ext (uint2): x: x0 | x1 << 16 (clipped pixel columns [x0, x1)), y: y0 | y1 << 10 | group << 20 | flags << 28 */
template <int W, int PIX>
__device__ __forceinline__ bool make_native_entry(const TbxPrim &p, int g, int gmode, int rA, int rB, uint4 &e, uint2 &ext) {
  Clip c;
  if (!clip_prim<W>(p, rA, rB, c)) return false;
  const uint32_t flags = ((gmode & TBX_GROUP_SERIAL) ? 0u : TBX_ENTRY_PAR) | ((p.bw == 0 && (c.x1 - c.x0) * (c.y1 - c.y0) <= 64) ? TBX_ENTRY_SMALL : 0u);
  e.x = (uint32_t)(uint16_t)p.x | ((uint32_t)(uint16_t)p.y << 16);
  e.y = (uint32_t)(uint16_t)p.w | ((uint32_t)(uint16_t)p.h << 16);
  e.z = PIX == 1 ? tbx_luma(p.color) : p.color;
  e.w = (uint32_t)p.off | ((uint32_t)p.bw << 16) | ((uint32_t)p.scale << 24);
  ext.x = (uint32_t)c.x0 | ((uint32_t)c.x1 << 16);
  ext.y = (uint32_t)c.y0 | ((uint32_t)c.y1 << 10) | ((uint32_t)g << 20) | (flags << 28);
  return true;
}
/* tiles are numbered 16 per tile row: bit (ty & 1) * 16 + tx of word ty >> 1 */
__device__ __forceinline__ void mark_native_tiles(uint32_t *tmask, const uint2 &ext) {
  const int x0 = ext.x & 0xffffu, x1 = ext.x >> 16, y0 = ext.y & 1023u, y1 = (ext.y >> 10) & 1023u;
  const int txlo = x0 / TBX_NT_TW, txhi = (x1 - 1) / TBX_NT_TW;
  const uint32_t cols = ((2u << (txhi - txlo)) - 1u) << txlo;
  for (int ty = y0 / TBX_NT_TH; ty <= (y1 - 1) / TBX_NT_TH; ty++) atomicOr(&tmask[ty >> 1], cols << ((ty & 1) * 16));
}

/* the patches of one env over tile rows [ty_lo, ty_hi): every lane of the warp calls it */
template <int GAME, int PIX>
__device__ __forceinline__ void patch_env(const RenderArgs &a, const uint32_t *R, const typename Traits<GAME>::Cfg &cfg, const typename Traits<GAME>::Table *tables,
                                          int ty_lo, int ty_hi, uint8_t *out, uint4 *list, uint2 *exts, uint32_t *tmask, uint8_t *scr, int lane) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H;
  constexpr int STRIDE = TBX_NT_MAX_RUN(PIX) * TBX_NT_TW * PIX;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int base = T::base_id(R, cfg, tables);
  const uint8_t *bfr = base ? a.base[1] : a.base[0];
  const int nty = ty_hi - ty_lo;
  int nsw = 1;
  for (int sw = 0; sw < nsw; sw++) {
    const int tyA = ty_lo + (sw * nty) / nsw, tyB = ty_lo + ((sw + 1) * nty) / nsw;
    if (tyA >= tyB) continue;
    const int rA = tyA * TBX_NT_TH, rB = min(H, tyB * TBX_NT_TH);
    __syncwarp();
    if (lane < 16) tmask[lane] = 0;
    __syncwarp();
    int n = 0;
    bool overflow = false;
    for (int g = 0; g < T::NG && !overflow; g++) {
      int gb, ge, gmode;
      T::group(g, R, tables, base, gb, ge, gmode);
      T::trim(g, R, cfg, base, gb, ge);
      for (int s0 = gb; s0 < ge; s0 += 32) {
        const int s = s0 + lane;
        TbxPrim p = tbx_prim_none();
        if (s < ge) p = T::prim(R, cfg, tables, s, base);
        uint4 e;
        uint2 ext;
        const bool ok = make_native_entry<W, PIX>(p, g, gmode, rA, rB, e, ext);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m == 0) continue;
        if (n + __popc(m) > a.list_cap) { overflow = true; break; }
        if (ok) {
          const int slot = n + __popc(m & lt_mask);
          list[slot] = e;
          exts[slot] = ext;
          mark_native_tiles(tmask, ext);
        }
        n += __popc(m);
      }
    }
    if (overflow && nsw < nty) { nsw = min(nty, nsw * 2); sw = -1; continue; }
    __syncwarp();
    if (overflow) {
      /* one tile row with more entries than the list holds: rebuild the primitives per tile, paint in draw order */
      const int wy0 = rA, wy1 = rB;
      for (int wx0 = 0; wx0 < W; wx0 += TBX_NT_TW) {
        const int wx1 = min(W, wx0 + TBX_NT_TW);
        const int nwr = ((wx1 - wx0) * PIX) >> 2;
        for (int y = 0; y < wy1 - wy0; y++)
          for (int c = lane; c < nwr; c += 32)
            reinterpret_cast<uint32_t *>(scr + y * STRIDE)[c] = __ldg(reinterpret_cast<const uint32_t *>(bfr + ((size_t)(wy0 + y) * W + wx0) * PIX) + c);
        __syncwarp();
        for (int g = 0; g < T::NG; g++) {
          int gb, ge, gmode;
          T::group(g, R, tables, base, gb, ge, gmode);
          T::trim(g, R, cfg, base, gb, ge);
          for (int s0 = gb; s0 < ge; s0 += 32) {
            const int s = s0 + lane;
            TbxPrim p = tbx_prim_none();
            if (s < ge) p = T::prim(R, cfg, tables, s, base);
            uint4 e;
            uint2 ext;
            bool hit = make_native_entry<W, PIX>(p, g, gmode, rA, rB, e, ext);
            hit = hit && (int)(ext.x & 0xffffu) < wx1 && (int)(ext.x >> 16) > wx0;
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
              const int l = __ffs(m) - 1;
              m &= m - 1;
              uint4 q;
              q.x = __shfl_sync(0xffffffffu, e.x, l); q.y = __shfl_sync(0xffffffffu, e.y, l);
              q.z = __shfl_sync(0xffffffffu, e.z, l); q.w = __shfl_sync(0xffffffffu, e.w, l);
              paint_window<PIX>(scr, STRIDE, wx0, wy0, wx1, wy1, q, R, lane);
              __syncwarp();
            }
          }
        }
        for (int y = 0; y < wy1 - wy0; y++)
          for (int c = lane; c < nwr; c += 32)
            reinterpret_cast<uint32_t *>(out + ((size_t)(wy0 + y) * W + wx0) * PIX)[c] = reinterpret_cast<const uint32_t *>(scr + y * STRIDE)[c];
        __syncwarp();
      }
      continue;
    }
    if (n == 0) continue;

    /* the marked tiles, one run of horizontally adjacent tiles of a tile row at a time */
    for (int ty = tyA; ty < tyB; ty++) {
      uint32_t rowbits = (tmask[ty >> 1] >> ((ty & 1) * 16)) & 0xffffu;
      while (rowbits) {
        const int t0 = __ffs(rowbits) - 1;
        const int len = min(__ffs(~(rowbits >> t0)) - 1, TBX_NT_MAX_RUN(PIX));
        rowbits &= ~(((1u << len) - 1u) << t0);
        const int rx0 = t0 * TBX_NT_TW, rx1 = min(W, rx0 + len * TBX_NT_TW), ry0 = ty * TBX_NT_TH, ry1 = min(H, ry0 + TBX_NT_TH);
        uint32_t hm[TBX_NT_LCAP / 32];
        int bx0 = 65535, bx1 = 0, by0 = 65535, by1 = 0;
#pragma unroll
        for (int c = 0; c < TBX_NT_LCAP / 32; c++) {
          hm[c] = 0;
          if (c * 32 >= n) continue;
          bool hit = false;
          if (c * 32 + lane < n) {
            const uint2 ext = exts[c * 32 + lane];
            const int x0 = ext.x & 0xffffu, x1 = ext.x >> 16, y0 = ext.y & 1023u, y1 = (ext.y >> 10) & 1023u;
            hit = x0 < rx1 && x1 > rx0 && y0 < ry1 && y1 > ry0;
            if (hit) { bx0 = min(bx0, x0); bx1 = max(bx1, x1); by0 = min(by0, y0); by1 = max(by1, y1); }
          }
          hm[c] = __ballot_sync(0xffffffffu, hit);
        }
        /* the window: the hit entries' bounding box inside the run, columns widened to multiples of 4 pixels so that
         * its rows are whole 32-bit words in every pixel format */
        const int wx0 = max(rx0, __reduce_min_sync(0xffffffffu, bx0)) & ~3, wx1 = min(rx1, (__reduce_max_sync(0xffffffffu, bx1) + 3) & ~3);
        const int wy0 = max(ry0, __reduce_min_sync(0xffffffffu, by0)), wy1 = min(ry1, __reduce_max_sync(0xffffffffu, by1));
        if (wx0 >= wx1 || wy0 >= wy1) continue;
        const int nwr = ((wx1 - wx0) * PIX) >> 2, nrows = wy1 - wy0;
        {
          const int lg = nwr > 16 ? 5 : nwr > 8 ? 4 : nwr > 4 ? 3 : 2;
          const int rstep = 32 >> lg;
          for (int cc = lane & ((1 << lg) - 1); cc < nwr; cc += 32)
            for (int y = lane >> lg; y < nrows; y += rstep)
              reinterpret_cast<uint32_t *>(scr + y * STRIDE)[cc] = __ldg(reinterpret_cast<const uint32_t *>(bfr + ((size_t)(wy0 + y) * W + wx0) * PIX) + cc);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < TBX_NT_LCAP / 32; c++) {
          uint32_t m = hm[c];
          if (m == 0) continue;
          uint4 e = make_uint4(0, 0, 0, 0);
          uint32_t gf = 0; /* group << 20 | flags << 28 */
          if ((m >> lane) & 1u) { e = list[c * 32 + lane]; gf = exts[c * 32 + lane].y >> 20; }
          while (m) {
            const int l = __ffs(m) - 1;
            const uint32_t gl = __shfl_sync(0xffffffffu, gf, l);
            if (!((gl >> 8) & TBX_ENTRY_PAR)) { /* in-order group: one entry at a time */
              const uint4 q = list[c * 32 + l];
              paint_window<PIX>(scr, STRIDE, wx0, wy0, wx1, wy1, q, R, lane);
              __syncwarp();
              m &= m - 1;
              continue;
            }
            const bool mine = ((m >> lane) & 1u) && (gf & 255u) == (gl & 255u);
            const uint32_t same = __ballot_sync(0xffffffffu, mine);
            bool small = mine && ((gf >> 8) & TBX_ENTRY_SMALL);
            uint32_t smalls = __ballot_sync(0xffffffffu, small);
            if (__popc(smalls) < 3) { smalls = 0; small = false; }
            uint32_t big = same & ~smalls;
            if (small) paint_window_lane<PIX>(scr, STRIDE, wx0, wy0, wx1, wy1, e);
            __syncwarp();
            while (big) {
              const int lb = __ffs(big) - 1;
              big &= big - 1;
              const uint4 q = list[c * 32 + lb];
              paint_window<PIX>(scr, STRIDE, wx0, wy0, wx1, wy1, q, R, lane);
              __syncwarp();
            }
            m &= ~same;
          }
        }
        { /* overwrite the window in the frame */
          const int lg = nwr > 16 ? 5 : nwr > 8 ? 4 : nwr > 4 ? 3 : 2;
          const int rstep = 32 >> lg;
          for (int cc = lane & ((1 << lg) - 1); cc < nwr; cc += 32)
            for (int y = lane >> lg; y < nrows; y += rstep)
              reinterpret_cast<uint32_t *>(out + ((size_t)(wy0 + y) * W + wx0) * PIX)[cc] = reinterpret_cast<const uint32_t *>(scr + y * STRIDE)[cc];
        }
        __syncwarp();
      }
    }
  }
}

template <int GAME, int PIX>
__global__ void __launch_bounds__(256, 4) native_patch_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ typename Traits<GAME>::Cfg cfg_c) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H, RW = T::RW;
  constexpr int NTY = (H + TBX_NT_TH - 1) / TBX_NT_TH; /* tile rows (<= 32) */
  static_assert(NTY <= 32 && (W + TBX_NT_TW - 1) / TBX_NT_TW <= 16, "tile mask layout");
  extern __shared__ uint4 smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(smem_raw);
  uint32_t *recs = reinterpret_cast<uint32_t *>(smem);
  const typename T::Cfg &cfg = cfg_c;
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int e0 = blockIdx.x * TBX_EPC;
  const int ne = min(TBX_EPC, a.n - e0);
  for (int i = tid; i < RW * TBX_EPC; i += blockDim.x) {
    const int w = i / TBX_EPC, j = i - w * TBX_EPC;
    if (j < ne) recs[j * RW + w] = a.planes[(size_t)w * a.n_pad + e0 + j];
  }
  __syncthreads(); /* the only CTA barrier */

  uint8_t *wmem = smem + a.smem_canvas + wid * a.warp_bytes;
  uint4 *list = reinterpret_cast<uint4 *>(wmem);
  uint2 *exts = reinterpret_cast<uint2 *>(wmem + TBX_NT_LCAP * 16);
  uint32_t *tmask = reinterpret_cast<uint32_t *>(wmem + TBX_NT_LCAP * 24);
  uint8_t *scr = wmem + TBX_NT_LCAP * 24 + 64;

  for (int j = wid; j < ne; j += nwarps)
    patch_env<GAME, PIX>(a, recs + j * RW, cfg, tables, 0, NTY, a.dst + (size_t)(e0 + j) * a.frame_bytes, list, exts, tmask, scr, lane);
}

} /* namespace tbxk */
#endif
