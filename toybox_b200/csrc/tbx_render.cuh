/* tbx_render.cuh -- the render kernel (Toybox.get_state / get_rgb_frame for a whole pool; fused WarpFrame).
 *
 * One CTA renders a chunk of TBX_EPC = 8 consecutive envs (8 envs x 4 B = one 32-byte sector per state word, so
 * the word-major planes are read at full sector efficiency).  Per env:
 *   1. the STATIC part of the frame (config-only draw-list slots: walls, ground line) is not painted at all:
 *      a pre-rendered base canvas (and, for the INTER_AREA layout, its pre-computed down-sample) is copied in
 *      from L2;
 *   2. the dynamic draw list is built straight from the env's record and painted into the shared-memory canvas
 *      group by group, in draw order.  Inside a group whose primitives cannot conflict (disjoint or same colour)
 *      every thread paints its own small rectangle with 32-bit span stores and larger / masked primitives are
 *      painted warp-cooperatively; groups that may conflict are painted by one warp, strictly in order.  The
 *      result equals the painter's algorithm over the whole draw list;
 *   3. native layouts (RGBA / RGB / gray) stream the canvas band to HBM with 16-byte coalesced stores;
 *      the INTER_AREA layout recomputes only the output pixels whose taps touch a dynamic primitive (dirty
 *      rectangles tracked while painting), with the exact f32 tap order of cv2's area resize, patches them
 *      into the staged base output and streams the 84x84 frame out with 16-byte stores.
 */
#ifndef TBX_RENDER_CUH
#define TBX_RENDER_CUH
#include <cuda_runtime.h>
#include "tbx_breakout.h"
#include "tbx_space_invaders.h"
#include "tbx_amidar.h"

namespace tbxk {

#ifndef TBX_RENDER_THREADS
#define TBX_RENDER_THREADS 256
#endif
#ifndef TBX_RENDER_MIN_CTAS
#define TBX_RENDER_MIN_CTAS 4
#endif
#define TBX_RENDER_WARPS (TBX_RENDER_THREADS / 32)
#define TBX_EPC 8
#define TBX_MAX_GROUPS 4
#define TBX_MAX_RECTS (TBX_MAX_GROUPS + 32)

__device__ const uint32_t d_bank[TBX_BANK_WORDS] = TBX_BANK_INIT;

template <int GAME> struct Traits;
template <> struct Traits<TBX_BREAKOUT> {
  typedef BrkCfg Cfg; typedef BrkTable Table; typedef BrkRec Rec;
  static constexpr int W = TBX_BRK_W, H = TBX_BRK_H, NS = BRK_N_SLOTS, RW = TBX_WORDS(BrkRec), NG = BRK_N_GROUPS;
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *t, int in) { brk_step(S, c, t, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *) { brk_new_game(S, c); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &c, const Table *t, int s) { return brk_prim(R, c, t, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &c) { return c.bg_color; }
  static __device__ __forceinline__ void group(int g, const uint32_t *R, const Table *t, int &b, int &e, int &mode) { brk_group(g, R, t, b, e, mode); }
};
template <> struct Traits<TBX_SPACE_INVADERS> {
  typedef SiCfg Cfg; typedef int Table; typedef SiRec Rec;
  static constexpr int W = TBX_SI_W, H = TBX_SI_H, NS = SI_N_SLOTS, RW = TBX_WORDS(SiRec), NG = SI_N_GROUPS;
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *, int in) { si_step(S, c, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *) { si_new_game(S, c); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &, const Table *, int s) { return si_prim(R, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &) { return SI_COLOR_BLACK; }
  static __device__ __forceinline__ void group(int g, const uint32_t *, const Table *, int &b, int &e, int &mode) { si_group(g, b, e, mode); }
};
template <> struct Traits<TBX_AMIDAR> {
  typedef AmiCfg Cfg; typedef AmiTable Table; typedef AmiRec Rec;
  static constexpr int W = TBX_AMI_W, H = TBX_AMI_H, NS = AMI_N_SLOTS, RW = TBX_WORDS(AmiRec), NG = AMI_N_GROUPS;
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *t, int in) { ami_step(S, c, t, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *t) { ami_new_game(S, c, t); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &c, const Table *t, int s) { return ami_prim(R, c, t, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &c) { return c.bg_color; }
  static __device__ __forceinline__ void group(int g, const uint32_t *, const Table *, int &b, int &e, int &mode) { ami_group(g, b, e, mode); }
};

struct RenderArgs {
  const uint32_t *planes;
  int n, n_pad;
  const void *cfg, *tables;
  uint8_t *dst;
  size_t frame_bytes;
  const uint8_t *base;     /* static frame: gray bytes (gray layouts) or RGBA (colour layouts) */
  const uint8_t *base_out; /* INTER_AREA: down-sample of the static frame */
  const TbxAreaPlan *plan; /* INTER_AREA */
  int band_rows;           /* native layouts: canvas rows per pass */
  int smem_canvas, smem_out, smem_plan, smem_rects; /* byte offsets into dynamic shared memory */
};

template <int PIX> struct PixT;
template <> struct PixT<1> { typedef uint8_t T; };
template <> struct PixT<4> { typedef uint32_t T; };

struct Clip { int x0, y0, x1, y1; };
template <int W> __device__ __forceinline__ bool clip_prim(const TbxPrim &p, int r0, int r1, Clip &c) {
  c.x0 = max((int)p.x, 0); c.x1 = min((int)p.x + (int)p.w, W);
  c.y0 = max((int)p.y, r0); c.y1 = min((int)p.y + (int)p.h, r1);
  return p.h > 0 && c.x0 < c.x1 && c.y0 < c.y1;
}

/* one thread fills a small solid rectangle: aligned 32-bit span stores for the 1-byte canvas */
template <int PIX, int W>
__device__ __forceinline__ void paint_small(typename PixT<PIX>::T *canvas, int r0, const Clip &c, uint32_t val) {
  for (int y = c.y0; y < c.y1; y++) {
    if (PIX == 1) {
      uint8_t *row = reinterpret_cast<uint8_t *>(canvas) + (size_t)(y - r0) * W;
      const uint32_t v4 = (val & 255u) * 0x01010101u;
      int x = c.x0;
      for (; x < c.x1 && (x & 3); x++) row[x] = (uint8_t)val;
      for (; x + 4 <= c.x1; x += 4) *reinterpret_cast<uint32_t *>(row + x) = v4;
      for (; x < c.x1; x++) row[x] = (uint8_t)val;
    } else {
      uint32_t *row = reinterpret_cast<uint32_t *>(canvas) + (size_t)(y - r0) * W;
      for (int x = c.x0; x < c.x1; x++) row[x] = val;
    }
  }
}

/* a whole warp paints one primitive (solid or sprite-masked), lanes over its clipped pixels */
template <int PIX, int W>
__device__ __forceinline__ void paint_coop(typename PixT<PIX>::T *canvas, int r0, const Clip &c, int qx, int qy, uint32_t val, uint32_t q3,
                                           const uint32_t *rec, int lane) {
  typedef typename PixT<PIX>::T P;
  const int nw = c.x1 - c.x0, cnt = nw * (c.y1 - c.y0);
  const float inv_nw = 1.0f / (float)nw;
  const int bw = (q3 >> 16) & 255;
  if (bw == 0) {
    for (int i = lane; i < cnt; i += 32) {
      int yy = (int)(((float)i + 0.5f) * inv_nw), xx = i - yy * nw;
      canvas[(size_t)(c.y0 + yy - r0) * W + c.x0 + xx] = (P)val;
    }
  } else {
    const uint32_t off = q3 & 0xffffu;
    const uint32_t *rows = (off & TBX_PRIM_STATE) ? rec + (off & 0x7fffu) : d_bank + off;
    const int sx = (q3 >> 24) & 15, sy = q3 >> 28;
    const float inv_sx = 1.0f / (float)sx, inv_sy = 1.0f / (float)sy;
    for (int i = lane; i < cnt; i += 32) {
      int yy = (int)(((float)i + 0.5f) * inv_nw), xx = i - yy * nw;
      int sy_i = (int)(((float)(c.y0 + yy - qy) + 0.5f) * inv_sy), sx_i = (int)(((float)(c.x0 + xx - qx) + 0.5f) * inv_sx);
      if ((rows[sy_i] >> (bw - 1 - sx_i)) & 1u) canvas[(size_t)(c.y0 + yy - r0) * W + c.x0 + xx] = (P)val;
    }
  }
}

/* warp-wide bounding box of the lanes' clipped rectangles (hardware integer warp reductions) */
__device__ __forceinline__ int4 warp_bbox(int bx0, int by0, int bx1, int by1) {
  return make_int4(__reduce_min_sync(0xffffffffu, bx0), __reduce_min_sync(0xffffffffu, by0), __reduce_max_sync(0xffffffffu, bx1),
                   __reduce_max_sync(0xffffffffu, by1));
}

/* Paint every dynamic draw-list group of one env into canvas rows [r0,r1), in draw order.  When `rects` is given
 * the dirty rectangles (native coordinates) are appended to it: one bounding box per group, except that an in-order
 * group of at most 32 slots contributes one rectangle per primitive (a ball far from the paddle must not dirty
 * everything in between).  All threads of the CTA must call this; it ends with a barrier. */
template <int GAME, int PIX>
__device__ __forceinline__ void paint_env(const uint32_t *R, const typename Traits<GAME>::Cfg &cfg, const typename Traits<GAME>::Table *tables,
                                          typename PixT<PIX>::T *canvas, int r0, int r1, int4 *rects, int *n_rects) {
  typedef Traits<GAME> T;
  constexpr int W = T::W;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int g = 0; g < T::NG; g++) {
    int gb, ge, gmode;
    T::group(g, R, tables, gb, ge, gmode);
    if (gb >= ge) continue; /* empty group (merged into its neighbour): uniform across the CTA, no barrier */
    if (!(gmode & TBX_GROUP_SERIAL)) {
      int bx0 = 32767, by0 = 32767, bx1 = -1, by1 = -1;
      for (int s0 = gb; s0 < ge; s0 += TBX_RENDER_THREADS) {
        const int s = s0 + tid;
        TbxPrim p = tbx_prim_none();
        if (s < ge) p = T::prim(R, cfg, tables, s);
        Clip c;
        const bool ok = clip_prim<W>(p, r0, r1, c);
        if (!__any_sync(0xffffffffu, ok)) continue;
        const uint32_t val = PIX == 1 ? tbx_luma(p.color) : p.color;
        if (ok) { bx0 = min(bx0, c.x0); by0 = min(by0, c.y0); bx1 = max(bx1, c.x1); by1 = max(by1, c.y1); }
        const bool small = ok && p.bw == 0 && (c.x1 - c.x0) * (c.y1 - c.y0) <= 96;
        if (small) paint_small<PIX, W>(canvas, r0, c, val);
        unsigned big = __ballot_sync(0xffffffffu, ok && !small);
        if (big) {
          const uint32_t w0 = (uint16_t)p.x | ((uint32_t)(uint16_t)p.y << 16), w1 = (uint16_t)p.w | ((uint32_t)(uint16_t)p.h << 16);
          const uint32_t w3 = (uint32_t)p.off | ((uint32_t)p.bw << 16) | ((uint32_t)p.scale << 24);
          while (big) {
            const int l = __ffs(big) - 1;
            big &= big - 1;
            TbxPrim q;
            const uint32_t q0 = __shfl_sync(0xffffffffu, w0, l), q1 = __shfl_sync(0xffffffffu, w1, l);
            const uint32_t qv = __shfl_sync(0xffffffffu, val, l), q3 = __shfl_sync(0xffffffffu, w3, l);
            q.x = (int16_t)(q0 & 0xffffu); q.y = (int16_t)(q0 >> 16); q.w = (int16_t)(q1 & 0xffffu); q.h = (int16_t)(q1 >> 16);
            Clip qc;
            clip_prim<W>(q, r0, r1, qc);
            paint_coop<PIX, W>(canvas, r0, qc, q.x, q.y, qv, q3, R, lane);
          }
        }
      }
      if (rects && __any_sync(0xffffffffu, bx1 > bx0)) {
        const int4 bb = warp_bbox(bx0, by0, bx1, by1);
        if (lane == 0) { atomicMin(&rects[g].x, bb.x); atomicMin(&rects[g].y, bb.y); atomicMax(&rects[g].z, bb.z); atomicMax(&rects[g].w, bb.w); }
      }
    } else if (wid == 0) {
      /* in order: lane l builds primitive gb + 32*batch + l, then the warp paints them one at a time */
      int bx0 = 32767, by0 = 32767, bx1 = -1, by1 = -1;
      const bool per_prim_rects = (ge - gb) <= 32;
      for (int s0 = gb; s0 < ge; s0 += 32) {
        const int s = s0 + lane;
        TbxPrim p = tbx_prim_none();
        if (s < ge) p = T::prim(R, cfg, tables, s);
        Clip c;
        const bool ok = clip_prim<W>(p, r0, r1, c);
        const uint32_t val = PIX == 1 ? tbx_luma(p.color) : p.color;
        if (ok) { bx0 = min(bx0, c.x0); by0 = min(by0, c.y0); bx1 = max(bx1, c.x1); by1 = max(by1, c.y1); }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (rects && per_prim_rects && ok) rects[T::NG + __popc(m & ((1u << lane) - 1u))] = make_int4(c.x0, c.y0, c.x1, c.y1);
        if (rects && per_prim_rects && lane == 0) *n_rects = T::NG + __popc(m);
        const uint32_t w0 = (uint16_t)p.x | ((uint32_t)(uint16_t)p.y << 16), w1 = (uint16_t)p.w | ((uint32_t)(uint16_t)p.h << 16);
        const uint32_t w3 = (uint32_t)p.off | ((uint32_t)p.bw << 16) | ((uint32_t)p.scale << 24);
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          TbxPrim q;
          const uint32_t q0 = __shfl_sync(0xffffffffu, w0, l), q1 = __shfl_sync(0xffffffffu, w1, l);
          const uint32_t qv = __shfl_sync(0xffffffffu, val, l), q3 = __shfl_sync(0xffffffffu, w3, l);
          q.x = (int16_t)(q0 & 0xffffu); q.y = (int16_t)(q0 >> 16); q.w = (int16_t)(q1 & 0xffffu); q.h = (int16_t)(q1 >> 16);
          Clip qc;
          clip_prim<W>(q, r0, r1, qc);
          paint_coop<PIX, W>(canvas, r0, qc, q.x, q.y, qv, q3, R, lane);
          __syncwarp();
        }
      }
      if (rects && !per_prim_rects) {
        const int4 bb = warp_bbox(bx0, by0, bx1, by1);
        if (lane == 0 && bb.z > bb.x) rects[g] = bb;
      }
    }
    if (!(gmode & TBX_GROUP_NOSYNC)) __syncthreads();
  }
}

/* MODE: TBX_OBS_RGBA (0), TBX_OBS_RGB (1), TBX_OBS_GRAY (2), TBX_OBS_GRAY_AREA (3).
 * TX, TY: taps per output column / row of the INTER_AREA plan (>= plan.tx, plan.ty; surplus taps have zero weight). */
template <int GAME, int MODE, int TX, int TY>
__global__ void __launch_bounds__(TBX_RENDER_THREADS, TBX_RENDER_MIN_CTAS) render_kernel(RenderArgs a) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H, RW = T::RW;
  constexpr int PIX = (MODE == 0 || MODE == 1) ? 4 : 1;
  typedef typename PixT<PIX>::T P;
  extern __shared__ uint4 smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(smem_raw);
  uint32_t *recs = reinterpret_cast<uint32_t *>(smem);
  P *canvas = reinterpret_cast<P *>(smem + a.smem_canvas);
  const typename T::Cfg &cfg = *(const typename T::Cfg *)a.cfg;
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int e0 = blockIdx.x * TBX_EPC;
  const int ne = min(TBX_EPC, a.n - e0);

  /* coalesced load of the chunk's state words: thread -> (word, env) with env fastest */
  for (int i = tid; i < RW * TBX_EPC; i += TBX_RENDER_THREADS) {
    int w = i / TBX_EPC, j = i - w * TBX_EPC;
    if (j < ne) recs[j * RW + w] = a.planes[(size_t)w * a.n_pad + e0 + j];
  }

  if (MODE != 3) {
    __syncthreads();
    for (int j = 0; j < ne; j++) {
      const uint32_t *R = recs + j * RW;
      uint8_t *out = a.dst + (size_t)(e0 + j) * a.frame_bytes;
      for (int r0 = 0; r0 < H; r0 += a.band_rows) {
        const int r1 = min(H, r0 + a.band_rows);
        const int n16 = (r1 - r0) * W * PIX / 16;
        {
          const uint4 *src = reinterpret_cast<const uint4 *>(a.base + (size_t)r0 * W * PIX);
          uint4 *dst = reinterpret_cast<uint4 *>(canvas);
          for (int i = tid; i < n16; i += TBX_RENDER_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        paint_env<GAME, PIX>(R, cfg, tables, canvas, r0, r1, (int4 *)0, (int *)0);
        if (MODE == 1) {
          /* 16 pixels (64 B of RGBA) -> 48 B of RGB, three 16-byte stores per thread */
          const int groups = (r1 - r0) * W / 16;
          const uint4 *src = reinterpret_cast<const uint4 *>(canvas);
          uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)r0 * W * 3);
          for (int g = tid; g < groups; g += TBX_RENDER_THREADS) {
            uint32_t p[16];
#pragma unroll
            for (int k = 0; k < 4; k++) { uint4 v = src[g * 4 + k]; p[4 * k] = v.x; p[4 * k + 1] = v.y; p[4 * k + 2] = v.z; p[4 * k + 3] = v.w; }
            uint32_t o[12];
#pragma unroll
            for (int k = 0; k < 4; k++) { /* 4 pixels -> 3 words */
              uint32_t c0 = p[4 * k] & 0xffffffu, c1 = p[4 * k + 1] & 0xffffffu, c2 = p[4 * k + 2] & 0xffffffu, c3 = p[4 * k + 3] & 0xffffffu;
              o[3 * k] = c0 | (c1 << 24);
              o[3 * k + 1] = (c1 >> 8) | (c2 << 16);
              o[3 * k + 2] = (c2 >> 16) | (c3 << 8);
            }
            __stcs(dst + g * 3, make_uint4(o[0], o[1], o[2], o[3]));
            __stcs(dst + g * 3 + 1, make_uint4(o[4], o[5], o[6], o[7]));
            __stcs(dst + g * 3 + 2, make_uint4(o[8], o[9], o[10], o[11]));
          }
        } else {
          const uint4 *src = reinterpret_cast<const uint4 *>(canvas);
          uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)r0 * W * PIX);
          for (int i = tid; i < n16; i += TBX_RENDER_THREADS) __stcs(dst + i, src[i]);
        }
        __syncthreads();
      }
    }
    return;
  }

  /* ---- INTER_AREA layout.  The staged output is double-buffered so that streaming frame j out overlaps with
   * staging the base frame for env j+1 (one barrier fewer per env). */
  const TbxAreaPlan *__restrict__ plan = a.plan; /* read through L1: small, shared by every CTA */
  int4 *rects = reinterpret_cast<int4 *>(smem + a.smem_rects);
  int *n_rects = reinterpret_cast<int *>(rects + TBX_MAX_RECTS);
  const int dw = plan->dw, dh = plan->dh;
  const int nout16 = (dw * dh + 15) / 16;
  const int ostage_bytes = nout16 * 16;
  for (int j = 0; j <= ne; j++) {
    if (j > 0) { /* stream frame j-1 out */
      const uint8_t *ostage = smem + a.smem_out + ((j - 1) & 1) * ostage_bytes;
      uint8_t *out = a.dst + (size_t)(e0 + j - 1) * a.frame_bytes;
      if ((a.frame_bytes & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ostage);
        uint4 *dst = reinterpret_cast<uint4 *>(out);
        for (int i = tid; i < dw * dh / 16; i += TBX_RENDER_THREADS) __stcs(dst + i, src[i]);
      } else {
        for (int i = tid; i < dw * dh; i += TBX_RENDER_THREADS) out[i] = ostage[i];
      }
    }
    if (j == ne) break;
    const uint32_t *R = recs + j * RW;
    uint8_t *ostage = smem + a.smem_out + (j & 1) * ostage_bytes;
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(a.base);
      uint4 *dst = reinterpret_cast<uint4 *>(canvas);
#pragma unroll 4
      for (int i = tid; i < W * H / 16; i += TBX_RENDER_THREADS) dst[i] = __ldg(src + i);
      const uint4 *src2 = reinterpret_cast<const uint4 *>(a.base_out);
      uint4 *dst2 = reinterpret_cast<uint4 *>(ostage);
      for (int i = tid; i < nout16; i += TBX_RENDER_THREADS) dst2[i] = __ldg(src2 + i);
      if (tid < T::NG) rects[tid] = make_int4(32767, 32767, -1, -1);
      if (tid == 0) *n_rects = T::NG;
    }
    __syncthreads();
    paint_env<GAME, 1>(R, cfg, tables, reinterpret_cast<uint8_t *>(canvas), 0, H, rects, n_rects);
    /* recompute the output pixels fed by a dirty rectangle.  Lanes own output columns (their taps stay in
     * registers), warps own output rows; narrow rectangles pack several rows into one warp.  The tap loops are
     * straight-line: TX x TY taps, surplus taps carry zero weights (x + 0*b == x exactly for these sums). */
    const int nr = *n_rects;
    for (int r = 0; r < nr; r++) {
      const int4 rc = rects[r];
      if (rc.z <= rc.x) continue;
      const int dx0 = __ldg(&plan->xdlo[rc.x]), dx1 = __ldg(&plan->xdhi[rc.z - 1]);
      const int dy0 = __ldg(&plan->ydlo[rc.y]), dy1 = __ldg(&plan->ydhi[rc.w - 1]);
      const int ncols = dx1 - dx0 + 1;
      const int lg = ncols > 16 ? 5 : ncols > 8 ? 4 : ncols > 4 ? 3 : 2; /* columns per warp pass = 1 << lg */
      const int cpl = 1 << lg, rpi = 32 >> lg, sub = lane >> lg, c = lane & (cpl - 1);
      for (int dxb = dx0; dxb <= dx1; dxb += cpl) {
        const bool colok = dxb + c <= dx1;
        const int dx = colok ? dxb + c : dx1;
        const uint8_t *col = reinterpret_cast<const uint8_t *>(canvas) + __ldg(&plan->xs0[dx]);
        float al[TX];
#pragma unroll
        for (int t = 0; t < TX; t++) al[t] = __ldg(&plan->xalpha[t][dx]);
        for (int dy = dy0 + wid * rpi + sub; dy <= dy1; dy += TBX_RENDER_WARPS * rpi) {
          const uint8_t *row = col + (size_t)__ldg(&plan->ys0[dy]) * W;
          float v = 0.0f;
#pragma unroll
          for (int k = 0; k < TY; k++) {
            float h = tbx_fmul((float)row[k * W], al[0]);
#pragma unroll
            for (int t = 1; t < TX; t++) h = tbx_fadd(h, tbx_fmul((float)row[k * W + t], al[t]));
            const float bh = tbx_fmul(__ldg(&plan->yalpha[k][dy]), h);
            v = k == 0 ? bh : tbx_fadd(v, bh);
          }
          const int iv = tbx_f2i_rn(v);
          if (colok) ostage[dy * dw + dx] = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
      }
    }
    __syncthreads();
  }
}

} /* namespace tbxk */
#endif
