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
 *   2. native_patch_kernel (stream-ordered after 1.): one WARP per env builds the entries that differ from the base and
 *      paints them straight into the frame in draw order (painter's algorithm on top of the base that is already
 *      there).  Work and traffic are proportional to what differs from the base; every sprite is painted once.
 * (Measured and dropped: the patches of env chunk c on a second stream while chunk c+1 is broadcast -- no gain
 * with 2, 4 or 8 chunks, even with the broadcast CTAs padded to 104 KB of shared memory so that both grids fit an SM; and both steps fused per (8 envs, band) CTA -- 100+ registers, the patches wait for the stores to
 * LAND and every band rebuilds the entries: Amidar RGB 24.7 M frames/s against 35.3 M for the two launches.)
 */
#ifndef TBX_RENDER_NATIVE_CUH
#define TBX_RENDER_NATIVE_CUH
#include "tbx_render_area.cuh"

namespace tbxk {

#define TBX_FILL_ENVS 32   /* envs per CTA of the broadcast kernel */

/* ---- 0. which envs differ from their base frame in so many places that the canvas kernel is the cheaper renderer:
 * a warp per env evaluates the game's estimate straight from the state planes, flags the env and appends it to the list */
template <int GAME>
__global__ void __launch_bounds__(256) dense_classify_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ typename Traits<GAME>::Cfg cfg_c) {
  typedef Traits<GAME> T;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (env >= a.n) return;
  const uint32_t *col = a.planes + env;
  const size_t n_pad = (size_t)a.n_pad;
  const int hint = T::dense_hint([col, n_pad](int w) { return col[(size_t)w * n_pad]; }, cfg_c, (const typename T::Table *)a.tables);
  if (lane == 0) {
    const bool dense = hint > a.dense_threshold;
    a.dense_flag[env] = dense;
    if (dense) a.dense_list[atomicAdd(a.dense_count, 1)] = env;
  }
}

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
    base_of[tid] = (a.dense_flag && a.dense_flag[e0 + tid]) ? -1 : T::base_id(hdr, cfg_c, tables); /* -1: the canvas kernel's env */
  }
  __syncthreads();
  int first = -1;
  for (int j = 0; j < ne && first < 0; j++) first = base_of[j];
  if (first < 0) return; /* every env of the chunk goes to the canvas kernel */
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
    if (base_of[j] == first || base_of[j] < 0) continue;
    const uint4 *src = reinterpret_cast<const uint4 *>((base_of[j] ? a.base[1] : a.base[0]) + (size_t)r0 * W * PIX);
    uint4 *dst = reinterpret_cast<uint4 *>(a.dst + (size_t)(e0 + j) * a.frame_bytes + (size_t)r0 * W * PIX);
    for (int i = tid; i < (int)(nbytes >> 4); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* the band must outlive the engine's reads */
}

/* ---- 2. patches: the frame already holds the base, so every entry that differs from it is painted STRAIGHT INTO THE
 * FRAME in draw order -- no scratch canvas, no tiles, every sprite exactly once.  A warp owns an env; entries of a
 * conflict-free group are painted with no ordering between them (small solid ones one per lane), everything else
 * one entry at a time; __syncwarp() between dependent paints orders the warp's global stores. */
template <int PIX, int W, int H>
__device__ __forceinline__ void paint_frame(uint8_t *out, const TbxPrim &q, const uint32_t *rec, int lane) {
  const int qx = q.x, qy = q.y;
  const int x0 = max(qx, 0), x1 = min(qx + (int)q.w, W), y0 = max(qy, 0), y1 = min(qy + (int)q.h, H);
  const int nw = x1 - x0, nh = y1 - y0;
  if (nw <= 0 || nh <= 0) return;
  const int lg = nw > 16 ? 5 : nw > 8 ? 4 : nw > 4 ? 3 : nw > 2 ? 2 : nw > 1 ? 1 : 0;
  const int cpl = 1 << lg, rpp = 32 >> lg, sub = lane >> lg, cx = lane & (cpl - 1);
  const uint32_t val = PIX == 1 ? tbx_luma(q.color) : q.color;
  if (q.bw == 0) {
    for (int yy = sub; yy < nh; yy += rpp) {
      uint8_t *row = out + (size_t)(y0 + yy) * W * PIX;
      for (int xb = cx; xb < nw; xb += cpl) put_pixel<PIX>(row, (size_t)(x0 + xb), val);
    }
  } else {
    const uint32_t off = q.off;
    const bool state = (off & TBX_PRIM_STATE) != 0; /* sprite rows in the env's record (shared) or in the bank (global) */
    const int o = state ? (int)(off & 0x7fffu) : (int)off;
    const int bw = q.bw, sx = q.scale & 15, sy = q.scale >> 4;
    const uint32_t ix = d_inv16[sx], iy = d_inv16[sy];
    for (int yy = sub; yy < nh; yy += rpp) {
      const int py = y0 + yy - qy;
      const int sy_i = sy == 1 ? py : (int)(((uint32_t)py * iy) >> 16);
      const uint32_t bits = state ? rec[o + sy_i] : __ldg(&d_bank[o + sy_i]);
      uint8_t *row = out + (size_t)(y0 + yy) * W * PIX;
      for (int xb = cx; xb < nw; xb += cpl) {
        const int px = x0 + xb - qx;
        const int sx_i = sx == 1 ? px : (int)(((uint32_t)px * ix) >> 16);
        if ((bits >> (bw - 1 - sx_i)) & 1u) put_pixel<PIX>(row, (size_t)(x0 + xb), val);
      }
    }
  }
}
/* one lane fills its own small solid rectangle */
template <int PIX, int W, int H>
__device__ __forceinline__ void paint_frame_lane(uint8_t *out, const TbxPrim &q) {
  const int x0 = max((int)q.x, 0), x1 = min((int)q.x + (int)q.w, W), y0 = max((int)q.y, 0), y1 = min((int)q.y + (int)q.h, H);
  const uint32_t val = PIX == 1 ? tbx_luma(q.color) : q.color;
  for (int y = y0; y < y1; y++) {
    uint8_t *row = out + (size_t)y * W * PIX;
    for (int x = x0; x < x1; x++) put_pixel<PIX>(row, (size_t)x, val);
  }
}

template <int GAME, int PIX>
__device__ __forceinline__ void patch_env(const uint32_t *R, const typename Traits<GAME>::Cfg &cfg, const typename Traits<GAME>::Table *tables, uint8_t *out, int lane) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H;
  const int base = T::base_id(R, cfg, tables);
  for (int g = 0; g < T::NG; g++) {
    int gb, ge, gmode;
    T::group(g, R, tables, base, gb, ge, gmode);
    T::trim(g, R, cfg, base, gb, ge);
    const bool par = !(gmode & TBX_GROUP_SERIAL);
    for (int s0 = gb; s0 < ge; s0 += 32) {
      const int s = s0 + lane;
      TbxPrim p = tbx_prim_none();
      if (s < ge) p = T::prim(R, cfg, tables, s, base);
      Clip c;
      const bool ok = clip_prim<W>(p, 0, H, c);
      unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m == 0) continue;
      if (par) { /* no two members conflict: the small solid ones each by its own lane, when there are enough of them */
        bool small = ok && p.bw == 0 && (c.x1 - c.x0) * (c.y1 - c.y0) <= 64;
        unsigned smalls = __ballot_sync(0xffffffffu, small);
        if (__popc(smalls) < 3) { smalls = 0; small = false; }
        if (small) paint_frame_lane<PIX, W, H>(out, p);
        m &= ~smalls;
      }
      const uint32_t w0 = (uint16_t)p.x | ((uint32_t)(uint16_t)p.y << 16), w1 = (uint16_t)p.w | ((uint32_t)(uint16_t)p.h << 16);
      const uint32_t w3 = (uint32_t)p.off | ((uint32_t)p.bw << 16) | ((uint32_t)p.scale << 24);
      while (m) { /* the warp paints one entry at a time, in slot order */
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t q0 = __shfl_sync(0xffffffffu, w0, l), q1 = __shfl_sync(0xffffffffu, w1, l);
        const uint32_t qc = __shfl_sync(0xffffffffu, p.color, l), q3 = __shfl_sync(0xffffffffu, w3, l);
        TbxPrim q;
        q.x = (int16_t)(q0 & 0xffffu); q.y = (int16_t)(q0 >> 16); q.w = (int16_t)(q1 & 0xffffu); q.h = (int16_t)(q1 >> 16);
        q.color = qc; q.off = (uint16_t)(q3 & 0xffffu); q.bw = (uint8_t)((q3 >> 16) & 255u); q.scale = (uint8_t)(q3 >> 24);
        paint_frame<PIX, W, H>(out, q, R, lane);
        if (!par) __syncwarp(); /* in-order group: the next entry may overwrite this one */
      }
      __syncwarp(); /* later passes and groups paint over this one */
    }
  }
}

template <int GAME, int PIX>
__global__ void __launch_bounds__(256, 4) native_patch_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ typename Traits<GAME>::Cfg cfg_c) {
  typedef Traits<GAME> T;
  constexpr int RW = T::RW;
  extern __shared__ uint4 smem_raw[];
  uint32_t *recs = reinterpret_cast<uint32_t *>(smem_raw);
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
  for (int j = wid; j < ne; j += nwarps) {
    const uint32_t *R = recs + j * RW;
    if (a.dense_flag && a.dense_flag[e0 + j]) continue; /* differs from its base in many places: the canvas kernel's env */
    patch_env<GAME, PIX>(R, cfg, tables, a.dst + (size_t)(e0 + j) * a.frame_bytes, lane);
  }
}

} /* namespace tbxk */
#endif
