/* tbx_kernels.cuh -- the sm_100a kernels: thread-per-env step / new_game / bookkeeping kernels over the
 * word-major state planes, and the render kernel (one CTA per chunk of 8 envs: coalesced state load,
 * draw-list build, shared-memory canvas painted by warp-owned row bands, 16-byte coalesced stores of
 * native RGBA / RGB / gray frames or the fused INTER_AREA down-sample).
 */
#ifndef TBX_KERNELS_CUH
#define TBX_KERNELS_CUH
#include <cuda_runtime.h>
#include "tbx_breakout.h"
#include "tbx_space_invaders.h"
#include "tbx_amidar.h"
#include "tbx_host.h"

namespace tbxk {

__device__ const uint32_t d_bank[TBX_BANK_WORDS] = TBX_BANK_INIT;

template <int GAME> struct Traits;
template <> struct Traits<TBX_BREAKOUT> {
  typedef BrkCfg Cfg; typedef BrkTable Table; typedef BrkRec Rec;
  static constexpr int W = TBX_BRK_W, H = TBX_BRK_H, NS = BRK_N_SLOTS, RW = TBX_WORDS(BrkRec);
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *t, int in) { brk_step(S, c, t, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *) { brk_new_game(S, c); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &c, const Table *t, int s) { return brk_prim(R, c, t, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &c) { return c.bg_color; }
};
template <> struct Traits<TBX_SPACE_INVADERS> {
  typedef SiCfg Cfg; typedef int Table; typedef SiRec Rec;
  static constexpr int W = TBX_SI_W, H = TBX_SI_H, NS = SI_N_SLOTS, RW = TBX_WORDS(SiRec);
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *, int in) { si_step(S, c, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *) { si_new_game(S, c); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &, const Table *, int s) { return si_prim(R, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &) { return SI_COLOR_BLACK; }
};
template <> struct Traits<TBX_AMIDAR> {
  typedef AmiCfg Cfg; typedef AmiTable Table; typedef AmiRec Rec;
  static constexpr int W = TBX_AMI_W, H = TBX_AMI_H, NS = AMI_N_SLOTS, RW = TBX_WORDS(AmiRec);
  static __device__ __forceinline__ void step(const TbxAcc &S, const Cfg &c, const Table *t, int in) { ami_step(S, c, t, in); }
  static __device__ __forceinline__ void new_game(const TbxAcc &S, const Cfg &c, const Table *t) { ami_new_game(S, c, t); }
  static __device__ __forceinline__ TbxPrim prim(const uint32_t *R, const Cfg &c, const Table *t, int s) { return ami_prim(R, c, t, s); }
  static __device__ __forceinline__ uint32_t clear_color(const Cfg &c) { return c.bg_color; }
};

/* ------------------------------------------------------------------ step */
struct StepArgs {
  uint32_t *planes;
  int n, n_pad;
  const void *cfg, *tables;
  const int32_t *actions; /* ALE ids, or NULL */
  const uint8_t *inputs;  /* Input bitmasks, or NULL */
  int auto_reset;
  int32_t *reward, *score, *lives;
  uint8_t *done;
  unsigned long long *stats; /* episodes, sum return, sum length, max return */
  int *bad_actions;
};

template <int GAME>
__global__ void __launch_bounds__(128) step_kernel(StepArgs a) {
  typedef Traits<GAME> T;
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.n) return;
  TbxAcc S;
  S.p = a.planes + env;
  S.stride = (size_t)a.n_pad;
  const typename T::Cfg &cfg = *(const typename T::Cfg *)a.cfg;
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  int in = a.actions ? tbx_ale_action_to_input(a.actions[env]) : (int)a.inputs[env];
  int lives_before = S.ldi(TBX_HW(lives));
  if (in < 0) atomicAdd(a.bad_actions, 1);
  else T::step(S, cfg, tables, in);
  TbxStepOut o = tbx_bookkeep(S, lives_before);
  if (a.reward) a.reward[env] = o.reward;
  if (a.done) a.done[env] = (uint8_t)o.done;
  if (a.score) a.score[env] = o.score;
  if (a.lives) a.lives[env] = o.lives;
  if (o.episode_ended) {
    atomicAdd(a.stats + 0, 1ull);
    atomicAdd(a.stats + 1, (unsigned long long)(long long)o.ep_return);
    atomicAdd(a.stats + 2, (unsigned long long)o.ep_len);
    atomicMax((long long *)(a.stats + 3), (long long)o.ep_return);
  }
  if (o.done && a.auto_reset) T::new_game(S, cfg, tables);
}

template <int GAME>
__global__ void __launch_bounds__(128) new_game_kernel(uint32_t *planes, int n, int n_pad, const void *cfg_, const void *tables_, const uint8_t *mask) {
  typedef Traits<GAME> T;
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n || (mask && !mask[env])) return;
  TbxAcc S;
  S.p = planes + env;
  S.stride = (size_t)n_pad;
  T::new_game(S, *(const typename T::Cfg *)cfg_, (const typename T::Table *)tables_);
}

/* sim_rand <- the pool config's rand (pool creation), or <- seed mapping (tbx_seed) */
__global__ void set_sim_rand_kernel(uint32_t *planes, int n, int n_pad, uint64_t s0, uint64_t s1) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  TbxAcc S; S.p = planes + env; S.stride = (size_t)n_pad;
  S.st64(TBX_HW(sim_rand), s0); S.st64(TBX_HW(sim_rand) + 2, s1);
}
__global__ void seed_kernel(uint32_t *planes, int n_pad, const uint32_t *seeds, const int32_t *ids, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  int env = ids ? ids[i] : i;
  TbxAcc S; S.p = planes + env; S.stride = (size_t)n_pad;
  TbxRng g;
  tbx_rng_seed(g, seeds[i]);
  S.st64(TBX_HW(sim_rand), g.s0); S.st64(TBX_HW(sim_rand) + 2, g.s1);
}
__global__ void read_scalars_kernel(const uint32_t *planes, int n, int n_pad, int32_t *score, int32_t *lives, int32_t *level) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  if (score) score[env] = (int32_t)planes[(size_t)TBX_HW(score) * n_pad + env];
  if (lives) lives[env] = (int32_t)planes[(size_t)TBX_HW(lives) * n_pad + env];
  if (level) level[env] = (int32_t)planes[(size_t)TBX_HW(level) * n_pad + env];
}
__global__ void fill_actions_kernel(int32_t *actions, int n, uint64_t seed, uint64_t env0, uint64_t t, const int32_t *legal, int n_legal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  actions[i] = legal[tbx_action_index(seed, env0 + (uint64_t)i, t, (uint32_t)n_legal)];
}
/* records (AoS, rw words each) <-> planes, for the JSON import/export of a few envs */
__global__ void gather_kernel(const uint32_t *planes, int n_pad, const int32_t *ids, int k, int rw, uint32_t *recs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * rw) return;
  int e = i / rw, w = i - e * rw;
  recs[i] = planes[(size_t)w * n_pad + ids[e]];
}
__global__ void scatter_kernel(uint32_t *planes, int n_pad, const int32_t *ids, int k, int rw, const uint32_t *recs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * rw) return;
  int e = i / rw, w = i - e * rw;
  planes[(size_t)w * n_pad + ids[e]] = recs[i];
}

/* ------------------------------------------------------------------ render */
#define TBX_RENDER_THREADS 256
#define TBX_RENDER_WARPS (TBX_RENDER_THREADS / 32)
#define TBX_EPC 8 /* envs per CTA: 8 envs x 4 B = one 32-byte sector per state word */

struct RenderArgs {
  const uint32_t *planes;
  int n, n_pad;
  const void *cfg, *tables;
  uint8_t *dst;
  size_t frame_bytes;
  const TbxResizeTab *rs; /* TBX_OBS_GRAY_AREA only */
  int out_w, out_h;
  int band_rows;     /* output rows per band */
  int canvas_rows;   /* canvas capacity in source rows */
  int smem_prims, smem_canvas, smem_buf, smem_out; /* byte offsets into dynamic shared memory */
};

template <int PIX> struct PixT;
template <> struct PixT<1> { typedef uint8_t T; };
template <> struct PixT<4> { typedef uint32_t T; };

/* Paint source rows [r0,r1) of the frame into `canvas` (row r0 at offset 0).  Each warp owns a contiguous
 * slice of the rows and paints every primitive that touches it, in draw-list order, so overlapping
 * primitives resolve exactly as the painter's algorithm does and no block-level barrier is needed. */
template <int PIX, int W>
__device__ __forceinline__ void paint_band(const TbxPrim *prims, int ns, const uint32_t *rec, typename PixT<PIX>::T *canvas,
                                           int r0, int r1, uint32_t clearv) {
  typedef typename PixT<PIX>::T P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rpw = (r1 - r0 + TBX_RENDER_WARPS - 1) / TBX_RENDER_WARPS;
  const int wy0 = r0 + wid * rpw, wy1 = min(r1, wy0 + rpw);
  if (wy0 >= wy1) return;
  {
    uint32_t c32 = PIX == 4 ? clearv : (clearv & 255u) * 0x01010101u;
    uint4 cv = make_uint4(c32, c32, c32, c32);
    uint4 *dst = reinterpret_cast<uint4 *>(canvas + (size_t)(wy0 - r0) * W);
    const int n16 = (wy1 - wy0) * W * PIX / 16;
    for (int i = lane; i < n16; i += 32) dst[i] = cv;
  }
  __syncwarp();
  for (int base = 0; base < ns; base += 32) {
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (base + lane < ns) raw = reinterpret_cast<const uint4 *>(prims)[base + lane];
    int px = (int16_t)(raw.x & 0xffffu), py = (int16_t)(raw.x >> 16), pw = (int16_t)(raw.y & 0xffffu), ph = (int16_t)(raw.y >> 16);
    bool hit = ph > 0 && py < wy1 && py + ph > wy0 && px < W && px + pw > 0;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t q0 = __shfl_sync(0xffffffffu, raw.x, l), q1 = __shfl_sync(0xffffffffu, raw.y, l);
      const uint32_t col = __shfl_sync(0xffffffffu, raw.z, l), q3 = __shfl_sync(0xffffffffu, raw.w, l);
      const int qx = (int16_t)(q0 & 0xffffu), qy = (int16_t)(q0 >> 16), qw = (int16_t)(q1 & 0xffffu), qh = (int16_t)(q1 >> 16);
      const int x0 = max(qx, 0), x1 = min(qx + qw, W), y0 = max(qy, wy0), y1 = min(qy + qh, wy1);
      const int nw = x1 - x0, cnt = nw * (y1 - y0);
      const float inv_nw = 1.0f / (float)nw;
      const int bw = (q3 >> 16) & 255;
      const P val = (P)col;
      if (bw == 0) {
        for (int i = lane; i < cnt; i += 32) {
          int yy = (int)(((float)i + 0.5f) * inv_nw), xx = i - yy * nw;
          canvas[(size_t)(y0 + yy - r0) * W + x0 + xx] = val;
        }
      } else {
        const uint32_t off = q3 & 0xffffu;
        const uint32_t *rows = (off & TBX_PRIM_STATE) ? rec + (off & 0x7fffu) : d_bank + off;
        const int sx = (q3 >> 24) & 15, sy = q3 >> 28;
        const float inv_sx = 1.0f / (float)sx, inv_sy = 1.0f / (float)sy;
        for (int i = lane; i < cnt; i += 32) {
          int yy = (int)(((float)i + 0.5f) * inv_nw), xx = i - yy * nw;
          int sy_i = (int)(((float)(y0 + yy - qy) + 0.5f) * inv_sy), sx_i = (int)(((float)(x0 + xx - qx) + 0.5f) * inv_sx);
          if ((rows[sy_i] >> (bw - 1 - sx_i)) & 1u) canvas[(size_t)(y0 + yy - r0) * W + x0 + xx] = val;
        }
      }
      __syncwarp();
    }
  }
}

/* MODE: TBX_OBS_RGBA, TBX_OBS_RGB, TBX_OBS_GRAY, TBX_OBS_GRAY_AREA */
template <int GAME, int MODE>
__global__ void __launch_bounds__(TBX_RENDER_THREADS) render_kernel(RenderArgs a) {
  typedef Traits<GAME> T;
  constexpr int W = T::W, H = T::H, NS = T::NS, RW = T::RW;
  constexpr int PIX = (MODE == TBX_OBS_RGBA || MODE == TBX_OBS_RGB) ? 4 : 1;
  typedef typename PixT<PIX>::T P;
  extern __shared__ uint4 smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(smem_raw);
  uint32_t *recs = reinterpret_cast<uint32_t *>(smem);
  TbxPrim *prims = reinterpret_cast<TbxPrim *>(smem + a.smem_prims);
  P *canvas = reinterpret_cast<P *>(smem + a.smem_canvas);
  const typename T::Cfg &cfg = *(const typename T::Cfg *)a.cfg;
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * TBX_EPC;
  const int ne = min(TBX_EPC, a.n - e0);

  /* coalesced load of the chunk's state words: thread -> (word, env) with env fastest */
  for (int i = tid; i < RW * TBX_EPC; i += TBX_RENDER_THREADS) {
    int w = i / TBX_EPC, j = i - w * TBX_EPC;
    if (j < ne) recs[j * RW + w] = a.planes[(size_t)w * a.n_pad + e0 + j];
  }
  __syncthreads();
  uint32_t clearv = T::clear_color(cfg);
  if (PIX == 1) clearv = tbx_luma(clearv);

  for (int j = 0; j < ne; j++) {
    const uint32_t *R = recs + j * RW;
    uint8_t *out = a.dst + (size_t)(e0 + j) * a.frame_bytes;
    for (int s = tid; s < NS; s += TBX_RENDER_THREADS) {
      TbxPrim p = T::prim(R, cfg, tables, s);
      if (PIX == 1) p.color = tbx_luma(p.color);
      prims[s] = p;
    }
    __syncthreads();
    if (MODE != TBX_OBS_GRAY_AREA) {
      for (int r0 = 0; r0 < H; r0 += a.band_rows) {
        const int r1 = min(H, r0 + a.band_rows);
        paint_band<PIX, W>(prims, NS, R, canvas, r0, r1, clearv);
        __syncthreads();
        if (MODE == TBX_OBS_RGB) {
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
            dst[g * 3] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[g * 3 + 1] = make_uint4(o[4], o[5], o[6], o[7]);
            dst[g * 3 + 2] = make_uint4(o[8], o[9], o[10], o[11]);
          }
        } else {
          const int n16 = (r1 - r0) * W * PIX / 16;
          const uint4 *src = reinterpret_cast<const uint4 *>(canvas);
          uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)r0 * W * PIX);
          for (int i = tid; i < n16; i += TBX_RENDER_THREADS) dst[i] = src[i];
        }
        __syncthreads();
      }
    } else {
      float *buf = reinterpret_cast<float *>(smem + a.smem_buf);
      uint8_t *ostage = smem + a.smem_out;
      const TbxResizeAxis &ax = a.rs->x, &ay = a.rs->y;
      const int dw = a.out_w, dh = a.out_h;
      for (int d0 = 0; d0 < dh; d0 += a.band_rows) {
        const int d1 = min(dh, d0 + a.band_rows);
        const int ys0 = ay.si[ay.start[d0]], ys1 = ay.si[ay.start[d1] - 1] + 1;
        paint_band<PIX, W>(prims, NS, R, canvas, ys0, ys1, clearv);
        __syncthreads();
        /* horizontal pass: buf[r][dx] = sum_k src[r][si_k] * alpha_k, taps accumulated in table order */
        const int nrows = ys1 - ys0;
        for (int i = tid; i < nrows * dw; i += TBX_RENDER_THREADS) {
          int r = i / dw, dx = i - r * dw;
          float acc = tbx_area_h(reinterpret_cast<const uint8_t *>(canvas) + (size_t)r * W, ax, dx);
          buf[i] = acc;
        }
        __syncthreads();
        /* vertical pass, round half to even, saturate */
        for (int i = tid; i < (d1 - d0) * dw; i += TBX_RENDER_THREADS) {
          int dy = d0 + i / dw, dx = i - (dy - d0) * dw;
          ostage[dy * dw + dx] = tbx_area_v(buf, dw, ys0, ay, dy, dx);
        }
        __syncthreads();
      }
      const int nbytes = dw * dh;
      if ((a.frame_bytes & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ostage);
        uint4 *dst = reinterpret_cast<uint4 *>(out);
        for (int i = tid; i < nbytes / 16; i += TBX_RENDER_THREADS) dst[i] = src[i];
      } else {
        for (int i = tid; i < nbytes; i += TBX_RENDER_THREADS) out[i] = ostage[i];
      }
      __syncthreads();
    }
  }
}

} /* namespace tbxk */
#endif
