/* tbx_render_direct.cuh -- the DIRECT INTER_AREA kernels (84x84 gray WarpFrame observation): one warp per env, no
 * canvas, no draw list, no tiles.  See tbx_direct.h for the closed forms; this file is their warp-level schedule.
 *
 * Replaces get_state() + cv2.resize(INTER_AREA) (toybox/envs/atari/base.py:109, baselines/baselines/common/
 * atari_wrappers.py:243) for the envs the closed forms cover; the others are appended to a list that the general tile
 * kernel (tbx_render_area.cuh, env-list mode) renders right after.
 *
 * Breakout, per env (warp):
 *   1. frame <- pre-computed down-sample of base frame 1 (every brick alive), 16-byte copies;
 *   2. if a brick is dead: the wall's H rows (one look-up per brick row and output column) go to shared memory and the
 *      output words (4 pixels) x rows that a dead brick feeds are recomputed from them -- cost bounded by the wall's
 *      output area, whatever the number of holes;
 *   3. HUD digits: pre-resolved patches;
 *   4. paddle and balls: the output pixels their rectangles feed, every tap evaluated analytically
 *      (brk_direct_pixel: base frame 0, brick grid, movers in draw order).
 */
#ifndef TBX_RENDER_DIRECT_CUH
#define TBX_RENDER_DIRECT_CUH
#include "tbx_render_area.cuh"
#include "tbx_direct.h"

namespace tbxk {

struct DirectArgs {
  const void *aux;    /* the game's closed-form tables on the device (TbxBrkDirect ...) */
  int32_t *fb_list;   /* envs handed to the tile kernel */
  int *fb_count;
  int hstride;        /* floats per H row in shared memory */
  int warp_bytes;     /* shared memory per warp */
};

#define TBX_DIRECT_THREADS 256

template <int TX, int TY>
__global__ void __launch_bounds__(TBX_DIRECT_THREADS, 4) brk_direct_kernel(const __grid_constant__ RenderArgs a, const __grid_constant__ BrkCfg cfg_c,
                                                                            const __grid_constant__ TbxAreaPlan plan_c, const __grid_constant__ DirectArgs d) {
  constexpr int RW = TBX_WORDS(BrkRec);
  extern __shared__ uint4 smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(smem_raw);
  uint32_t *recs = reinterpret_cast<uint32_t *>(smem);
  const TbxBrkDirect *__restrict__ Ap = reinterpret_cast<const TbxBrkDirect *>(d.aux);
  const TbxBrkDirect &A = *Ap;
  const TbxAreaPlan *__restrict__ plan = a.plan; /* per-lane indexed reads */
  const TbxAreaPlan &cp = plan_c;                /* warp-uniform reads: constant bank */
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int e0 = blockIdx.x * TBX_EPC;
  const int ne = min(TBX_EPC, a.n - e0);
  for (int i = tid; i < RW * TBX_EPC; i += blockDim.x) {
    const int w = i / TBX_EPC, j = i - w * TBX_EPC;
    if (j < ne) recs[j * RW + w] = a.planes[(size_t)w * a.n_pad + e0 + j];
  }
  __syncthreads(); /* the only CTA barrier */

  const int ok = A.ok, ncols = A.ncols, nrows = A.nrows, wdy0 = A.wdy0, wdy1 = A.wdy1, hud_dyhi = A.hud_dyhi;
  const int wy0 = A.wy0, wy1 = A.wy0 + A.nrows * A.bh;
  const int dw = cp.dw, dh = cp.dh, nwords = dw >> 2, hs = d.hstride;
  float *hw = reinterpret_cast<float *>(smem + a.smem_canvas + wid * d.warp_bytes);
  const TbxDigitPatch *__restrict__ patches = a.patches[1];

  for (int j = wid; j < ne; j += nwarps) {
    const uint32_t *R = recs + j * RW;
    const int env = e0 + j;
    /* movers: lane 0 the paddle, lanes 1..4 the balls; footprint = the output pixels the rectangle feeds */
    const TbxMover mine = brk_mover(R, cfg_c, A, lane < BRK_N_MOVERS ? lane : -1);
    const bool valid = mine.x0 < mine.x1;
    int fx0 = 0, fx1 = 0, fy0 = 0, fy1 = 0;
    if (valid) { fx0 = __ldg(&plan->xdlo[mine.x0]); fx1 = __ldg(&plan->xdhi[mine.x1 - 1]); fy0 = __ldg(&plan->ydlo[mine.y0]); fy1 = __ldg(&plan->ydhi[mine.y1 - 1]); }
    bool bad = valid && fy0 <= hud_dyhi;
    /* HUD digits: lanes 0..9 the score, 10..19 the lives */
    int dig = -1;
    const TbxDigitPatch *P = patches;
    if (lane < 2 * TBX_MAX_DIGITS && patches) {
      const int field = lane >= TBX_MAX_DIGITS;
      dig = tbx_digit_at((int32_t)R[field ? TBX_HW(lives) : TBX_HW(score)], lane - field * TBX_MAX_DIGITS);
      if (dig >= 0) { P = patches + lane * 10 + dig; bad |= __ldg(&P->w) == 0; }
    }
    const bool covered = ok && patches && (int32_t)R[TBX_HW(tbl)] == cfg_c.default_tbl;
    if (!covered || __any_sync(0xffffffffu, bad)) { /* the general kernel's */
      if (lane == 0) d.fb_list[atomicAdd(d.fb_count, 1)] = env;
      continue;
    }
    uint8_t *out = a.dst + (size_t)env * a.env_stride + (size_t)a.stack_slot * a.frame_bytes;
    { /* 1. every brick alive, nothing else */
      const uint4 *src = reinterpret_cast<const uint4 *>(a.base_out[1]);
      const int nb = dw * dh;
#pragma unroll 4
      for (int i = lane; i < (nb >> 4); i += 32) reinterpret_cast<uint4 *>(out)[i] = __ldg(src + i);
      for (int i = ((nb >> 4) << 2) + lane; i < (nb >> 2); i += 32) reinterpret_cast<uint32_t *>(out)[i] = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
    }
    /* 2. the wall */
    const uint32_t fullm = (1u << nrows) - 1u;
    const uint32_t colbits = lane < ncols ? brk_col_bits(R + BRK_W(alive), nrows, lane) : fullm;
    const uint32_t deadcols = __ballot_sync(0xffffffffu, colbits != fullm);
    __syncwarp(); /* orders the base copy before the patches other lanes write below */
    if (deadcols) {
      const uint32_t deadrows = __reduce_or_sync(0xffffffffu, ~colbits & fullm);
      uint32_t rowmask[TBX_BRK_MAX_ROWS];
#pragma unroll
      for (int r = 0; r < TBX_BRK_MAX_ROWS; r++) rowmask[r] = __ballot_sync(0xffffffffu, (colbits >> r) & 1u);
      const uint32_t wmask = __ballot_sync(0xffffffffu, lane < nwords && (__ldg(&A.wordcols[lane]) & deadcols));
      const uint32_t rmask = __ballot_sync(0xffffffffu, wdy0 + lane <= wdy1 && (__ldg(&A.dyrows[wdy0 + lane]) & deadrows));
      if (wmask && rmask) {
        const int wlo = __ffs(wmask) - 1, whi = 31 - __clz(wmask);
        for (int dx = 4 * wlo + lane; dx < 4 * whi + 4; dx += 32) {
          const int c0 = __ldg(&A.col0[dx]);
#pragma unroll
          for (int r = 0; r < TBX_BRK_MAX_ROWS; r++)
            if (r < nrows) hw[r * hs + dx] = __ldg(&A.hlut[r][(rowmask[r] >> c0) & 3u][dx]);
        }
        __syncwarp();
        const int naw = __popc(wmask), rlo = wdy0 + __ffs(rmask) - 1, rhi = wdy0 + 31 - __clz(rmask);
        const int lg = naw > 16 ? 5 : naw > 8 ? 4 : naw > 4 ? 3 : naw > 2 ? 2 : naw > 1 ? 1 : 0;
        const int jw = lane & ((1 << lg) - 1), sub = lane >> lg, rpi = 32 >> lg;
        const int myword = jw < naw ? (int)__fns(wmask, 0, jw + 1) : -1;
        for (int dyb = rlo; dyb <= rhi; dyb += rpi) {
          const int dy = dyb + sub;
          if (myword < 0 || dy > rhi) continue;
          float acc[4];
#pragma unroll
          for (int k = 0; k < TY; k++) {
            const int sel = __ldg(&A.hsel[dy][k]);
            const float4 h = sel < nrows ? *reinterpret_cast<const float4 *>(hw + sel * hs + 4 * myword)
                                         : __ldg(reinterpret_cast<const float4 *>(&A.hstatic[sel - nrows][4 * myword]));
            const float b = __ldg(&plan->yalpha[k][dy]);
            const float p0 = tbx_fmul(b, h.x), p1 = tbx_fmul(b, h.y), p2 = tbx_fmul(b, h.z), p3 = tbx_fmul(b, h.w);
            if (k == 0) { acc[0] = p0; acc[1] = p1; acc[2] = p2; acc[3] = p3; }
            else { acc[0] = tbx_fadd(acc[0], p0); acc[1] = tbx_fadd(acc[1], p1); acc[2] = tbx_fadd(acc[2], p2); acc[3] = tbx_fadd(acc[3], p3); }
          }
          uint32_t word = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int iv = tbx_f2i_rn_small(acc[q]);
            word |= (uint32_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv) << (8 * q);
          }
          *reinterpret_cast<uint32_t *>(out + dy * dw + 4 * myword) = word;
        }
      }
    }
    /* 3. HUD digits: one patch at a time, lanes over (row, column) */
    {
      uint32_t dm = __ballot_sync(0xffffffffu, dig >= 0);
      while (dm) {
        const int l = __ffs(dm) - 1;
        dm &= dm - 1;
        const TbxDigitPatch *Q = reinterpret_cast<const TbxDigitPatch *>(__shfl_sync(0xffffffffu, (unsigned long long)P, l));
        const int px0 = __ldg(&Q->x0), py0 = __ldg(&Q->y0), pw = __ldg(&Q->w), ph = __ldg(&Q->h);
        const int cc = lane & 7;
        if (cc < pw)
          for (int r = lane >> 3; r < ph; r += 4) out[(py0 + r) * dw + px0 + cc] = __ldg(&Q->px[r * pw + cc]);
      }
    }
    __syncwarp(); /* the movers' pixels go over the wall's */
    /* 4. paddle and balls */
    uint32_t mm = __ballot_sync(0xffffffffu, valid) & ((1u << BRK_N_MOVERS) - 1u);
    if (mm) {
      TbxMover mv[BRK_N_MOVERS];
#pragma unroll
      for (int m = 0; m < BRK_N_MOVERS; m++) {
        mv[m].x0 = __shfl_sync(0xffffffffu, mine.x0, m); mv[m].x1 = __shfl_sync(0xffffffffu, mine.x1, m);
        mv[m].y0 = __shfl_sync(0xffffffffu, mine.y0, m); mv[m].y1 = __shfl_sync(0xffffffffu, mine.y1, m);
        mv[m].gray = __shfl_sync(0xffffffffu, mine.gray, m);
      }
      const uint8_t *__restrict__ base0 = a.base[0];
      const uint32_t vm = mm;
      while (mm) {
        const int m = __ffs(mm) - 1;
        mm &= mm - 1;
        const int dx0 = __shfl_sync(0xffffffffu, fx0, m), dx1 = __shfl_sync(0xffffffffu, fx1, m);
        const int dy0 = __shfl_sync(0xffffffffu, fy0, m), dy1 = __shfl_sync(0xffffffffu, fy1, m);
        /* the movers that reach into the source window of this footprint, and whether the wall does */
        const int sx0 = cp.xs0[dx0], sx1 = cp.xs0[dx1] + TX, sy0 = cp.ys0[dy0], sy1 = cp.ys0[dy1] + TY;
        uint32_t near = 0;
#pragma unroll
        for (int q = 0; q < BRK_N_MOVERS; q++)
          if (((vm >> q) & 1u) && mv[q].x0 < sx1 && mv[q].x1 > sx0 && mv[q].y0 < sy1 && mv[q].y1 > sy0) near |= 1u << q;
        const bool wall = sy0 < wy1 && sy1 > wy0;
        const int ncol = dx1 - dx0 + 1;
        const int lg = ncol > 16 ? 5 : ncol > 8 ? 4 : ncol > 4 ? 3 : ncol > 2 ? 2 : ncol > 1 ? 1 : 0;
        const int c = lane & ((1 << lg) - 1), cpl = 1 << lg, rstep = 32 >> lg;
        for (int dxb = dx0; dxb <= dx1; dxb += cpl) {
          const int dx = dxb + c;
          if (dx > dx1) continue;
          for (int dy = dy0 + (lane >> lg); dy <= dy1; dy += rstep)
            out[dy * dw + dx] = brk_direct_pixel<TX, TY>(A, *plan, base0, R + BRK_W(alive), mv, near, wall, dx, dy);
        }
      }
    }
  }
}

} /* namespace tbxk */
#endif
