/* tbx_breakout.h -- Breakout transition, new_game and draw list, one env per call.
 *
 * Replaces what Toybox('breakout').apply_ale_action / new_game / get_state reach in ctoybox
 * (reference call sites toybox/envs/atari/base.py:126,153,109; state fields
 * toybox/interventions/breakout.py:49-68, :132 Paddle, :198 Brick, :276 Ball; constants and the initial
 * state from toybox/interventions/defaults/breakout_{config,state}_default.json).
 * All f64 arithmetic goes through tbx_d* (round-to-nearest, no FMA contraction); trig comes from the
 * host-evaluated tables in BrkCfg so CPU and GPU use identical bits.
 */
#ifndef TBX_BREAKOUT_H
#define TBX_BREAKOUT_H
#include "tbx_common.h"

#define BRK_W(f) TBX_W(BrkRec, f)
#define BRK_LEFT_X 12.0
#define BRK_RIGHT_X 228.0
#define BRK_TOP_Y 25.0
#define BRK_BOTTOM_Y 160.0
#define BRK_FRAME_Y 13
#define BRK_PADDLE_HALF_H 1.5

/* draw-list slots (fixed positions; unused slots are empty prims) */
#define BRK_SLOT_FRAME 0                                     /* 3 rects */
#define BRK_SLOT_SCORE 3                                     /* TBX_MAX_DIGITS */
#define BRK_SLOT_LIVES (BRK_SLOT_SCORE + TBX_MAX_DIGITS)     /* TBX_MAX_DIGITS */
#define BRK_SLOT_BRICKS (BRK_SLOT_LIVES + TBX_MAX_DIGITS)    /* TBX_BRK_MAX_BRICKS */
#define BRK_SLOT_PADDLE (BRK_SLOT_BRICKS + TBX_BRK_MAX_BRICKS)
#define BRK_SLOT_BALLS (BRK_SLOT_PADDLE + 1)                 /* TBX_BRK_MAX_BALLS */
#define BRK_N_SLOTS (BRK_SLOT_BALLS + TBX_BRK_MAX_BALLS)
#define BRK_N_STATIC 3 /* leading slots that depend on the config only */
#define BRK_N_GROUPS 3
/* HUD digits (one colour) | bricks (parallel iff the table's rectangles are disjoint) | paddle, balls (in order) */
TBX_HD void brk_group(int g, const uint32_t *R, const BrkTable *tables, int base, int &b, int &e, int &mode) {
  const BrkTable &T = tables[(int32_t)R[TBX_W(TbxHdr, tbl)]];
  if (g == 0) { b = BRK_SLOT_SCORE; e = BRK_SLOT_BRICKS; mode = TBX_GROUP_PARALLEL | (T.hud_clear ? TBX_GROUP_NOSYNC : 0); }
  else if (g == 1) {
    b = BRK_SLOT_BRICKS; e = BRK_SLOT_PADDLE; mode = T.disjoint ? TBX_GROUP_PARALLEL : TBX_GROUP_SERIAL;
    if (base == 1) { /* delta rendering: nothing to paint while every brick is alive */
      uint32_t dead = 0;
      for (int k = 0; k < 5; k++) dead |= T.all_mask[k] & ~R[TBX_W(BrkRec, alive) + k];
      if (!dead) e = b;
    }
  }
  else { b = BRK_SLOT_PADDLE; e = BRK_N_SLOTS; mode = TBX_GROUP_SERIAL; }
}

TBX_HD double brk_vmag(double vx, double vy) { return tbx_dsqrt(tbx_dadd(tbx_dmul(vx, vx), tbx_dmul(vy, vy))); }

/* park a fresh ball on a random start position (config ball_start_positions), slow speed */
TBX_HD void brk_start_ball(const TbxAcc &S, const BrkCfg &c, TbxRng &rng) {
  uint32_t i = tbx_rng_index(rng, (uint32_t)c.n_starts);
  S.sti(BRK_W(n_balls), 1);
  int w = BRK_W(ball);
  S.std_(w + 0, c.start_x[i]);
  S.std_(w + 2, c.start_y[i]);
  S.std_(w + 4, tbx_dmul(c.ball_speed_slow, c.start_cos[i]));
  S.std_(w + 6, tbx_dmul(c.ball_speed_slow, c.start_sin[i]));
}

/* new_game: child rng = two draws of the env's simulator rng, bricks from the default table */
TBX_HD void brk_new_game(const TbxAcc &S, const BrkCfg &c) {
  TbxRng sim = tbx_rng_load(S, TBX_HW(sim_rand));
  TbxRng rng;
  rng.s0 = tbx_rng_next_u64(sim);
  rng.s1 = tbx_rng_next_u64(sim);
  tbx_rng_store(S, TBX_HW(sim_rand), sim);
  S.sti(TBX_HW(lives), c.start_lives);
  S.sti(TBX_HW(score), 0);
  S.sti(TBX_HW(level), 1);
  S.sti(TBX_HW(prev_score), 0);
  S.sti(TBX_HW(ep_len), 0);
  S.sti(TBX_HW(ep_return), 0);
  S.sti(TBX_HW(tbl), c.default_tbl);
  S.sti(BRK_W(is_dead), 1);
  S.sti(BRK_W(reset), 1);
  S.std_(BRK_W(paddle_px), 120.0);
  S.std_(BRK_W(paddle_py), 143.0);
  S.std_(BRK_W(paddle_vx), 0.0);
  S.std_(BRK_W(paddle_vy), 0.0);
  S.std_(BRK_W(paddle_width), 24.0);
  S.std_(BRK_W(paddle_speed), 4.0);
  S.std_(BRK_W(ball_radius), 2.0);
  int nb = 18 * c.n_rows;
  for (int k = 0; k < 5; k++) {
    int lo = 32 * k, n = nb - lo;
    S.st(BRK_W(alive) + k, n >= 32 ? 0xffffffffu : n <= 0 ? 0u : ((1u << n) - 1u));
  }
  for (int b = 1; b < TBX_BRK_MAX_BALLS; b++)
    for (int k = 0; k < 4; k++) S.std_(BRK_W(ball) + 8 * b + 2 * k, 0.0);
  brk_start_ball(S, c, rng);
  tbx_rng_store(S, TBX_HW(rand), rng);
}

struct BrkBall { double px, py, vx, vy; };

/* one time slice of one ball; returns false when the ball left the field through the bottom */
TBX_HD bool brk_slice(const TbxAcc &S, const BrkCfg &c, const BrkTable &T, BrkBall &b, double dt, double r,
                      double paddle_x, double paddle_y, double paddle_width, uint32_t *alive, int &score) {
  b.px = tbx_dadd(b.px, tbx_dmul(b.vx, dt));
  b.py = tbx_dadd(b.py, tbx_dmul(b.vy, dt));
  double xl = tbx_dsub(b.px, r), xr = tbx_dadd(b.px, r), yt = tbx_dsub(b.py, r), yb = tbx_dadd(b.py, r);
  if (xl < BRK_LEFT_X && b.vx < 0.0) b.vx = -b.vx;
  if (xr > BRK_RIGHT_X && b.vx > 0.0) b.vx = -b.vx;
  if (yt < BRK_TOP_Y && b.vy < 0.0) b.vy = -b.vy;
  if (b.vy > 0.0) {
    double half = tbx_dmul(paddle_width, 0.5);
    double pl = tbx_dsub(paddle_x, half), pr = tbx_dadd(paddle_x, half);
    double pt = tbx_dsub(paddle_y, BRK_PADDLE_HALF_H), pb = tbx_dadd(paddle_y, BRK_PADDLE_HALF_H);
    if (yb >= pt && yt <= pb && xr >= pl && xl <= pr) {
      int nseg = c.paddle_discrete_segments;
      double frac = tbx_ddiv(tbx_dsub(b.px, pl), paddle_width);
      int seg = tbx_d2i(floor(tbx_dmul(frac, (double)nseg)));
      if (seg < 0) seg = 0;
      if (seg > nseg - 1) seg = nseg - 1;
      if (seg > TBX_BRK_MAX_SEGS - 1) seg = TBX_BRK_MAX_SEGS - 1;
      if (seg < 0) seg = 0;
      double speed = brk_vmag(b.vx, b.vy);
      b.vx = tbx_dmul(speed, c.seg_cos[seg]);
      b.vy = -tbx_dmul(speed, c.seg_sin[seg]);
    }
  }
  /* bricks: first alive brick in index order whose box meets the ball's box */
  if (xr > T.bb_x0 && xl < T.bb_x1 && yb > T.bb_y0 && yt < T.bb_y1) {
    int hit_i = -1;
    if (T.grid) {
      /* a regular column-major grid: only the cells around the ball can meet its box.  The cell range is a superset (one cell
       * of margin absorbs the rounding of the multiply), every candidate takes the exact test, and columns then rows ascending
       * IS index order: the same first hit as the full scan below. */
      int c0 = tbx_d2i(tbx_dmul(tbx_dsub(xl, T.gx0), T.ginv_w)) - 1, c1 = tbx_d2i(tbx_dmul(tbx_dsub(xr, T.gx0), T.ginv_w)) + 1;
      int r0 = tbx_d2i(tbx_dmul(tbx_dsub(yt, T.gy0), T.ginv_h)) - 1, r1 = tbx_d2i(tbx_dmul(tbx_dsub(yb, T.gy0), T.ginv_h)) + 1;
      if (c0 < 0) c0 = 0;
      if (r0 < 0) r0 = 0;
      if (c1 > T.g_ncols - 1) c1 = T.g_ncols - 1;
      if (r1 > T.g_nrows - 1) r1 = T.g_nrows - 1;
      for (int c = c0; c <= c1 && hit_i < 0; c++)
        for (int r = r0; r <= r1; r++) {
          const int i = c * T.g_nrows + r, k = i >> 5;
          const uint32_t w = k == 0 ? alive[0] : k == 1 ? alive[1] : k == 2 ? alive[2] : k == 3 ? alive[3] : alive[4];
          if (!((w >> (i & 31)) & 1u)) continue;
          if (xr > T.px[i] && xl < T.x1[i] && yb > T.py[i] && yt < T.y1[i]) { hit_i = i; break; }
        }
    } else {
      for (int k = 0; k < 5 && hit_i < 0; k++) {
        uint32_t m = alive[k];
        while (m) {
          int bit = tbx_ffs(m) - 1, i = 32 * k + bit;
          m &= m - 1;
          if (xr > T.px[i] && xl < T.x1[i] && yb > T.py[i] && yt < T.y1[i]) { hit_i = i; break; }
        }
      }
    }
    if (hit_i >= 0) {
      const int i = hit_i, k = i >> 5, bit = i & 31;
      const uint32_t destr = k == 0 ? T.destructible[0] : k == 1 ? T.destructible[1] : k == 2 ? T.destructible[2] : k == 3 ? T.destructible[3] : T.destructible[4];
      if ((destr >> bit) & 1u) {
        for (int q = 0; q < 5; q++) if (q == k) alive[q] &= ~(1u << bit);
        score += T.points[i];
      }
      b.vy = -b.vy;
      if (T.depth[i] >= c.ball_speed_row_depth) {
        double mag = brk_vmag(b.vx, b.vy);
        if (mag > 0.0) {
          double ux = tbx_ddiv(b.vx, mag), uy = tbx_ddiv(b.vy, mag);
          b.vx = tbx_dmul(ux, c.ball_speed_fast);
          b.vy = tbx_dmul(uy, c.ball_speed_fast);
        }
      }
    }
  }
  (void)S;
  return !(yt > BRK_BOTTOM_Y);
}

TBX_HD void brk_step(const TbxAcc &S, const BrkCfg &c, const BrkTable *tables, int in) {
  /* paddle */
  double paddle_speed = S.ldd(BRK_W(paddle_speed)), paddle_width = S.ldd(BRK_W(paddle_width));
  double vx = 0.0;
  bool left = (in & TBX_IN_LEFT) != 0, right = (in & TBX_IN_RIGHT) != 0;
  if (left && !right) vx = -paddle_speed;
  if (right && !left) vx = paddle_speed;
  S.std_(BRK_W(paddle_vx), vx);
  S.std_(BRK_W(paddle_vy), 0.0);
  double half = tbx_dmul(paddle_width, 0.5), lo = tbx_dadd(BRK_LEFT_X, half), hi = tbx_dsub(BRK_RIGHT_X, half);
  double paddle_x = tbx_dadd(S.ldd(BRK_W(paddle_px)), vx);
  if (paddle_x > hi) paddle_x = hi;
  if (paddle_x < lo) paddle_x = lo;
  S.std_(BRK_W(paddle_px), paddle_x);
  int lives = S.ldi(TBX_HW(lives));
  if (lives <= 0) return; /* game over: only the paddle moves */
  int n_balls = S.ldi(BRK_W(n_balls));
  if (S.ldi(BRK_W(is_dead))) { /* waiting for FIRE */
    if (in & TBX_IN_BUTTON1) {
      S.sti(BRK_W(is_dead), 0);
      S.sti(BRK_W(reset), 0);
      if (n_balls == 0) {
        TbxRng rng = tbx_rng_load(S, TBX_HW(rand));
        brk_start_ball(S, c, rng);
        tbx_rng_store(S, TBX_HW(rand), rng);
      }
    }
    return;
  }
  const BrkTable &T = tables[S.ldi(TBX_HW(tbl))];
  double paddle_y = S.ldd(BRK_W(paddle_py)), r = S.ldd(BRK_W(ball_radius));
  uint32_t alive[5];
  for (int k = 0; k < 5; k++) alive[k] = S.ld(BRK_W(alive) + k);
  int score = S.ldi(TBX_HW(score));
  if (n_balls > TBX_BRK_MAX_BALLS) n_balls = TBX_BRK_MAX_BALLS;
  /* balls, each sub-stepped so it never travels more than its radius per slice */
  uint32_t keep = 0;
  double lim = r < 0.25 ? 0.25 : r;
  for (int bi = 0; bi < n_balls; bi++) {
    int w = BRK_W(ball) + 8 * bi;
    BrkBall b;
    b.px = S.ldd(w); b.py = S.ldd(w + 2); b.vx = S.ldd(w + 4); b.vy = S.ldd(w + 6);
    double t_left = 1.0;
    bool ok = true;
    for (int it = 0; it < 64 && ok && t_left > 0.0; it++) {
      double speed = brk_vmag(b.vx, b.vy);
      double dt = speed > lim ? tbx_ddiv(lim, speed) : 1.0;
      if (dt > t_left) dt = t_left;
      ok = brk_slice(S, c, T, b, dt, r, paddle_x, paddle_y, paddle_width, alive, score);
      t_left = tbx_dsub(t_left, dt);
    }
    S.std_(w, b.px); S.std_(w + 2, b.py); S.std_(w + 4, b.vx); S.std_(w + 6, b.vy);
    if (ok) keep |= 1u << bi;
  }
  /* drop lost balls; losing the last one costs a life and parks a fresh ball */
  int n = tbx_popc(keep);
  if (n != n_balls) {
    int dst = 0;
    for (int bi = 0; bi < n_balls; bi++)
      if ((keep >> bi) & 1u) {
        if (dst != bi)
          for (int k = 0; k < 8; k++) S.st(BRK_W(ball) + 8 * dst + k, S.ld(BRK_W(ball) + 8 * bi + k));
        dst++;
      }
    S.sti(BRK_W(n_balls), n);
  }
  if (n == 0) {
    S.sti(TBX_HW(lives), lives - 1);
    S.sti(BRK_W(is_dead), 1);
    S.sti(BRK_W(reset), 1);
    TbxRng rng = tbx_rng_load(S, TBX_HW(rand));
    brk_start_ball(S, c, rng);
    tbx_rng_store(S, TBX_HW(rand), rng);
  }
  /* board cleared -> refill, next level */
  uint32_t left_to_break = 0;
  for (int k = 0; k < 5; k++) left_to_break |= alive[k] & T.destructible[k];
  if (left_to_break == 0 && T.n_bricks > 0) {
    for (int k = 0; k < 5; k++) alive[k] = T.all_mask[k];
    S.sti(TBX_HW(level), S.ldi(TBX_HW(level)) + 1);
  }
  for (int k = 0; k < 5; k++) S.st(BRK_W(alive) + k, alive[k]);
  S.sti(TBX_HW(score), score);
}

/* draw list, painted in slot order: frame, score, lives, bricks, paddle, balls.
 * `R` is the env's record as contiguous words (stride 1). */
TBX_HD TbxPrim brk_prim(const uint32_t *R, const BrkCfg &c, const BrkTable *tables, int slot) {
  TbxAcc S; S.p = const_cast<uint32_t *>(R); S.stride = 1;
  if (slot < BRK_SLOT_SCORE) {
    if (slot == 0) return tbx_prim_rect(c.frame_color, 0, BRK_FRAME_Y, TBX_BRK_W, 12);
    if (slot == 1) return tbx_prim_rect(c.frame_color, 0, BRK_FRAME_Y, 12, TBX_BRK_H - BRK_FRAME_Y);
    return tbx_prim_rect(c.frame_color, TBX_BRK_W - 12, BRK_FRAME_Y, 12, TBX_BRK_H - BRK_FRAME_Y);
  }
  if (slot < BRK_SLOT_LIVES) return tbx_prim_digit(c.frame_color, 108, 2, S.ldi(TBX_HW(score)), 4, 2, slot - BRK_SLOT_SCORE);
  if (slot < BRK_SLOT_BRICKS) return tbx_prim_digit(c.frame_color, 180, 2, S.ldi(TBX_HW(lives)), 4, 2, slot - BRK_SLOT_LIVES);
  if (slot < BRK_SLOT_PADDLE) {
    int i = slot - BRK_SLOT_BRICKS;
    const BrkTable &T = tables[S.ldi(TBX_HW(tbl))];
    if (i >= T.n_bricks || !((S.ld(BRK_W(alive) + (i >> 5)) >> (i & 31)) & 1u)) return tbx_prim_none();
    return tbx_prim_rect(T.color[i], T.ix[i], T.iy[i], T.iw[i], T.ih[i]);
  }
  if (slot == BRK_SLOT_PADDLE) {
    double pw = S.ldd(BRK_W(paddle_width)), half = tbx_dmul(pw, 0.5);
    return tbx_prim_rect(c.paddle_color, tbx_d2i(tbx_dsub(S.ldd(BRK_W(paddle_px)), half)),
                         tbx_d2i(tbx_dsub(S.ldd(BRK_W(paddle_py)), BRK_PADDLE_HALF_H)), tbx_d2i(pw), 3);
  }
  int bi = slot - BRK_SLOT_BALLS;
  if (bi >= S.ldi(BRK_W(n_balls))) return tbx_prim_none();
  double r = S.ldd(BRK_W(ball_radius)), d = tbx_dmul(r, 2.0);
  int w = BRK_W(ball) + 8 * bi;
  return tbx_prim_rect(c.ball_color, tbx_d2i(tbx_dsub(S.ldd(w), r)), tbx_d2i(tbx_dsub(S.ldd(w + 2), r)), tbx_d2i(d), tbx_d2i(d));
}

/* Delta rendering.  Base frame 0 holds the static slots only; base frame 1 additionally holds every brick of the
 * config's default table alive.  An env on that table (when the table is `delta_ok`: disjoint bricks lying on
 * pure background, clear of the HUD) is drawn as base 1 plus its DEAD bricks painted in the background colour --
 * the same pixels the painter's algorithm yields, with work proportional to what differs from a fresh game. */
TBX_HD int brk_base_id(const uint32_t *R, const BrkCfg &c, const BrkTable *tables) {
  int t = (int32_t)R[TBX_HW(tbl)];
  return (t == c.default_tbl && tables[t].delta_ok) ? 1 : 0;
}
TBX_HD bool brk_alive_bit(const uint32_t *R, int i) { return (R[BRK_W(alive) + (i >> 5)] >> (i & 31)) & 1u; }
/* brick b sits directly below brick a: same columns, a's bottom edge is b's top edge */
TBX_HD bool brk_stacked(const BrkTable &T, int a, int b) { return T.ix[a] == T.ix[b] && T.iw[a] == T.iw[b] && T.iy[a] + T.ih[a] == T.iy[b]; }
TBX_HD TbxPrim brk_prim_delta(const uint32_t *R, const BrkCfg &c, const BrkTable *tables, int slot, int base) {
  if (base == 1 && slot >= BRK_SLOT_BRICKS && slot < BRK_SLOT_PADDLE) {
    int i = slot - BRK_SLOT_BRICKS;
    const BrkTable &T = tables[(int32_t)R[TBX_HW(tbl)]];
    if (i >= T.n_bricks || brk_alive_bit(R, i)) return tbx_prim_none();
    /* Dead bricks that sit directly on top of each other (consecutive slots: bricks are column-major) are ONE
     * background rectangle, emitted by the topmost of the run: the ball tunnels vertical channels, so a worn wall is
     * a few tall holes rather than dozens of brick-sized ones.  Same pixels: the run covers exactly its bricks. */
    if (i > 0 && brk_stacked(T, i - 1, i) && !brk_alive_bit(R, i - 1)) return tbx_prim_none();
    int h = T.ih[i];
    for (int j = i + 1; j < T.n_bricks && brk_stacked(T, j - 1, j) && !brk_alive_bit(R, j); j++) h += T.ih[j];
    return tbx_prim_rect(c.bg_color, T.ix[i], T.iy[i], T.iw[i], h);
  }
  return brk_prim(R, c, tables, slot);
}

#endif
