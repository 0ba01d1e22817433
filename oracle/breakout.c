/* breakout.c -- CPU ORACLE (test infrastructure, not product code).
 *
 * Restates the Breakout engine behind Toybox('breakout').apply_ale_action / get_state
 * (reference call sites toybox/envs/atari/base.py:126,109; state schema
 * toybox/interventions/breakout.py:49-68 Breakout, :132 Paddle, :198 Brick, :276 Ball;
 * constants and the initial state from toybox/interventions/defaults/breakout_*_default.json).
 *
 * Pinned by the fixtures ([FIX], checked in tests/test_oracle_golden.py): the whole new_game()
 * state incl. RNG lineage and the 2-draw ball-start choice, brick layout (column-major), paddle
 * (120,143).  The per-frame rules B1..B7 below are this project's restatement: PARITY UNPINNED.
 *
 * All f64 arithmetic is written one IEEE operation per statement so the CUDA path can mirror it
 * with __dmul_rn/__dadd_rn (no FMA contraction on either side; build with -ffp-contract=off).
 */
#include "tbo.h"
#include <math.h>
#include <string.h>

#define LEFT_X 12.0
#define RIGHT_X 228.0
#define TOP_Y 25.0
#define BOTTOM_Y 160.0
#define FRAME_Y 13
#define PADDLE_HALF_H 1.5
#define DEG2RAD 0.017453292519943295   /* PI/180 as f64, i.e. f64::to_radians */

static int d2i(double v) {            /* `as i32`: truncate, saturate, NaN -> 0 */
  if (!(v == v)) return 0;
  if (v >= 536870912.0) return 536870912;
  if (v <= -536870912.0) return -536870912;
  return (int)v;
}

void tbo_brk_default_cfg(tbo_brk_cfg *c) {
  static const tbo_color rows[6] = { {200, 72, 72, 255}, {198, 108, 58, 255}, {180, 122, 48, 255},
                                     {162, 162, 42, 255}, {72, 160, 72, 255}, {66, 72, 200, 255} };
  static const int32_t scores[6] = { 7, 7, 4, 4, 1, 1 };
  static const tbo_ball_start starts[4] = { {24, 80, 30}, {120, 80, 30}, {120, 80, 150}, {216, 80, 150} };
  memset(c, 0, sizeof *c);
  c->bg_color = (tbo_color){0, 0, 0, 255};
  c->frame_color = (tbo_color){144, 144, 144, 255};
  c->paddle_color = (tbo_color){200, 72, 72, 255};
  c->ball_color = (tbo_color){200, 72, 72, 255};
  c->n_rows = 6;
  for (int i = 0; i < 6; i++) { c->row_colors[i] = rows[i]; c->row_scores[i] = scores[i]; }
  c->start_lives = 5; c->paddle_discrete_segments = 5; c->ball_speed_row_depth = 3;
  c->ball_speed_slow = 2.0; c->ball_speed_fast = 4.0;
  c->n_starts = 4;
  for (int i = 0; i < 4; i++) c->ball_start_positions[i] = starts[i];
  tbo_rng_seed(&c->rand, 13);
}

static void start_ball(const tbo_brk_cfg *c, tbo_brk_state *s) {
  uint32_t i = tbo_rng_index(&s->rand, (uint32_t)c->n_starts);
  const tbo_ball_start *p = &c->ball_start_positions[i];
  double rad = p->angle_degrees * DEG2RAD;
  s->n_balls = 1;
  s->balls[0].position.x = p->x; s->balls[0].position.y = p->y;
  s->balls[0].velocity.x = c->ball_speed_slow * cos(rad);
  s->balls[0].velocity.y = c->ball_speed_slow * sin(rad);
}

void tbo_brk_new_game(tbo_brk_cfg *c, tbo_brk_state *s) {
  memset(s, 0, sizeof *s);
  s->rand = tbo_rng_child(&c->rand);                 /* [FIX] App. A.2 */
  s->lives = c->start_lives; s->score = 0; s->level = 1; s->is_dead = 1; s->reset = 1;
  s->paddle.position.x = 120.0; s->paddle.position.y = 143.0;
  s->paddle_width = 24.0; s->paddle_speed = 4.0; s->ball_radius = 2.0;
  s->n_bricks = 18 * c->n_rows;
  for (int col = 0; col < 18; col++)
    for (int row = 0; row < c->n_rows; row++) {      /* [FIX] column-major: bricks[i].col = i / n_rows */
      tbo_brk_brick *b = &s->bricks[col * c->n_rows + row];
      b->position.x = 12.0 + 12.0 * col; b->position.y = 43.0 + 4.0 * row;
      b->size.x = 12.0; b->size.y = 4.0; b->color = c->row_colors[row];
      b->points = c->row_scores[row]; b->depth = c->n_rows - 1 - row; b->row = row; b->col = col;
      b->alive = 1; b->destructible = 1;
    }
  start_ball(c, s);                                   /* [FIX] App. A.3: two draws, index 2 */
}

static double vmag(double vx, double vy) {
  double a = vx * vx; double b = vy * vy; double q = a + b;
  return sqrt(q);
}

/* B4: one time slice for ball `bi`; returns 0 when the ball left the field */
static int slice(const tbo_brk_cfg *c, tbo_brk_state *s, int bi, double dt) {
  tbo_brk_body *b = &s->balls[bi];
  double r = s->ball_radius;
  double mx = b->velocity.x * dt; double my = b->velocity.y * dt;
  b->position.x = b->position.x + mx; b->position.y = b->position.y + my;
  double xl = b->position.x - r, xr = b->position.x + r, yt = b->position.y - r, yb = b->position.y + r;
  /* walls */
  if (xl < LEFT_X && b->velocity.x < 0.0) b->velocity.x = -b->velocity.x;
  if (xr > RIGHT_X && b->velocity.x > 0.0) b->velocity.x = -b->velocity.x;
  if (yt < TOP_Y && b->velocity.y < 0.0) b->velocity.y = -b->velocity.y;
  /* paddle: the contact offset picks one of `paddle_discrete_segments` outgoing angles, 150 deg (left) .. 30 deg */
  if (b->velocity.y > 0.0) {
    double half = s->paddle_width * 0.5;
    double pl = s->paddle.position.x - half, pr = s->paddle.position.x + half;
    double pt = s->paddle.position.y - PADDLE_HALF_H, pb = s->paddle.position.y + PADDLE_HALF_H;
    if (yb >= pt && yt <= pb && xr >= pl && xl <= pr) {
      int nseg = c->paddle_discrete_segments;
      double hit = b->position.x - pl;
      double frac = hit / s->paddle_width;
      double fs = frac * (double)nseg;
      int seg = d2i(floor(fs));
      if (seg < 0) seg = 0;
      if (seg > nseg - 1) seg = nseg - 1;
      double ang = nseg > 1 ? 150.0 - (double)seg * (120.0 / (double)(nseg - 1)) : 90.0;
      double rad = ang * DEG2RAD;
      double speed = vmag(b->velocity.x, b->velocity.y);
      b->velocity.x = speed * cos(rad);
      b->velocity.y = -(speed * sin(rad));
    }
  }
  /* bricks: first alive brick (index order) whose box meets the ball's box */
  for (int i = 0; i < s->n_bricks; i++) {
    tbo_brk_brick *k = &s->bricks[i];
    if (!k->alive) continue;
    double bx1 = k->position.x + k->size.x, by1 = k->position.y + k->size.y;
    if (xr > k->position.x && xl < bx1 && yb > k->position.y && yt < by1) {
      if (k->destructible) { k->alive = 0; s->score += k->points; }
      b->velocity.y = -b->velocity.y;
      if (k->depth >= c->ball_speed_row_depth) {
        double m = vmag(b->velocity.x, b->velocity.y);
        if (m > 0.0) {
          double ux = b->velocity.x / m, uy = b->velocity.y / m;
          b->velocity.x = ux * c->ball_speed_fast; b->velocity.y = uy * c->ball_speed_fast;
        }
      }
      break;
    }
  }
  return !(yt > BOTTOM_Y);
}

void tbo_brk_step(const tbo_brk_cfg *c, tbo_brk_state *s, int in) {
  /* B1 paddle */
  double vx = 0.0;
  int left = (in & TBO_IN_LEFT) != 0, right = (in & TBO_IN_RIGHT) != 0;
  if (left && !right) vx = -s->paddle_speed;
  if (right && !left) vx = s->paddle_speed;
  s->paddle.velocity.x = vx; s->paddle.velocity.y = 0.0;
  {
    double half = s->paddle_width * 0.5, lo = LEFT_X + half, hi = RIGHT_X - half;
    double x = s->paddle.position.x + vx;
    if (x > hi) x = hi;
    if (x < lo) x = lo;
    s->paddle.position.x = x;
  }
  /* B2 game over: only the paddle moves */
  if (s->lives <= 0) return;
  /* B3 waiting for FIRE */
  if (s->is_dead) {
    if (in & TBO_IN_BUTTON1) { s->is_dead = 0; s->reset = 0; if (s->n_balls == 0) start_ball(c, s); }
    return;
  }
  /* B4 balls, each sub-stepped so it never travels more than its radius per slice */
  int keep[TBO_BRK_MAX_BALLS];
  for (int bi = 0; bi < s->n_balls; bi++) {
    double t_left = 1.0; int alive = 1;
    double lim = s->ball_radius < 0.25 ? 0.25 : s->ball_radius;
    for (int it = 0; it < 64 && alive && t_left > 0.0; it++) {
      double speed = vmag(s->balls[bi].velocity.x, s->balls[bi].velocity.y);
      double dt = speed > lim ? lim / speed : 1.0;
      if (dt > t_left) dt = t_left;
      alive = slice(c, s, bi, dt);
      t_left = t_left - dt;
    }
    keep[bi] = alive;
  }
  /* B5 drop lost balls; losing the last one costs a life and parks a fresh ball */
  int n = 0;
  for (int bi = 0; bi < s->n_balls; bi++) if (keep[bi]) s->balls[n++] = s->balls[bi];
  s->n_balls = n;
  if (n == 0) { s->lives -= 1; s->is_dead = 1; s->reset = 1; start_ball(c, s); }
  /* B6 board cleared -> refill, next level */
  int left_to_break = 0;
  for (int i = 0; i < s->n_bricks; i++) left_to_break += (s->bricks[i].alive && s->bricks[i].destructible);
  if (left_to_break == 0 && s->n_bricks > 0) {
    for (int i = 0; i < s->n_bricks; i++) s->bricks[i].alive = 1;
    s->level += 1;
  }
}

/* B7 draw list, painted in order */
void tbo_brk_render(const tbo_brk_cfg *c, const tbo_brk_state *s, uint8_t *rgba) {
  tbo_canvas cv = { TBO_BRK_W, TBO_BRK_H, rgba };
  tbo_clear(&cv, c->bg_color);
  tbo_rect(&cv, c->frame_color, 0, FRAME_Y, TBO_BRK_W, 12);
  tbo_rect(&cv, c->frame_color, 0, FRAME_Y, 12, TBO_BRK_H - FRAME_Y);
  tbo_rect(&cv, c->frame_color, TBO_BRK_W - 12, FRAME_Y, 12, TBO_BRK_H - FRAME_Y);
  tbo_digits(&cv, c->frame_color, 108, 2, s->score, 4, 2);
  tbo_digits(&cv, c->frame_color, 180, 2, s->lives, 4, 2);
  for (int i = 0; i < s->n_bricks; i++) {
    const tbo_brk_brick *k = &s->bricks[i];
    if (k->alive) tbo_rect(&cv, k->color, d2i(k->position.x), d2i(k->position.y), d2i(k->size.x), d2i(k->size.y));
  }
  {
    double half = s->paddle_width * 0.5;
    tbo_rect(&cv, c->paddle_color, d2i(s->paddle.position.x - half), d2i(s->paddle.position.y - PADDLE_HALF_H),
             d2i(s->paddle_width), 3);
  }
  for (int bi = 0; bi < s->n_balls; bi++) {
    double r = s->ball_radius, d = r * 2.0;
    tbo_rect(&cv, c->ball_color, d2i(s->balls[bi].position.x - r), d2i(s->balls[bi].position.y - r), d2i(d), d2i(d));
  }
}
