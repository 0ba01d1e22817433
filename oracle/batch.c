#define _POSIX_C_SOURCE 199309L
/* batch.c -- CPU ORACLE (test infrastructure, not product code): array-of-envs drivers around the
 * scalar engines, with the env-level semantics of toybox/envs/atari/base.py:115-156 (reward =
 * max(score - prev, 0), done = lives <= 0) and the auto-reset of
 * baselines/baselines/common/vec_env/subproc_vec_env.py:11-15.  OpenMP only spreads independent envs. */
#include "tbo.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { G_BREAKOUT = 0, G_AMIDAR = 1, G_SPACE_INVADERS = 2 };
enum { M_RGBA = 0, M_RGB = 1, M_GRAY = 2, M_GRAY84 = 3 };

static size_t state_size(int game) { return tbo_sizeof(game == G_BREAKOUT ? 1 : game == G_AMIDAR ? 5 : 3); }
static void dims(int game, int *w, int *h) {
  if (game == G_BREAKOUT) { *w = TBO_BRK_W; *h = TBO_BRK_H; }
  else if (game == G_AMIDAR) { *w = TBO_AMI_W; *h = TBO_AMI_H; }
  else { *w = TBO_SI_W; *h = TBO_SI_H; }
}
static void new_game1(int game, const void *cfg, void *st, tbo_rng *sim) {
  /* the simulator-level rand is per env; everything else in the config is shared */
  if (game == G_BREAKOUT) { tbo_brk_cfg c = *(const tbo_brk_cfg *)cfg; c.rand = *sim; tbo_brk_new_game(&c, (tbo_brk_state *)st); *sim = c.rand; }
  else if (game == G_AMIDAR) { tbo_ami_cfg c = *(const tbo_ami_cfg *)cfg; c.rand = *sim; tbo_ami_new_game(&c, (tbo_ami_state *)st); *sim = c.rand; }
  else { tbo_si_cfg c = *(const tbo_si_cfg *)cfg; c.rand = *sim; tbo_si_new_game(&c, (tbo_si_state *)st); *sim = c.rand; }
}
static void step1(int game, const void *cfg, void *st, int input) {
  if (game == G_BREAKOUT) tbo_brk_step((const tbo_brk_cfg *)cfg, (tbo_brk_state *)st, input);
  else if (game == G_AMIDAR) tbo_ami_step((const tbo_ami_cfg *)cfg, (tbo_ami_state *)st, input);
  else tbo_si_step((const tbo_si_cfg *)cfg, (tbo_si_state *)st, input);
}
static void render1(int game, const void *cfg, const void *st, uint8_t *rgba) {
  if (game == G_BREAKOUT) tbo_brk_render((const tbo_brk_cfg *)cfg, (const tbo_brk_state *)st, rgba);
  else if (game == G_AMIDAR) tbo_ami_render((const tbo_ami_cfg *)cfg, (const tbo_ami_state *)st, rgba);
  else tbo_si_render((const tbo_si_cfg *)cfg, (const tbo_si_state *)st, rgba);
}
static void score_lives(int game, const void *st, int32_t *score, int32_t *lives) {
  if (game == G_BREAKOUT) { *score = ((const tbo_brk_state *)st)->score; *lives = ((const tbo_brk_state *)st)->lives; }
  else if (game == G_AMIDAR) { *score = ((const tbo_ami_state *)st)->score; *lives = ((const tbo_ami_state *)st)->lives; }
  else { *score = ((const tbo_si_state *)st)->score; *lives = ((const tbo_si_state *)st)->lives; }
}

void tbo_batch_seed(tbo_rng *sim, const uint32_t *seeds, int n) { for (int i = 0; i < n; i++) tbo_rng_seed(&sim[i], seeds[i]); }

void tbo_batch_new_game(int game, const void *cfg, void *states, tbo_rng *sim, int n) {
  size_t ss = state_size(game);
  #pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) new_game1(game, cfg, (char *)states + ss * i, &sim[i]);
}

/* returns -1 if any action id is invalid (nothing is stepped in that case) */
int tbo_batch_step(int game, const void *cfg, void *states, tbo_rng *sim, int n, const int32_t *ale_actions,
                   int auto_reset, int32_t *prev_score, int32_t *reward, uint8_t *done, int32_t *score, int32_t *lives) {
  size_t ss = state_size(game);
  for (int i = 0; i < n; i++) if (tbo_ale_action_to_input(ale_actions[i]) < 0) return -1;
  #pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    void *st = (char *)states + ss * i;
    int32_t sc, lv;
    step1(game, cfg, st, tbo_ale_action_to_input(ale_actions[i]));
    score_lives(game, st, &sc, &lv);
    int32_t r = sc - prev_score[i];
    reward[i] = r > 0 ? r : 0; done[i] = lv <= 0; score[i] = sc; lives[i] = lv;
    prev_score[i] = sc;
    if (done[i] && auto_reset) { new_game1(game, cfg, st, &sim[i]); score_lives(game, st, &sc, &lv); prev_score[i] = sc; }
  }
  return 0;
}

void tbo_frame_convert(int game, const uint8_t *rgba, int mode, uint8_t *out, int out_w, int out_h) {
  int w, h; dims(game, &w, &h);
  if (mode == M_RGBA) memcpy(out, rgba, (size_t)w * h * 4);
  else if (mode == M_RGB) tbo_rgba_to_rgb(rgba, w * h, out);
  else if (mode == M_GRAY) tbo_rgba_to_gray(rgba, w * h, out);
  else {
    uint8_t *g = (uint8_t *)malloc((size_t)w * h);
    tbo_rgba_to_gray(rgba, w * h, g);
    tbo_resize_area_u8(g, w, h, 1, out, out_w, out_h);
    free(g);
  }
}
size_t tbo_frame_bytes(int game, int mode, int out_w, int out_h) {
  int w, h; dims(game, &w, &h);
  return mode == M_RGBA ? (size_t)w * h * 4 : mode == M_RGB ? (size_t)w * h * 3 : mode == M_GRAY ? (size_t)w * h : (size_t)out_w * out_h;
}
void tbo_batch_render(int game, const void *cfg, const void *states, int n, int mode, uint8_t *out, int out_w, int out_h) {
  size_t ss = state_size(game), fb = tbo_frame_bytes(game, mode, out_w, out_h);
  int w, h; dims(game, &w, &h);
  #pragma omp parallel
  {
    uint8_t *rgba = (uint8_t *)malloc((size_t)w * h * 4);
    #pragma omp for schedule(static)
    for (int i = 0; i < n; i++) {
      render1(game, cfg, (const char *)states + ss * i, rgba);
      tbo_frame_convert(game, rgba, mode, out + fb * i, out_w, out_h);
    }
    free(rgba);
  }
}

/* cpu_baseline leg of bench.py: `steps` frames of n envs with the shared synthetic action stream,
 * auto-reset, optional render each frame; returns seconds of wall time. */
double tbo_batch_rollout(int game, const void *cfg, void *states, tbo_rng *sim, int n, int steps, uint64_t t0,
                         uint64_t action_seed, uint64_t env0, const int32_t *legal, int n_legal,
                         int render_mode, uint8_t *frames, int out_w, int out_h, int32_t *prev_score, int64_t *episodes) {
  size_t ss = state_size(game), fb = render_mode >= 0 ? tbo_frame_bytes(game, render_mode, out_w, out_h) : 0;
  int w, h; dims(game, &w, &h);
  struct timespec a, b;
  int64_t eps = 0;
  clock_gettime(CLOCK_MONOTONIC, &a);
  #pragma omp parallel reduction(+:eps)
  {
    uint8_t *rgba = (uint8_t *)malloc((size_t)w * h * 4);
    #pragma omp for schedule(static)
    for (int i = 0; i < n; i++) {
      void *st = (char *)states + ss * i;
      for (int t = 0; t < steps; t++) {
        int32_t sc, lv;
        int act = legal[tbo_action_index(action_seed, env0 + (uint64_t)i, t0 + (uint64_t)t, (uint32_t)n_legal)];
        step1(game, cfg, st, tbo_ale_action_to_input(act));
        score_lives(game, st, &sc, &lv);
        prev_score[i] = sc;
        if (lv <= 0) { new_game1(game, cfg, st, &sim[i]); score_lives(game, st, &sc, &lv); prev_score[i] = sc; eps++; }
        if (render_mode >= 0) { render1(game, cfg, st, rgba); tbo_frame_convert(game, rgba, render_mode, frames + fb * i, out_w, out_h); }
      }
    }
    free(rgba);
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  if (episodes) *episodes = eps;
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
