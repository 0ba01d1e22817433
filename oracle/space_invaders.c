/* space_invaders.c -- CPU ORACLE (test infrastructure, not product code).
 *
 * Restates the Space Invaders engine behind Toybox('space_invaders') (reference call sites
 * toybox/envs/atari/base.py:126,109; state schema toybox/interventions/space_invaders.py:16-32
 * SpaceInvaders, :38 Player, :60 Laser, :101 Ufo, :116 Enemy, :146 EnemiesMovementState; constants and
 * initial state from toybox/interventions/defaults/space_invaders_*_default.json).
 *
 * Pinned by the fixtures ([FIX]): new_game() state incl. RNG lineage (seed 17), ship, 36 enemies,
 * shields bitmap/colour/positions, ufo, timers.  Frame rules S0..S12: PARITY UNPINNED restatement.
 * All integer arithmetic is i32.
 */
#include "tbo.h"
#include <string.h>

#define SHIP_START_X 68
#define SHIP_MIN_X 38
#define SHIP_MAX_X 266
#define ENEMY_W 16
#define ENEMY_H 10
#define FORM_MIN_X 22
#define FORM_MAX_X 298
#define FORM_DX 4
#define FORM_DY 10
#define GROUND_Y 195
#define UFO_W 16
#define UFO_H 7
#define UFO_POINTS 100
#define UFO_PERIOD 500
#define UFO_START_X (-2)
#define UFO_SPEED 2
#define SHOT_DELAY 50
#define LIFE_DISPLAY 128
#define SHIP_DEATH_TIME 30
#define ENEMY_DEATH_TIME 8
#define UFO_DEATH_TIME 16
#define LASER_W 2
#define LASER_H 8
#define SHIP_LASER_SPEED 8
#define ENEMY_LASER_SPEED 3
#define MAX_ACTIVE_ENEMY_LASERS 3

const uint16_t TBO_SI_SHIELD_ROWS[TBO_SI_SHIELD_H] = {
  0x0FF0, 0x0FF0, 0x3FFC, 0x3FFC, 0x3FFC, 0x3FFC, 0x3FFC, 0x3FFC, 0x3FFC, 0x3FFC,
  0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF, 0xF00F, 0xF00F };
const tbo_color TBO_SI_SHIELD_COLOR = { 172, 80, 48, 255 };
static const tbo_color ENEMY_LASER_COLOR = { 252, 252, 84, 255 };
static const tbo_color ENEMY_COLOR = { 134, 134, 29, 255 };
static const tbo_color UFO_COLOR = { 151, 25, 122, 255 };
static const tbo_color GROUND_COLOR = { 80, 89, 22, 255 };
static const tbo_color HUD_COLOR = { 50, 132, 50, 255 };
static const tbo_color BLACK = { 0, 0, 0, 255 };

static const uint32_t SPR_INVADER[3][2][ENEMY_H] = {
  { {0x0810, 0x0420, 0x0FF0, 0x1BD8, 0x3FFC, 0x2FF4, 0x2814, 0x0660, 0, 0},
    {0x0810, 0x2424, 0x2FF4, 0x3BDC, 0x3FFC, 0x1FF8, 0x0810, 0x1008, 0, 0} },
  { {0x0180, 0x03C0, 0x07E0, 0x0DB0, 0x0FF0, 0x0240, 0x05A0, 0x0A50, 0, 0},
    {0x0180, 0x03C0, 0x07E0, 0x0DB0, 0x0FF0, 0x05A0, 0x0810, 0x0420, 0, 0} },
  { {0x03C0, 0x1FF8, 0x3FFC, 0x39CC, 0x3FFC, 0x0660, 0x0DB0, 0x300C, 0, 0},
    {0x03C0, 0x1FF8, 0x3FFC, 0x39CC, 0x3FFC, 0x0E70, 0x1998, 0x0C30, 0, 0} } };
static const uint32_t SPR_SHIP[ENEMY_H] = { 0x0100, 0x0380, 0x0380, 0x3FF8, 0x7FFC, 0x7FFC, 0x7FFC, 0x7FFC, 0x7FFC, 0x7FFC };
static const uint32_t SPR_UFO[UFO_H] = { 0x07E0, 0x1FF8, 0x3FFC, 0x6DB6, 0xFFFF, 0x399C, 0x1008 };
static const uint32_t SPR_BOOM[2][ENEMY_H] = {
  {0x0890, 0x4512, 0x2244, 0x1008, 0xC003, 0x1008, 0x2244, 0x4512, 0x0890, 0x0000},
  {0x1248, 0x0420, 0x4812, 0x2004, 0x0240, 0x9009, 0x0420, 0x2814, 0x4002, 0x1248} };

void tbo_si_default_cfg(tbo_si_cfg *c) {
  static const int32_t rs[6] = { 30, 30, 20, 20, 10, 10 };
  memset(c, 0, sizeof *c);
  c->jitter = 0.5; c->enemy_protocol = TBO_SI_PROTO_TARGET_PLAYER; c->start_lives = 3;
  c->shields[0][0] = 84; c->shields[1][0] = 148; c->shields[2][0] = 212;
  for (int i = 0; i < 3; i++) c->shields[i][1] = 157;
  for (int i = 0; i < 6; i++) c->row_scores[i] = rs[i];
  tbo_rng_seed(&c->rand, 17);
}

static void reset_enemies(const tbo_si_cfg *c, tbo_si_state *s) {
  for (int row = 0; row < 6; row++)
    for (int col = 0; col < 6; col++) {
      tbo_si_enemy *e = &s->enemies[row * 6 + col];         /* [FIX] row-major, id = row*6+col */
      e->x = 44 + 32 * col; e->y = 31 + 18 * row; e->row = row; e->col = col; e->id = row * 6 + col;
      e->alive = 1; e->points = c->row_scores[row]; e->death_counter = TBO_NONE;
    }
  s->enemies_movement.move_counter = 32; s->enemies_movement.move_dir = TBO_DIR_RIGHT;
  s->enemies_movement.visual_orientation = 1;
}
static void reset_shields(const tbo_si_cfg *c, tbo_si_state *s) {
  for (int i = 0; i < TBO_SI_N_SHIELDS; i++) {
    s->shields[i].x = c->shields[i][0]; s->shields[i].y = c->shields[i][1];
    memcpy(s->shields[i].rows, TBO_SI_SHIELD_ROWS, sizeof TBO_SI_SHIELD_ROWS);
  }
}

void tbo_si_new_game(tbo_si_cfg *c, tbo_si_state *s) {
  memset(s, 0, sizeof *s);
  s->rand = tbo_rng_child(&c->rand);
  s->ship.x = SHIP_START_X; s->ship.y = 185; s->ship.w = 16; s->ship.h = 10; s->ship.speed = 3;
  s->ship.alive = 0; s->ship.death_hit_1 = 1; s->ship.death_counter = TBO_NONE;
  s->ship.color = (tbo_color){35, 129, 59, 255};
  s->has_ship_laser = 0;
  reset_enemies(c, s); reset_shields(c, s);
  s->n_enemy_lasers = 0;
  s->ufo.x = UFO_START_X; s->ufo.y = 12; s->ufo.appearance_counter = UFO_PERIOD; s->ufo.death_counter = TBO_NONE;
  s->life_display_timer = LIFE_DISPLAY; s->enemy_shot_delay = SHOT_DELAY;
  s->score = 0; s->lives = c->start_lives; s->level = 1;
}

static int overlap(int ax, int ay, int aw, int ah, int bx, int by, int bw, int bh) {
  return ax < bx + bw && bx < ax + aw && ay < by + bh && by < ay + ah;
}
/* laser vs shields: any opaque pixel inside the laser box => erase the box grown by 1 px, report hit */
static int hit_shields(tbo_si_state *s, int lx, int ly, int lw, int lh) {
  for (int i = 0; i < TBO_SI_N_SHIELDS; i++) {
    int sx = s->shields[i].x, sy = s->shields[i].y, hit = 0;
    if (!overlap(lx, ly, lw, lh, sx, sy, TBO_SI_SHIELD_W, TBO_SI_SHIELD_H)) continue;
    for (int r = 0; r < TBO_SI_SHIELD_H && !hit; r++)
      for (int q = 0; q < TBO_SI_SHIELD_W; q++) {
        int px = sx + q, py = sy + r;
        if (px >= lx && px < lx + lw && py >= ly && py < ly + lh && ((s->shields[i].rows[r] >> (15 - q)) & 1)) { hit = 1; break; }
      }
    if (!hit) continue;
    for (int r = 0; r < TBO_SI_SHIELD_H; r++)
      for (int q = 0; q < TBO_SI_SHIELD_W; q++) {
        int px = sx + q, py = sy + r;
        if (px >= lx - 1 && px < lx + lw + 1 && py >= ly - 1 && py < ly + lh + 1)
          s->shields[i].rows[r] &= (uint16_t)~(1u << (15 - q));
      }
    return 1;
  }
  return 0;
}
static void move_laser(tbo_si_laser *l) {
  switch (l->movement) {
    case TBO_DIR_UP: l->y -= l->speed; break;
    case TBO_DIR_DOWN: l->y += l->speed; break;
    case TBO_DIR_LEFT: l->x -= l->speed; break;
    default: l->x += l->speed; break;
  }
  l->t += 1;
}
static int offscreen(const tbo_si_laser *l) {
  return l->y + l->h <= 0 || l->y >= GROUND_Y || l->x + l->w <= 0 || l->x >= TBO_SI_W;
}

void tbo_si_step(const tbo_si_cfg *c, tbo_si_state *s, int in) {
  /* S0 */
  if (s->lives <= 0) return;
  /* S1 lives display: the world is frozen */
  if (s->life_display_timer > 0) {
    s->life_display_timer -= 1;
    if (s->life_display_timer == 0) s->ship.alive = 1;
    return;
  }
  /* S2 ship death animation: the world is frozen */
  if (s->ship.death_counter != TBO_NONE) {
    s->ship.death_counter -= 1;
    if ((s->ship.death_counter & 3) == 0) s->ship.death_hit_1 = !s->ship.death_hit_1;
    if (s->ship.death_counter <= 0) {
      s->ship.death_counter = TBO_NONE; s->ship.death_hit_1 = 1;
      s->lives -= 1; s->ship.x = SHIP_START_X;
      s->n_enemy_lasers = 0; s->has_ship_laser = 0;
      if (s->lives > 0) s->life_display_timer = LIFE_DISPLAY;
    }
    return;
  }
  /* S3 ship motion */
  if (s->ship.alive) {
    int left = (in & TBO_IN_LEFT) != 0, right = (in & TBO_IN_RIGHT) != 0;
    if (left && !right) s->ship.x -= s->ship.speed;
    if (right && !left) s->ship.x += s->ship.speed;
    if (s->ship.x < SHIP_MIN_X) s->ship.x = SHIP_MIN_X;
    if (s->ship.x > SHIP_MAX_X) s->ship.x = SHIP_MAX_X;
  }
  /* S4 fire */
  if (s->ship.alive && (in & TBO_IN_BUTTON1) && !s->has_ship_laser) {
    tbo_si_laser *l = &s->ship_laser;
    l->x = s->ship.x + s->ship.w / 2 - 1; l->y = s->ship.y - LASER_H; l->w = LASER_W; l->h = LASER_H;
    l->t = 0; l->movement = TBO_DIR_UP; l->speed = SHIP_LASER_SPEED; l->color = s->ship.color;
    s->has_ship_laser = 1;
  }
  /* S5 ship laser */
  if (s->has_ship_laser) {
    tbo_si_laser *l = &s->ship_laser;
    move_laser(l);
    if (offscreen(l)) s->has_ship_laser = 0;
    else if (hit_shields(s, l->x, l->y, l->w, l->h)) s->has_ship_laser = 0;
    else {
      for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
        tbo_si_enemy *e = &s->enemies[i];
        if (e->alive && overlap(l->x, l->y, l->w, l->h, e->x, e->y, ENEMY_W, ENEMY_H)) {
          e->alive = 0; e->death_counter = ENEMY_DEATH_TIME; s->score += e->points; s->has_ship_laser = 0;
          break;
        }
      }
      if (s->has_ship_laser && s->ufo.appearance_counter == TBO_NONE && s->ufo.death_counter == TBO_NONE &&
          overlap(l->x, l->y, l->w, l->h, s->ufo.x, s->ufo.y, UFO_W, UFO_H)) {
        s->score += UFO_POINTS; s->ufo.death_counter = UFO_DEATH_TIME; s->has_ship_laser = 0;
      }
    }
  }
  /* S6 enemy explosions */
  for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
    tbo_si_enemy *e = &s->enemies[i];
    if (e->death_counter != TBO_NONE) { e->death_counter -= 1; if (e->death_counter <= 0) e->death_counter = TBO_NONE; }
  }
  /* S7 mothership */
  if (s->ufo.death_counter != TBO_NONE) {
    s->ufo.death_counter -= 1;
    if (s->ufo.death_counter <= 0) { s->ufo.death_counter = TBO_NONE; s->ufo.x = UFO_START_X; s->ufo.appearance_counter = UFO_PERIOD; }
  } else if (s->ufo.appearance_counter == TBO_NONE) {
    s->ufo.x += UFO_SPEED;
    if (s->ufo.x >= TBO_SI_W) { s->ufo.x = UFO_START_X; s->ufo.appearance_counter = UFO_PERIOD; }
  } else if (s->ufo.appearance_counter >= 0) {
    if (s->ufo.appearance_counter > 0) s->ufo.appearance_counter -= 1;
    if (s->ufo.appearance_counter == 0) s->ufo.appearance_counter = TBO_NONE;
  }
  /* S8 formation */
  int n_alive = 0;
  for (int i = 0; i < TBO_SI_N_ENEMIES; i++) n_alive += s->enemies[i].alive;
  s->enemies_movement.move_counter -= 1;
  if (s->enemies_movement.move_counter <= 0) {
    int dx = s->enemies_movement.move_dir == TBO_DIR_RIGHT ? FORM_DX : -FORM_DX, edge = 0;
    for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
      tbo_si_enemy *e = &s->enemies[i];
      if (e->alive && (e->x + dx < FORM_MIN_X || e->x + ENEMY_W + dx > FORM_MAX_X)) edge = 1;
    }
    for (int i = 0; i < TBO_SI_N_ENEMIES; i++) { if (edge) s->enemies[i].y += FORM_DY; else s->enemies[i].x += dx; }
    if (edge) s->enemies_movement.move_dir = s->enemies_movement.move_dir == TBO_DIR_RIGHT ? TBO_DIR_LEFT : TBO_DIR_RIGHT;
    s->enemies_movement.visual_orientation = !s->enemies_movement.visual_orientation;
    s->enemies_movement.move_counter = 2 + (30 * n_alive) / 36;
    /* S9 invaders eat shields and land */
    for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
      tbo_si_enemy *e = &s->enemies[i];
      if (!e->alive) continue;
      if (e->y + ENEMY_H >= s->ship.y) s->lives = 0;
      for (int k = 0; k < TBO_SI_N_SHIELDS; k++) {
        int sx = s->shields[k].x, sy = s->shields[k].y;
        if (!overlap(e->x, e->y, ENEMY_W, ENEMY_H, sx, sy, TBO_SI_SHIELD_W, TBO_SI_SHIELD_H)) continue;
        for (int r = 0; r < TBO_SI_SHIELD_H; r++)
          for (int q = 0; q < TBO_SI_SHIELD_W; q++) {
            int px = sx + q, py = sy + r;
            if (px >= e->x && px < e->x + ENEMY_W && py >= e->y && py < e->y + ENEMY_H)
              s->shields[k].rows[r] &= (uint16_t)~(1u << (15 - q));
          }
      }
    }
    if (s->lives <= 0) return;
  }
  /* S10 enemy fire */
  s->enemy_shot_delay -= 1;
  if (s->enemy_shot_delay <= 0) {
    s->enemy_shot_delay = SHOT_DELAY;
    if (n_alive > 0 && s->n_enemy_lasers < MAX_ACTIVE_ENEMY_LASERS) {
      int shooter[6], cols[6], nc = 0;
      for (int col = 0; col < 6; col++) shooter[col] = -1;
      for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
        tbo_si_enemy *e = &s->enemies[i];
        if (!e->alive || e->col < 0 || e->col > 5) continue;
        if (shooter[e->col] < 0 || e->row > s->enemies[shooter[e->col]].row) shooter[e->col] = i;
      }
      for (int col = 0; col < 6; col++) if (shooter[col] >= 0) cols[nc++] = col;
      if (nc > 0) {
        int pick = -1;
        if (c->enemy_protocol == TBO_SI_PROTO_TARGET_PLAYER) {
          double r = tbo_rng_f64(&s->rand);
          if (!(r < c->jitter)) {
            int best = 0x7fffffff, target = s->ship.x + s->ship.w / 2;
            for (int k = 0; k < nc; k++) {
              int d = s->enemies[shooter[cols[k]]].x + ENEMY_W / 2 - target;
              if (d < 0) d = -d;
              if (d < best) { best = d; pick = cols[k]; }
            }
          }
        }
        if (pick < 0) pick = cols[tbo_rng_index(&s->rand, (uint32_t)nc)];
        tbo_si_enemy *e = &s->enemies[shooter[pick]];
        tbo_si_laser *l = &s->enemy_lasers[s->n_enemy_lasers++];
        l->x = e->x + ENEMY_W / 2 - 1; l->y = e->y + ENEMY_H; l->w = LASER_W; l->h = LASER_H; l->t = 0;
        l->movement = TBO_DIR_DOWN; l->speed = ENEMY_LASER_SPEED; l->color = ENEMY_LASER_COLOR;
      }
    }
  }
  /* S11 enemy lasers */
  {
    int n = 0;
    for (int i = 0; i < s->n_enemy_lasers; i++) {
      tbo_si_laser l = s->enemy_lasers[i];
      int keep = 1;
      move_laser(&l);
      if (offscreen(&l)) keep = 0;
      else if (hit_shields(s, l.x, l.y, l.w, l.h)) keep = 0;
      else if (s->ship.alive && overlap(l.x, l.y, l.w, l.h, s->ship.x, s->ship.y, s->ship.w, s->ship.h)) {
        s->ship.alive = 0; s->ship.death_counter = SHIP_DEATH_TIME; s->ship.death_hit_1 = 1; keep = 0;
      }
      if (keep) s->enemy_lasers[n++] = l;
    }
    s->n_enemy_lasers = n;
  }
  /* S12 wave cleared */
  if (n_alive == 0) {
    int busy = 0;
    for (int i = 0; i < TBO_SI_N_ENEMIES; i++) busy |= (s->enemies[i].alive || s->enemies[i].death_counter != TBO_NONE);
    if (!busy) { s->level += 1; reset_enemies(c, s); reset_shields(c, s); s->n_enemy_lasers = 0; }
  }
}

void tbo_si_render(const tbo_si_cfg *c, const tbo_si_state *s, uint8_t *rgba) {
  tbo_canvas cv = { TBO_SI_W, TBO_SI_H, rgba };
  (void)c;
  tbo_clear(&cv, BLACK);
  tbo_rect(&cv, GROUND_COLOR, 0, GROUND_Y, TBO_SI_W, 2);
  tbo_digits(&cv, HUD_COLOR, 120, 2, s->score, 2, 1);
  if (s->life_display_timer > 0) tbo_digits(&cv, HUD_COLOR, 170, 199, s->lives, 3, 2);
  for (int i = 0; i < TBO_SI_N_SHIELDS; i++)
    for (int r = 0; r < TBO_SI_SHIELD_H; r++) {
      uint32_t row = s->shields[i].rows[r];
      tbo_sprite1(&cv, TBO_SI_SHIELD_COLOR, s->shields[i].x, s->shields[i].y + r, 16, 1, &row, 1, 1);
    }
  for (int i = 0; i < TBO_SI_N_ENEMIES; i++) {
    const tbo_si_enemy *e = &s->enemies[i];
    if (e->alive) {
      int kind = e->row < 0 ? 0 : e->row > 5 ? 2 : e->row / 2;
      tbo_sprite1(&cv, ENEMY_COLOR, e->x, e->y, 16, ENEMY_H, SPR_INVADER[kind][s->enemies_movement.visual_orientation ? 1 : 0], 1, 1);
    } else if (e->death_counter != TBO_NONE) {
      tbo_sprite1(&cv, ENEMY_COLOR, e->x, e->y, 16, ENEMY_H, SPR_BOOM[0], 1, 1);
    }
  }
  if (s->ship.alive) tbo_sprite1(&cv, s->ship.color, s->ship.x, s->ship.y, 16, ENEMY_H, SPR_SHIP, 1, 1);
  else if (s->ship.death_counter != TBO_NONE)
    tbo_sprite1(&cv, s->ship.color, s->ship.x, s->ship.y, 16, ENEMY_H, SPR_BOOM[s->ship.death_hit_1 ? 0 : 1], 1, 1);
  if (s->ufo.death_counter != TBO_NONE) tbo_sprite1(&cv, UFO_COLOR, s->ufo.x, s->ufo.y, 16, ENEMY_H, SPR_BOOM[1], 1, 1);
  else if (s->ufo.appearance_counter == TBO_NONE) tbo_sprite1(&cv, UFO_COLOR, s->ufo.x, s->ufo.y, 16, UFO_H, SPR_UFO, 1, 1);
  if (s->has_ship_laser) tbo_rect(&cv, s->ship_laser.color, s->ship_laser.x, s->ship_laser.y, s->ship_laser.w, s->ship_laser.h);
  for (int i = 0; i < s->n_enemy_lasers; i++) {
    const tbo_si_laser *l = &s->enemy_lasers[i];
    tbo_rect(&cv, l->color, l->x, l->y, l->w, l->h);
  }
}
