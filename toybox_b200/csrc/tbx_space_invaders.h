/* tbx_space_invaders.h -- Space Invaders transition, new_game and draw list, one env per call.
 *
 * Replaces what Toybox('space_invaders').apply_ale_action / new_game / get_state reach in ctoybox
 * (reference call sites toybox/envs/atari/base.py:126,153,109; state fields
 * toybox/interventions/space_invaders.py:16-32, :38 Player, :60 Laser, :101 Ufo, :116 Enemy,
 * :146 EnemiesMovementState; constants and the initial state from
 * toybox/interventions/defaults/space_invaders_{config,state}_default.json).  All arithmetic is i32.
 */
#ifndef TBX_SPACE_INVADERS_H
#define TBX_SPACE_INVADERS_H
#include "tbx_common.h"

#define SI_W(f) TBX_W(SiRec, f)
#define SI_SHIP_START_X 68
#define SI_SHIP_MIN_X 38
#define SI_SHIP_MAX_X 266
#define SI_ENEMY_W 16
#define SI_ENEMY_H 10
#define SI_FORM_MIN_X 22
#define SI_FORM_MAX_X 298
#define SI_FORM_DX 4
#define SI_FORM_DY 10
#define SI_GROUND_Y 195
#define SI_UFO_W 16
#define SI_UFO_H 7
#define SI_UFO_POINTS 100
#define SI_UFO_PERIOD 500
#define SI_UFO_START_X (-2)
#define SI_UFO_SPEED 2
#define SI_SHOT_DELAY 50
#define SI_LIFE_DISPLAY 128
#define SI_SHIP_DEATH_TIME 30
#define SI_ENEMY_DEATH_TIME 8
#define SI_UFO_DEATH_TIME 16
#define SI_LASER_W 2
#define SI_LASER_H 8
#define SI_SHIP_LASER_SPEED 8
#define SI_ENEMY_LASER_SPEED 3
#define SI_MAX_ACTIVE_ENEMY_LASERS 3
#define SI_LASER_WORDS 8

#define SI_COLOR_SHIELD 0xFF3050ACu      /* 172, 80, 48 */
#define SI_COLOR_ENEMY_LASER 0xFF54FCFCu /* 252,252, 84 */
#define SI_COLOR_ENEMY 0xFF1D8686u       /* 134,134, 29 */
#define SI_COLOR_UFO 0xFF7A1997u         /* 151, 25,122 */
#define SI_COLOR_GROUND 0xFF165950u      /*  80, 89, 22 */
#define SI_COLOR_HUD 0xFF328432u         /*  50,132, 50 */
#define SI_COLOR_SHIP 0xFF3B8123u        /*  35,129, 59 */
#define SI_COLOR_BLACK 0xFF000000u

/* draw-list slots */
#define SI_SLOT_GROUND 0
#define SI_SLOT_SCORE 1
#define SI_SLOT_LIVES (SI_SLOT_SCORE + TBX_MAX_DIGITS)
#define SI_SLOT_SHIELDS (SI_SLOT_LIVES + TBX_MAX_DIGITS)
#define SI_SLOT_ENEMIES (SI_SLOT_SHIELDS + TBX_SI_N_SHIELDS)
#define SI_SLOT_SHIP (SI_SLOT_ENEMIES + TBX_SI_N_ENEMIES)
#define SI_SLOT_UFO (SI_SLOT_SHIP + 1)
#define SI_SLOT_SHIP_LASER (SI_SLOT_UFO + 1)
#define SI_SLOT_ENEMY_LASERS (SI_SLOT_SHIP_LASER + 1)
#define SI_N_SLOTS (SI_SLOT_ENEMY_LASERS + TBX_SI_MAX_LASERS)
#define SI_N_STATIC 1
#define SI_N_GROUPS 4
/* HUD digits | shields (one colour) | invaders (one colour) | ship, ufo, lasers (in order) */
TBX_HD void si_group(int g, int &b, int &e, int &mode) {
  if (g == 0) { b = SI_SLOT_SCORE; e = SI_SLOT_SHIELDS; mode = TBX_GROUP_PARALLEL; }
  else if (g == 1) { b = SI_SLOT_SHIELDS; e = SI_SLOT_ENEMIES; mode = TBX_GROUP_PARALLEL; }
  else if (g == 2) { b = SI_SLOT_ENEMIES; e = SI_SLOT_SHIP; mode = TBX_GROUP_PARALLEL; }
  else { b = SI_SLOT_SHIP; e = SI_N_SLOTS; mode = TBX_GROUP_SERIAL; }
}

TBX_HD uint32_t si_shield_row_default(int r) { return r < 2 ? 0x0FF0u : r < 10 ? 0x3FFCu : r < 16 ? 0xFFFFu : 0xF00Fu; }

TBX_HD void si_reset_enemies(const TbxAcc &S, const SiCfg &c) {
  for (int row = 0; row < 6; row++)
    for (int col = 0; col < 6; col++) {
      int i = row * 6 + col;
      S.sti(SI_W(en_x) + i, 44 + 32 * col);
      S.sti(SI_W(en_y) + i, 31 + 18 * row);
      S.sti(SI_W(en_row) + i, row);
      S.sti(SI_W(en_col) + i, col);
      S.sti(SI_W(en_id) + i, i);
      S.sti(SI_W(en_points) + i, c.row_scores[row]);
      S.sti(SI_W(en_death) + i, TBX_NONE);
    }
  S.st(SI_W(en_alive), 0xffffffffu);
  S.st(SI_W(en_alive) + 1, 0xfu);
  S.sti(SI_W(move_counter), 32);
  S.sti(SI_W(move_dir), TBX_DIR_RIGHT);
  S.sti(SI_W(visual_orientation), 1);
}
TBX_HD void si_reset_shields(const TbxAcc &S, const SiCfg &c) {
  for (int i = 0; i < TBX_SI_N_SHIELDS; i++) {
    S.sti(SI_W(shield_x) + i, c.shields[i][0]);
    S.sti(SI_W(shield_y) + i, c.shields[i][1]);
    for (int r = 0; r < TBX_SI_SHIELD_H; r++) S.st(SI_W(shield_rows) + i * TBX_SI_SHIELD_H + r, si_shield_row_default(r));
  }
}
TBX_HD void si_store_laser(const TbxAcc &S, int w, int x, int y, int lw, int lh, int t, int movement, int speed, uint32_t color) {
  S.sti(w + 0, x); S.sti(w + 1, y); S.sti(w + 2, lw); S.sti(w + 3, lh); S.sti(w + 4, t);
  S.sti(w + 5, movement); S.sti(w + 6, speed); S.st(w + 7, color);
}

TBX_HD void si_new_game(const TbxAcc &S, const SiCfg &c) {
  TbxRng sim = tbx_rng_load(S, TBX_HW(sim_rand));
  TbxRng rng;
  rng.s0 = tbx_rng_next_u64(sim);
  rng.s1 = tbx_rng_next_u64(sim);
  tbx_rng_store(S, TBX_HW(sim_rand), sim);
  tbx_rng_store(S, TBX_HW(rand), rng);
  S.sti(TBX_HW(lives), c.start_lives);
  S.sti(TBX_HW(score), 0);
  S.sti(TBX_HW(level), 1);
  S.sti(TBX_HW(prev_score), 0);
  S.sti(TBX_HW(ep_len), 0);
  S.sti(TBX_HW(ep_return), 0);
  S.sti(TBX_HW(tbl), 0);
  S.sti(SI_W(ship_x), SI_SHIP_START_X); S.sti(SI_W(ship_y), 185); S.sti(SI_W(ship_w), 16); S.sti(SI_W(ship_h), 10);
  S.sti(SI_W(ship_speed), 3); S.sti(SI_W(ship_alive), 0); S.sti(SI_W(ship_death_hit_1), 1);
  S.sti(SI_W(ship_death_counter), TBX_NONE);
  S.st(SI_W(ship_color), SI_COLOR_SHIP);
  S.sti(SI_W(has_ship_laser), 0);
  si_store_laser(S, SI_W(ship_laser), 0, 0, 0, 0, 0, 0, 0, 0);
  S.sti(SI_W(n_enemy_lasers), 0);
  for (int i = 0; i < TBX_SI_MAX_LASERS; i++) si_store_laser(S, SI_W(enemy_lasers) + SI_LASER_WORDS * i, 0, 0, 0, 0, 0, 0, 0, 0);
  si_reset_enemies(S, c);
  si_reset_shields(S, c);
  S.sti(SI_W(ufo_x), SI_UFO_START_X); S.sti(SI_W(ufo_y), 12);
  S.sti(SI_W(ufo_appearance_counter), SI_UFO_PERIOD); S.sti(SI_W(ufo_death_counter), TBX_NONE);
  S.sti(SI_W(life_display_timer), SI_LIFE_DISPLAY);
  S.sti(SI_W(enemy_shot_delay), SI_SHOT_DELAY);
}

TBX_HD bool si_overlap(int ax, int ay, int aw, int ah, int bx, int by, int bw, int bh) {
  return ax < bx + bw && bx < ax + aw && ay < by + bh && by < ay + ah;
}
/* bits of a 16-wide row covered by pixel columns [x0, x1) given the row's left edge sx */
TBX_HD uint32_t si_row_span(int sx, int x0, int x1) {
  int a = x0 - sx, b = x1 - sx; /* columns [a,b) */
  if (a < 0) a = 0;
  if (b > 16) b = 16;
  if (a >= b) return 0;
  /* column q is bit (15-q) */
  return ((0xFFFFu >> a) & (0xFFFFu << (16 - b))) & 0xFFFFu;
}
/* laser vs shields: any opaque pixel inside the laser box => erase the box grown by 1 px, report hit */
TBX_HD bool si_hit_shields(const TbxAcc &S, int lx, int ly, int lw, int lh) {
  for (int i = 0; i < TBX_SI_N_SHIELDS; i++) {
    int sx = S.ldi(SI_W(shield_x) + i), sy = S.ldi(SI_W(shield_y) + i);
    if (!si_overlap(lx, ly, lw, lh, sx, sy, TBX_SI_SHIELD_W, TBX_SI_SHIELD_H)) continue;
    int base = SI_W(shield_rows) + i * TBX_SI_SHIELD_H;
    uint32_t span = si_row_span(sx, lx, lx + lw);
    bool hit = false;
    for (int r = 0; r < TBX_SI_SHIELD_H && !hit; r++) {
      int py = sy + r;
      if (py >= ly && py < ly + lh && (S.ld(base + r) & span)) hit = true;
    }
    if (!hit) continue;
    uint32_t grown = si_row_span(sx, lx - 1, lx + lw + 1);
    for (int r = 0; r < TBX_SI_SHIELD_H; r++) {
      int py = sy + r;
      if (py >= ly - 1 && py < ly + lh + 1) S.st(base + r, S.ld(base + r) & ~grown);
    }
    return true;
  }
  return false;
}
struct SiL { int x, y, w, h, t, movement, speed; uint32_t color; };
TBX_HD SiL si_load_laser(const TbxAcc &S, int w) {
  SiL l; l.x = S.ldi(w); l.y = S.ldi(w + 1); l.w = S.ldi(w + 2); l.h = S.ldi(w + 3); l.t = S.ldi(w + 4);
  l.movement = S.ldi(w + 5); l.speed = S.ldi(w + 6); l.color = S.ld(w + 7);
  return l;
}
TBX_HD void si_move_laser(SiL &l) {
  if (l.movement == TBX_DIR_UP) l.y -= l.speed;
  else if (l.movement == TBX_DIR_DOWN) l.y += l.speed;
  else if (l.movement == TBX_DIR_LEFT) l.x -= l.speed;
  else l.x += l.speed;
  l.t += 1;
}
TBX_HD bool si_offscreen(const SiL &l) { return l.y + l.h <= 0 || l.y >= SI_GROUND_Y || l.x + l.w <= 0 || l.x >= TBX_SI_W; }
TBX_HD bool si_en_alive(const TbxAcc &S, int i) { return (S.ld(SI_W(en_alive) + (i >> 5)) >> (i & 31)) & 1u; }

TBX_HD void si_step(const TbxAcc &S, const SiCfg &c, int in) {
  int lives = S.ldi(TBX_HW(lives));
  if (lives <= 0) return;
  /* lives display: the world is frozen */
  int ldt = S.ldi(SI_W(life_display_timer));
  if (ldt > 0) {
    ldt -= 1;
    S.sti(SI_W(life_display_timer), ldt);
    if (ldt == 0) S.sti(SI_W(ship_alive), 1);
    return;
  }
  /* ship death animation: the world is frozen */
  int sdc = S.ldi(SI_W(ship_death_counter));
  if (sdc != TBX_NONE) {
    sdc -= 1;
    if ((sdc & 3) == 0) S.sti(SI_W(ship_death_hit_1), !S.ldi(SI_W(ship_death_hit_1)));
    if (sdc <= 0) {
      sdc = TBX_NONE;
      S.sti(SI_W(ship_death_hit_1), 1);
      lives -= 1;
      S.sti(TBX_HW(lives), lives);
      S.sti(SI_W(ship_x), SI_SHIP_START_X);
      S.sti(SI_W(n_enemy_lasers), 0);
      S.sti(SI_W(has_ship_laser), 0);
      if (lives > 0) S.sti(SI_W(life_display_timer), SI_LIFE_DISPLAY);
    }
    S.sti(SI_W(ship_death_counter), sdc);
    return;
  }
  int ship_alive = S.ldi(SI_W(ship_alive)), ship_x = S.ldi(SI_W(ship_x)), ship_y = S.ldi(SI_W(ship_y));
  int ship_w = S.ldi(SI_W(ship_w)), ship_h = S.ldi(SI_W(ship_h));
  int score = S.ldi(TBX_HW(score));
  /* ship motion */
  if (ship_alive) {
    bool left = (in & TBX_IN_LEFT) != 0, right = (in & TBX_IN_RIGHT) != 0;
    int sp = S.ldi(SI_W(ship_speed));
    if (left && !right) ship_x -= sp;
    if (right && !left) ship_x += sp;
    if (ship_x < SI_SHIP_MIN_X) ship_x = SI_SHIP_MIN_X;
    if (ship_x > SI_SHIP_MAX_X) ship_x = SI_SHIP_MAX_X;
    S.sti(SI_W(ship_x), ship_x);
  }
  int has_laser = S.ldi(SI_W(has_ship_laser));
  uint32_t alive0 = S.ld(SI_W(en_alive)), alive1 = S.ld(SI_W(en_alive) + 1) & 0xfu;
  /* fire */
  if (ship_alive && (in & TBX_IN_BUTTON1) && !has_laser) {
    si_store_laser(S, SI_W(ship_laser), ship_x + ship_w / 2 - 1, ship_y - SI_LASER_H, SI_LASER_W, SI_LASER_H, 0, TBX_DIR_UP,
                   SI_SHIP_LASER_SPEED, S.ld(SI_W(ship_color)));
    has_laser = 1;
  }
  /* ship laser */
  if (has_laser) {
    SiL l = si_load_laser(S, SI_W(ship_laser));
    si_move_laser(l);
    S.sti(SI_W(ship_laser) + 0, l.x); S.sti(SI_W(ship_laser) + 1, l.y); S.sti(SI_W(ship_laser) + 4, l.t);
    if (si_offscreen(l)) has_laser = 0;
    else if (si_hit_shields(S, l.x, l.y, l.w, l.h)) has_laser = 0;
    else {
      for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
        bool a = ((i < 32 ? alive0 >> i : alive1 >> (i - 32)) & 1u) != 0;
        if (a && si_overlap(l.x, l.y, l.w, l.h, S.ldi(SI_W(en_x) + i), S.ldi(SI_W(en_y) + i), SI_ENEMY_W, SI_ENEMY_H)) {
          if (i < 32) alive0 &= ~(1u << i); else alive1 &= ~(1u << (i - 32));
          S.sti(SI_W(en_death) + i, SI_ENEMY_DEATH_TIME);
          score += S.ldi(SI_W(en_points) + i);
          has_laser = 0;
          break;
        }
      }
      if (has_laser && S.ldi(SI_W(ufo_appearance_counter)) == TBX_NONE && S.ldi(SI_W(ufo_death_counter)) == TBX_NONE &&
          si_overlap(l.x, l.y, l.w, l.h, S.ldi(SI_W(ufo_x)), S.ldi(SI_W(ufo_y)), SI_UFO_W, SI_UFO_H)) {
        score += SI_UFO_POINTS;
        S.sti(SI_W(ufo_death_counter), SI_UFO_DEATH_TIME);
        has_laser = 0;
      }
    }
  }
  S.sti(SI_W(has_ship_laser), has_laser);
  S.st(SI_W(en_alive), alive0);
  S.st(SI_W(en_alive) + 1, alive1);
  S.sti(TBX_HW(score), score);
  /* enemy explosions */
  for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
    int d = S.ldi(SI_W(en_death) + i);
    if (d != TBX_NONE) { d -= 1; if (d <= 0) d = TBX_NONE; S.sti(SI_W(en_death) + i, d); }
  }
  /* mothership */
  {
    int udc = S.ldi(SI_W(ufo_death_counter)), uac = S.ldi(SI_W(ufo_appearance_counter));
    if (udc != TBX_NONE) {
      udc -= 1;
      if (udc <= 0) { udc = TBX_NONE; S.sti(SI_W(ufo_x), SI_UFO_START_X); S.sti(SI_W(ufo_appearance_counter), SI_UFO_PERIOD); }
      S.sti(SI_W(ufo_death_counter), udc);
    } else if (uac == TBX_NONE) {
      int ux = S.ldi(SI_W(ufo_x)) + SI_UFO_SPEED;
      if (ux >= TBX_SI_W) { ux = SI_UFO_START_X; S.sti(SI_W(ufo_appearance_counter), SI_UFO_PERIOD); }
      S.sti(SI_W(ufo_x), ux);
    } else if (uac >= 0) {
      if (uac > 0) uac -= 1;
      if (uac == 0) uac = TBX_NONE;
      S.sti(SI_W(ufo_appearance_counter), uac);
    }
  }
  /* formation */
  int n_alive = tbx_popc(alive0) + tbx_popc(alive1);
  int mc = S.ldi(SI_W(move_counter)) - 1;
  if (mc <= 0) {
    int dir = S.ldi(SI_W(move_dir));
    int dx = dir == TBX_DIR_RIGHT ? SI_FORM_DX : -SI_FORM_DX;
    bool edge = false;
    for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
      bool a = ((i < 32 ? alive0 >> i : alive1 >> (i - 32)) & 1u) != 0;
      if (a) { int ex = S.ldi(SI_W(en_x) + i); if (ex + dx < SI_FORM_MIN_X || ex + SI_ENEMY_W + dx > SI_FORM_MAX_X) edge = true; }
    }
    for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
      if (edge) S.sti(SI_W(en_y) + i, S.ldi(SI_W(en_y) + i) + SI_FORM_DY);
      else S.sti(SI_W(en_x) + i, S.ldi(SI_W(en_x) + i) + dx);
    }
    if (edge) S.sti(SI_W(move_dir), dir == TBX_DIR_RIGHT ? TBX_DIR_LEFT : TBX_DIR_RIGHT);
    S.sti(SI_W(visual_orientation), !S.ldi(SI_W(visual_orientation)));
    mc = 2 + (30 * n_alive) / 36;
    /* invaders eat shields and land */
    for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
      bool a = ((i < 32 ? alive0 >> i : alive1 >> (i - 32)) & 1u) != 0;
      if (!a) continue;
      int ex = S.ldi(SI_W(en_x) + i), ey = S.ldi(SI_W(en_y) + i);
      if (ey + SI_ENEMY_H >= ship_y) lives = 0;
      for (int k = 0; k < TBX_SI_N_SHIELDS; k++) {
        int sx = S.ldi(SI_W(shield_x) + k), sy = S.ldi(SI_W(shield_y) + k);
        if (!si_overlap(ex, ey, SI_ENEMY_W, SI_ENEMY_H, sx, sy, TBX_SI_SHIELD_W, TBX_SI_SHIELD_H)) continue;
        uint32_t span = si_row_span(sx, ex, ex + SI_ENEMY_W);
        for (int r = 0; r < TBX_SI_SHIELD_H; r++) {
          int py = sy + r;
          if (py >= ey && py < ey + SI_ENEMY_H) {
            int w = SI_W(shield_rows) + k * TBX_SI_SHIELD_H + r;
            S.st(w, S.ld(w) & ~span);
          }
        }
      }
    }
    S.sti(SI_W(move_counter), mc);
    if (lives <= 0) { S.sti(TBX_HW(lives), 0); return; }
  } else {
    S.sti(SI_W(move_counter), mc);
  }
  /* enemy fire */
  int n_lasers = S.ldi(SI_W(n_enemy_lasers));
  int delay = S.ldi(SI_W(enemy_shot_delay)) - 1;
  if (delay <= 0) {
    delay = SI_SHOT_DELAY;
    if (n_alive > 0 && n_lasers < SI_MAX_ACTIVE_ENEMY_LASERS) {
      int shooter[6], srow[6];
      for (int col = 0; col < 6; col++) { shooter[col] = -1; srow[col] = 0; }
      for (int i = 0; i < TBX_SI_N_ENEMIES; i++) {
        bool a = ((i < 32 ? alive0 >> i : alive1 >> (i - 32)) & 1u) != 0;
        if (!a) continue;
        int col = S.ldi(SI_W(en_col) + i);
        if (col < 0 || col > 5) continue;
        int row = S.ldi(SI_W(en_row) + i);
        for (int q = 0; q < 6; q++)
          if (q == col && (shooter[q] < 0 || row > srow[q])) { shooter[q] = i; srow[q] = row; }
      }
      int nc = 0;
      for (int col = 0; col < 6; col++) nc += shooter[col] >= 0;
      if (nc > 0) {
        TbxRng rng = tbx_rng_load(S, TBX_HW(rand));
        int pick = -1;
        if (c.enemy_protocol == 0) {
          double rr = tbx_rng_f64(rng);
          if (!(rr < c.jitter)) {
            int best = 0x7fffffff, target = ship_x + ship_w / 2;
            for (int col = 0; col < 6; col++) {
              if (shooter[col] < 0) continue;
              int d = S.ldi(SI_W(en_x) + shooter[col]) + SI_ENEMY_W / 2 - target;
              if (d < 0) d = -d;
              if (d < best) { best = d; pick = shooter[col]; }
            }
          }
        }
        if (pick < 0) {
          int k = (int)tbx_rng_index(rng, (uint32_t)nc);
          for (int col = 0; col < 6; col++)
            if (shooter[col] >= 0) { if (k == 0) pick = shooter[col]; k--; }
        }
        tbx_rng_store(S, TBX_HW(rand), rng);
        si_store_laser(S, SI_W(enemy_lasers) + SI_LASER_WORDS * n_lasers, S.ldi(SI_W(en_x) + pick) + SI_ENEMY_W / 2 - 1,
                       S.ldi(SI_W(en_y) + pick) + SI_ENEMY_H, SI_LASER_W, SI_LASER_H, 0, TBX_DIR_DOWN, SI_ENEMY_LASER_SPEED,
                       SI_COLOR_ENEMY_LASER);
        n_lasers++;
      }
    }
  }
  S.sti(SI_W(enemy_shot_delay), delay);
  /* enemy lasers */
  {
    int n = 0;
    for (int i = 0; i < n_lasers && i < TBX_SI_MAX_LASERS; i++) {
      SiL l = si_load_laser(S, SI_W(enemy_lasers) + SI_LASER_WORDS * i);
      bool keep = true;
      si_move_laser(l);
      if (si_offscreen(l)) keep = false;
      else if (si_hit_shields(S, l.x, l.y, l.w, l.h)) keep = false;
      else if (ship_alive && si_overlap(l.x, l.y, l.w, l.h, ship_x, ship_y, ship_w, ship_h)) {
        ship_alive = 0;
        S.sti(SI_W(ship_alive), 0);
        S.sti(SI_W(ship_death_counter), SI_SHIP_DEATH_TIME);
        S.sti(SI_W(ship_death_hit_1), 1);
        keep = false;
      }
      if (keep) { si_store_laser(S, SI_W(enemy_lasers) + SI_LASER_WORDS * n, l.x, l.y, l.w, l.h, l.t, l.movement, l.speed, l.color); n++; }
    }
    S.sti(SI_W(n_enemy_lasers), n);
  }
  /* wave cleared */
  if (n_alive == 0) {
    bool busy = false;
    for (int i = 0; i < TBX_SI_N_ENEMIES; i++) busy = busy || S.ldi(SI_W(en_death) + i) != TBX_NONE;
    if (!busy) {
      S.sti(TBX_HW(level), S.ldi(TBX_HW(level)) + 1);
      si_reset_enemies(S, c);
      si_reset_shields(S, c);
      S.sti(SI_W(n_enemy_lasers), 0);
    }
  }
}

/* draw list in slot order: ground, score, lives, shields, enemies, ship, ufo, ship laser, enemy lasers */
TBX_HD TbxPrim si_prim(const uint32_t *R, int slot) {
  TbxAcc S; S.p = const_cast<uint32_t *>(R); S.stride = 1;
  if (slot == SI_SLOT_GROUND) return tbx_prim_rect(SI_COLOR_GROUND, 0, SI_GROUND_Y, TBX_SI_W, 2);
  if (slot < SI_SLOT_LIVES) return tbx_prim_digit(SI_COLOR_HUD, 120, 2, S.ldi(TBX_HW(score)), 2, 1, slot - SI_SLOT_SCORE);
  if (slot < SI_SLOT_SHIELDS) {
    if (S.ldi(SI_W(life_display_timer)) <= 0) return tbx_prim_none();
    return tbx_prim_digit(SI_COLOR_HUD, 170, 199, S.ldi(TBX_HW(lives)), 3, 2, slot - SI_SLOT_LIVES);
  }
  if (slot < SI_SLOT_ENEMIES) {
    int i = slot - SI_SLOT_SHIELDS;
    return tbx_prim_sprite(SI_COLOR_SHIELD, S.ldi(SI_W(shield_x) + i), S.ldi(SI_W(shield_y) + i), 16, TBX_SI_SHIELD_H,
                           TBX_PRIM_STATE | (uint32_t)(SI_W(shield_rows) + i * TBX_SI_SHIELD_H), 1, 1);
  }
  if (slot < SI_SLOT_SHIP) {
    int i = slot - SI_SLOT_ENEMIES;
    int ex = S.ldi(SI_W(en_x) + i), ey = S.ldi(SI_W(en_y) + i);
    if (si_en_alive(S, i)) {
      int row = S.ldi(SI_W(en_row) + i);
      int kind = row < 0 ? 0 : row > 5 ? 2 : row / 2;
      return tbx_prim_sprite(SI_COLOR_ENEMY, ex, ey, 16, SI_ENEMY_H,
                             TBX_BANK_INVADER + 20 * kind + (S.ldi(SI_W(visual_orientation)) ? 10 : 0), 1, 1);
    }
    if (S.ldi(SI_W(en_death) + i) != TBX_NONE) return tbx_prim_sprite(SI_COLOR_ENEMY, ex, ey, 16, SI_ENEMY_H, TBX_BANK_BOOM, 1, 1);
    return tbx_prim_none();
  }
  if (slot == SI_SLOT_SHIP) {
    uint32_t col = S.ld(SI_W(ship_color));
    if (S.ldi(SI_W(ship_alive))) return tbx_prim_sprite(col, S.ldi(SI_W(ship_x)), S.ldi(SI_W(ship_y)), 16, SI_ENEMY_H, TBX_BANK_SHIP, 1, 1);
    if (S.ldi(SI_W(ship_death_counter)) != TBX_NONE)
      return tbx_prim_sprite(col, S.ldi(SI_W(ship_x)), S.ldi(SI_W(ship_y)), 16, SI_ENEMY_H,
                             TBX_BANK_BOOM + (S.ldi(SI_W(ship_death_hit_1)) ? 0 : 10), 1, 1);
    return tbx_prim_none();
  }
  if (slot == SI_SLOT_UFO) {
    if (S.ldi(SI_W(ufo_death_counter)) != TBX_NONE)
      return tbx_prim_sprite(SI_COLOR_UFO, S.ldi(SI_W(ufo_x)), S.ldi(SI_W(ufo_y)), 16, SI_ENEMY_H, TBX_BANK_BOOM + 10, 1, 1);
    if (S.ldi(SI_W(ufo_appearance_counter)) == TBX_NONE)
      return tbx_prim_sprite(SI_COLOR_UFO, S.ldi(SI_W(ufo_x)), S.ldi(SI_W(ufo_y)), 16, SI_UFO_H, TBX_BANK_UFO, 1, 1);
    return tbx_prim_none();
  }
  int w;
  if (slot == SI_SLOT_SHIP_LASER) {
    if (!S.ldi(SI_W(has_ship_laser))) return tbx_prim_none();
    w = SI_W(ship_laser);
  } else {
    int i = slot - SI_SLOT_ENEMY_LASERS;
    if (i >= S.ldi(SI_W(n_enemy_lasers))) return tbx_prim_none();
    w = SI_W(enemy_lasers) + SI_LASER_WORDS * i;
  }
  return tbx_prim_rect(S.ld(w + 7), S.ldi(w), S.ldi(w + 1), S.ldi(w + 2), S.ldi(w + 3));
}

#endif
