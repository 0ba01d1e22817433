/* tbx_amidar.h -- Amidar transition, new_game and draw list, one env per call.
 *
 * Replaces what Toybox('amidar').apply_ale_action / new_game / get_state reach in ctoybox (reference
 * call sites toybox/envs/atari/base.py:126,153,109; state fields toybox/interventions/amidar.py:22-34,
 * :83-166 MovementAI (five protocols), :171 Enemy, :195 Player, :216 Board, :300 Box, :316 TilePoint;
 * queries tile_to_world / world_to_tile amidar.py:508-518; constants and the initial state from
 * toybox/interventions/defaults/amidar_{config,state}_default.json).  All arithmetic is i32.
 * Board geometry that new_game derives from the config (boxes, junctions) lives in AmiTable.
 */
#ifndef TBX_AMIDAR_H
#define TBX_AMIDAR_H
#include "tbx_common.h"

#define AMI_W(f) TBX_W(AmiRec, f)
#define AMI_TW 64 /* tile width in world units  (4 px * 16) */
#define AMI_TH 80 /* tile height in world units (5 px * 16) */
#define AMI_OFF_X 16
#define AMI_OFF_Y 37
#define AMI_MOB_WORDS TBX_WORDS(AmiMob)
#define AMI_MW(f) TBX_W(AmiMob, f)
#define AMI_AIW(f) (TBX_W(AmiMob, ai) + TBX_W(AmiAi, f))
#define AMI_PLAYER AMI_W(player)
#define AMI_ENEMY(i) (AMI_W(enemies) + (i) * AMI_MOB_WORDS)

/* draw-list slots */
#define AMI_SLOT_TILES 0                                     /* 31*32, row-major */
#define AMI_SLOT_BOXES (TBX_AMI_BW * TBX_AMI_BH)             /* TBX_AMI_MAX_BOXES */
#define AMI_SLOT_ENEMIES (AMI_SLOT_BOXES + TBX_AMI_MAX_BOXES)
#define AMI_SLOT_PLAYER (AMI_SLOT_ENEMIES + TBX_AMI_MAX_ENEMIES)
#define AMI_SLOT_SCORE (AMI_SLOT_PLAYER + 1)
#define AMI_SLOT_LIVES (AMI_SLOT_SCORE + TBX_MAX_DIGITS)
#define AMI_SLOT_JUMPS (AMI_SLOT_LIVES + TBX_MAX_DIGITS)
#define AMI_N_SLOTS (AMI_SLOT_JUMPS + TBX_MAX_DIGITS)
#define AMI_N_STATIC 0
#define AMI_N_GROUPS 4
/* tiles (a grid) | painted box interiors (one colour) | enemies (one colour) | player, score, lives, jumps (in order) */
TBX_HD void ami_group(int g, int &b, int &e, int &mode) {
  if (g == 0) { b = AMI_SLOT_TILES; e = AMI_SLOT_BOXES; mode = TBX_GROUP_PARALLEL; }
  else if (g == 1) { b = AMI_SLOT_BOXES; e = AMI_SLOT_ENEMIES; mode = TBX_GROUP_PARALLEL; }
  else if (g == 2) { b = AMI_SLOT_ENEMIES; e = AMI_SLOT_PLAYER; mode = TBX_GROUP_PARALLEL; }
  else { b = AMI_SLOT_PLAYER; e = AMI_N_SLOTS; mode = TBX_GROUP_SERIAL; }
}

TBX_HD int ami_floordiv(int a, int b) { int q = a / b; if ((a % b) != 0 && ((a < 0) != (b < 0))) q--; return q; }
TBX_HD int ami_dx(int d) { return d == TBX_DIR_LEFT ? -1 : d == TBX_DIR_RIGHT ? 1 : 0; }
TBX_HD int ami_dy(int d) { return d == TBX_DIR_UP ? -1 : d == TBX_DIR_DOWN ? 1 : 0; }
TBX_HD int ami_opposite(int d) { return d ^ 1; }

TBX_HD int ami_tile(const TbxAcc &S, int tx, int ty) { return (int)((S.ld(AMI_W(tiles) + 2 * ty + (tx >> 4)) >> (2 * (tx & 15))) & 3u); }
TBX_HD void ami_set_tile(const TbxAcc &S, int tx, int ty, int v) {
  int w = AMI_W(tiles) + 2 * ty + (tx >> 4), sh = 2 * (tx & 15);
  S.st(w, (S.ld(w) & ~(3u << sh)) | ((uint32_t)v << sh));
}
TBX_HD bool ami_walkable(const TbxAcc &S, int tx, int ty) {
  return tx >= 0 && tx < TBX_AMI_BW && ty >= 0 && ty < TBX_AMI_BH && ami_tile(S, tx, ty) != TBX_TILE_EMPTY;
}
/* a junction is a walkable tile that has both a horizontal and a vertical walkable neighbour */
TBX_HD bool ami_junction_tile(const TbxAcc &S, int tx, int ty) {
  if (!ami_walkable(S, tx, ty)) return false;
  bool h = ami_walkable(S, tx - 1, ty) || ami_walkable(S, tx + 1, ty);
  bool v = ami_walkable(S, tx, ty - 1) || ami_walkable(S, tx, ty + 1);
  return h && v;
}
/* membership in the board's junction id list */
TBX_HD bool ami_is_junction(const AmiTable &T, int id) {
  if (id >= 0 && id < TBX_AMI_BW * TBX_AMI_BH) return (T.junction_bits[id >> 5] >> (id & 31)) & 1u;
  for (int i = 0; i < T.n_junctions; i++) if (T.junctions[i] == id) return true; /* out-of-range ids (interventions) */
  return false;
}

TBX_HD void ami_enemy_start_tile(const TbxAcc &S, const AmiCfg &c, int m, int &tx, int &ty) {
  if (S.ldi(m + AMI_AIW(kind)) == TBX_AI_LOOKUP) {
    int id = 0;
    if (c.n_routes > 0) {
      int r = S.ldi(m + AMI_AIW(default_route_index)) % c.n_routes;
      if (r < 0) r += c.n_routes;
      if (c.route_len[r] > 0) id = c.routes[r][0];
    }
    tx = id % TBX_AMI_BW; ty = id / TBX_AMI_BW;
  } else {
    tx = S.ldi(m + AMI_AIW(start_tx)); ty = S.ldi(m + AMI_AIW(start_ty));
  }
}
TBX_HD void ami_reset_ai(const TbxAcc &S, int m) {
  int kind = S.ldi(m + AMI_AIW(kind));
  if (kind == TBX_AI_LOOKUP) S.sti(m + AMI_AIW(next), 0);
  else if (kind == TBX_AI_AMIDAR) { S.sti(m + AMI_AIW(vert), S.ldi(m + AMI_AIW(start_vert))); S.sti(m + AMI_AIW(horiz), S.ldi(m + AMI_AIW(start_horiz))); }
  else if (kind == TBX_AI_TARGET) { S.sti(m + AMI_AIW(dir), S.ldi(m + AMI_AIW(start_dir))); S.sti(m + AMI_AIW(has_seen), 0); S.sti(m + AMI_AIW(seen_tx), 0); S.sti(m + AMI_AIW(seen_ty), 0); }
  else if (kind == TBX_AI_RANDOM) S.sti(m + AMI_AIW(dir), S.ldi(m + AMI_AIW(start_dir)));
}
TBX_HD void ami_place_mob(const TbxAcc &S, int m, int tx, int ty) {
  S.sti(m + AMI_MW(x), tx * AMI_TW); S.sti(m + AMI_MW(y), ty * AMI_TH);
  S.sti(m + AMI_MW(has_step), 0); S.sti(m + AMI_MW(step_tx), 0); S.sti(m + AMI_MW(step_ty), 0); S.sti(m + AMI_MW(caught), 0);
}
TBX_HD void ami_reset_player(const TbxAcc &S, const AmiCfg &c, const AmiTable &T) {
  ami_place_mob(S, AMI_PLAYER, c.player_start_tx, c.player_start_ty);
  S.sti(AMI_PLAYER + AMI_MW(n_history), 0);
  /* history starts with the first junction at or below the start tile (fixture: [607]) */
  for (int ty = c.player_start_ty; ty < TBX_AMI_BH && ami_walkable(S, c.player_start_tx, ty); ty++)
    if (ami_is_junction(T, ty * TBX_AMI_BW + c.player_start_tx)) {
      S.sti(AMI_PLAYER + AMI_MW(history), ty * TBX_AMI_BW + c.player_start_tx);
      S.sti(AMI_PLAYER + AMI_MW(n_history), 1);
      break;
    }
}
TBX_HD void ami_reset_enemy(const TbxAcc &S, const AmiCfg &c, int m) {
  int tx, ty;
  ami_reset_ai(S, m);
  ami_enemy_start_tile(S, c, m, tx, ty);
  ami_place_mob(S, m, tx, ty);
  S.sti(m + AMI_MW(n_history), 0);
}
TBX_HD void ami_store_ai(const TbxAcc &S, int m, const AmiAi &a) {
  const int32_t *src = &a.kind;
  for (int k = 0; k < TBX_WORDS(AmiAi); k++) S.sti(m + TBX_W(AmiMob, ai) + k, src[k]);
}

/* new_game: the child rng is drawn from a COPY of the simulator rng (Amidar's config rand is not advanced,
 * SURVEY App. A.2) */
TBX_HD void ami_new_game(const TbxAcc &S, const AmiCfg &c, const AmiTable *tables) {
  TbxRng sim = tbx_rng_load(S, TBX_HW(sim_rand));
  TbxRng rng;
  rng.s0 = tbx_rng_next_u64(sim);
  rng.s1 = tbx_rng_next_u64(sim);
  tbx_rng_store(S, TBX_HW(rand), rng);
  const AmiTable &T = tables[c.default_tbl];
  S.sti(TBX_HW(lives), c.start_lives); S.sti(TBX_HW(score), 0); S.sti(TBX_HW(level), 1);
  S.sti(TBX_HW(prev_score), 0); S.sti(TBX_HW(ep_len), 0); S.sti(TBX_HW(ep_return), 0);
  S.sti(TBX_HW(tbl), c.default_tbl);
  S.sti(AMI_W(jumps), c.start_jumps); S.sti(AMI_W(jump_timer), 0); S.sti(AMI_W(chase_timer), 0);
  S.sti(AMI_W(n_enemies), c.n_enemies);
  S.st(AMI_W(box_painted), 0);
  for (int ty = 0; ty < TBX_AMI_BH; ty++) { S.st(AMI_W(tiles) + 2 * ty, c.board[ty][0]); S.st(AMI_W(tiles) + 2 * ty + 1, c.board[ty][1]); }
  for (int k = 0; k < AMI_MOB_WORDS * (1 + TBX_AMI_MAX_ENEMIES); k++) S.st(AMI_PLAYER + k, 0);
  S.sti(AMI_PLAYER + AMI_MW(speed), 8);
  S.sti(AMI_PLAYER + AMI_AIW(kind), TBX_AI_PLAYER);
  ami_reset_player(S, c, T);
  for (int i = 0; i < c.n_enemies; i++) {
    ami_store_ai(S, AMI_ENEMY(i), c.enemies[i]);
    S.sti(AMI_ENEMY(i) + AMI_MW(speed), 8);
    ami_reset_enemy(S, c, AMI_ENEMY(i));
  }
}

/* advance a mob toward its step tile; returns true on arrival */
TBX_HD bool ami_advance(const TbxAcc &S, int m) {
  int gx = S.ldi(m + AMI_MW(step_tx)) * AMI_TW, gy = S.ldi(m + AMI_MW(step_ty)) * AMI_TH;
  int x = S.ldi(m + AMI_MW(x)), y = S.ldi(m + AMI_MW(y)), sp = S.ldi(m + AMI_MW(speed));
  if (sp < 0) sp = 0;
  int dx = gx - x, dy = gy - y;
  if (dx > sp) dx = sp;
  if (dx < -sp) dx = -sp;
  if (dy > sp) dy = sp;
  if (dy < -sp) dy = -sp;
  x += dx; y += dy;
  S.sti(m + AMI_MW(x), x); S.sti(m + AMI_MW(y), y);
  if (x == gx && y == gy) { S.sti(m + AMI_MW(has_step), 0); return true; }
  return false;
}
TBX_HD bool ami_aligned(int x, int y) { return (x % AMI_TW) == 0 && (y % AMI_TH) == 0; }

TBX_HD void ami_check_boxes(const TbxAcc &S, const AmiCfg &c, const AmiTable &T, int &score) {
  uint32_t painted = S.ld(AMI_W(box_painted));
  bool newly_chase = false;
  for (int i = 0; i < T.n_boxes; i++) {
    if ((painted >> i) & 1u) continue;
    bool ok = true;
    int x0 = T.tl_tx[i], y0 = T.tl_ty[i], x1 = T.br_tx[i], y1 = T.br_ty[i];
    for (int tx = x0; tx <= x1 && ok; tx++) ok = ami_tile(S, tx, y0) == TBX_TILE_PAINTED && ami_tile(S, tx, y1) == TBX_TILE_PAINTED;
    for (int ty = y0; ty <= y1 && ok; ty++) ok = ami_tile(S, x0, ty) == TBX_TILE_PAINTED && ami_tile(S, x1, ty) == TBX_TILE_PAINTED;
    if (!ok) continue;
    painted |= 1u << i;
    score += c.box_bonus;
    if ((T.triggers_chase >> i) & 1u) newly_chase = true;
  }
  S.st(AMI_W(box_painted), painted);
  if (newly_chase && (T.triggers_chase & T.all_boxes & ~painted) == 0) S.sti(AMI_W(chase_timer), c.chase_time);
}

TBX_HD void ami_player_arrived(const TbxAcc &S, const AmiCfg &c, const AmiTable &T, int tx, int ty, int &score) {
  int id = ty * TBX_AMI_BW + tx;
  if (!ami_is_junction(T, id)) return;
  int nh = S.ldi(AMI_PLAYER + AMI_MW(n_history));
  int h0 = nh > 0 ? S.ldi(AMI_PLAYER + AMI_MW(history)) : 0;
  if (nh > 0 && h0 != id) {
    /* C-style truncating % and / so that (intervention-made) negative ids behave like the restated rule */
    int px = h0 % TBX_AMI_BW, py = h0 / TBX_AMI_BW;
    bool ok = (px == tx || py == ty);
    int sx = px < tx ? 1 : px > tx ? -1 : 0, sy = py < ty ? 1 : py > ty ? -1 : 0;
    if (ok)
      for (int x = px, y = py;; x += sx, y += sy) {
        if (!ami_walkable(S, x, y)) ok = false;
        if (x == tx && y == ty) break;
      }
    if (ok) {
      int newly = 0;
      for (int x = px, y = py;; x += sx, y += sy) {
        if (ami_tile(S, x, y) != TBX_TILE_PAINTED) { ami_set_tile(S, x, y, TBX_TILE_PAINTED); newly++; }
        if (x == tx && y == ty) break;
      }
      score += newly;
      if (newly > 0) ami_check_boxes(S, c, T, score);
    }
  }
  if (nh == 0 || h0 != id) {
    int n = nh < TBX_AMI_HIST ? nh : TBX_AMI_HIST - 1;
    for (int i = n; i > 0; i--) S.sti(AMI_PLAYER + AMI_MW(history) + i, S.ldi(AMI_PLAYER + AMI_MW(history) + i - 1));
    S.sti(AMI_PLAYER + AMI_MW(history), id);
    S.sti(AMI_PLAYER + AMI_MW(n_history), n + 1);
  }
}

/* head for tile (gx,gy): along the shared row/column when there is one, else x first then y */
TBX_HD int ami_toward(const TbxAcc &S, int tx, int ty, int gx, int gy) {
  int dh = gx > tx ? TBX_DIR_RIGHT : TBX_DIR_LEFT, dv = gy > ty ? TBX_DIR_DOWN : TBX_DIR_UP;
  if (gy == ty && gx != tx) return ami_walkable(S, tx + ami_dx(dh), ty) ? dh : -1;
  if (gx == tx && gy != ty) return ami_walkable(S, tx, ty + ami_dy(dv)) ? dv : -1;
  if (gx != tx && ami_walkable(S, tx + ami_dx(dh), ty)) return dh;
  if (gy != ty && ami_walkable(S, tx, ty + ami_dy(dv))) return dv;
  return -1;
}
/* walkable neighbours in Up,Down,Left,Right order as a 4-bit set, without the reverse of `dir` unless dead end */
TBX_HD uint32_t ami_options(const TbxAcc &S, int tx, int ty, int dir) {
  uint32_t set = 0;
  for (int d = 0; d < 4; d++)
    if (d != ami_opposite(dir) && ami_walkable(S, tx + ami_dx(d), ty + ami_dy(d))) set |= 1u << d;
  int o = ami_opposite(dir);
  if (set == 0 && ami_walkable(S, tx + ami_dx(o), ty + ami_dy(o))) set |= 1u << (o & 31);
  return set;
}
TBX_HD int ami_nth_bit(uint32_t set, int k) {
  for (int d = 0; d < 32; d++)
    if ((set >> d) & 1u) { if (k == 0) return d; k--; }
  return -1;
}
TBX_HD int ami_random_dir(const TbxAcc &S, TbxRng &rng, int tx, int ty, int dir) {
  uint32_t set = ami_options(S, tx, ty, dir);
  int n = tbx_popc(set);
  if (n == 0) return -1;
  if (!ami_junction_tile(S, tx, ty) || n == 1) {
    if (dir >= 0 && dir < 32 && ((set >> dir) & 1u)) return dir;
    return ami_nth_bit(set, 0);
  }
  return ami_nth_bit(set, (int)tbx_rng_index(rng, (uint32_t)n));
}
TBX_HD int ami_abs(int v) { return v < 0 ? -v : v; }

/* pick the next direction for the enemy at record offset m standing (aligned) on (tx,ty); -1 = stay */
TBX_HD int ami_enemy_choose(const TbxAcc &S, const AmiCfg &c, int m, int tx, int ty) {
  int kind = S.ldi(m + AMI_AIW(kind));
  if (kind == TBX_AI_LOOKUP) {
    if (c.n_routes <= 0) return -1;
    int r = S.ldi(m + AMI_AIW(default_route_index)) % c.n_routes;
    if (r < 0) r += c.n_routes;
    int len = c.route_len[r];
    if (len <= 0) return -1;
    int nx = S.ldi(m + AMI_AIW(next)) % len;
    if (nx < 0) nx += len;
    if (c.routes[r][nx] == ty * TBX_AMI_BW + tx) nx = (nx + 1) % len;
    S.sti(m + AMI_AIW(next), nx);
    int g = c.routes[r][nx];
    return ami_toward(S, tx, ty, g % TBX_AMI_BW, g / TBX_AMI_BW);
  }
  if (kind == TBX_AI_PERIMETER) {
    int d = -1;
    if (ty == 0 && tx < TBX_AMI_BW - 1) d = TBX_DIR_RIGHT;
    else if (tx == TBX_AMI_BW - 1 && ty < TBX_AMI_BH - 1) d = TBX_DIR_DOWN;
    else if (ty == TBX_AMI_BH - 1 && tx > 0) d = TBX_DIR_LEFT;
    else if (tx == 0 && ty > 0) d = TBX_DIR_UP;
    if (d >= 0 && ami_walkable(S, tx + ami_dx(d), ty + ami_dy(d))) return d;
    return ami_toward(S, tx, ty, S.ldi(m + AMI_AIW(start_tx)), S.ldi(m + AMI_AIW(start_ty)));
  }
  if (kind == TBX_AI_AMIDAR) {
    int vert = S.ldi(m + AMI_AIW(vert)), horiz = S.ldi(m + AMI_AIW(horiz)), d = -1;
    if (vert != TBX_DIR_UP && vert != TBX_DIR_DOWN) vert = TBX_DIR_DOWN;
    if (horiz != TBX_DIR_LEFT && horiz != TBX_DIR_RIGHT) horiz = TBX_DIR_RIGHT;
    if (ami_walkable(S, tx, ty + ami_dy(vert))) d = vert;
    else {
      if (ty == 0 || ty == TBX_AMI_BH - 1) vert = ami_opposite(vert);
      if (ami_walkable(S, tx + ami_dx(horiz), ty)) d = horiz;
      else {
        horiz = ami_opposite(horiz);
        if (ami_walkable(S, tx + ami_dx(horiz), ty)) d = horiz;
        else {
          vert = ami_opposite(vert);
          if (ami_walkable(S, tx, ty + ami_dy(vert))) d = vert;
        }
      }
    }
    S.sti(m + AMI_AIW(vert), vert); S.sti(m + AMI_AIW(horiz), horiz);
    return d;
  }
  if (kind == TBX_AI_TARGET) {
    int ptx = ami_floordiv(S.ldi(AMI_PLAYER + AMI_MW(x)), AMI_TW), pty = ami_floordiv(S.ldi(AMI_PLAYER + AMI_MW(y)), AMI_TH);
    int dist = ami_abs(ptx - tx) + ami_abs(pty - ty), d;
    int dir = S.ldi(m + AMI_AIW(dir));
    if (dir < 0 || dir > 3) dir = TBX_DIR_UP;
    int has_seen = S.ldi(m + AMI_AIW(has_seen)), stx = S.ldi(m + AMI_AIW(seen_tx)), sty = S.ldi(m + AMI_AIW(seen_ty));
    if (dist <= S.ldi(m + AMI_AIW(vision_distance))) { has_seen = 1; stx = ptx; sty = pty; }
    if (has_seen && stx == tx && sty == ty) { has_seen = 0; stx = 0; sty = 0; }
    if (has_seen) {
      uint32_t set = ami_options(S, tx, ty, dir);
      int best = 0x7fffffff;
      d = -1;
      /* candidates in the order the option list is built: non-reverse directions Up..Right, else the reverse */
      for (int k = 0; k < 4; k++)
        if ((set >> k) & 1u) {
          int nx = tx + ami_dx(k), ny = ty + ami_dy(k);
          int mm = ami_abs(stx - nx) + ami_abs(sty - ny);
          if (mm < best) { best = mm; d = k; }
        }
    } else {
      TbxRng rng = tbx_rng_load(S, TBX_HW(rand));
      d = ami_random_dir(S, rng, tx, ty, dir);
      tbx_rng_store(S, TBX_HW(rand), rng);
    }
    if (d >= 0) dir = d;
    S.sti(m + AMI_AIW(dir), dir); S.sti(m + AMI_AIW(has_seen), has_seen);
    S.sti(m + AMI_AIW(seen_tx), stx); S.sti(m + AMI_AIW(seen_ty), sty);
    return d;
  }
  if (kind == TBX_AI_RANDOM) {
    int dir = S.ldi(m + AMI_AIW(dir));
    if (dir < 0 || dir > 3) dir = TBX_DIR_UP;
    TbxRng rng = tbx_rng_load(S, TBX_HW(rand));
    int d = ami_random_dir(S, rng, tx, ty, dir);
    tbx_rng_store(S, TBX_HW(rand), rng);
    if (d >= 0) dir = d;
    S.sti(m + AMI_AIW(dir), dir);
    return d;
  }
  return -1;
}

/* off-grid mob (after an intervention): walk to the containing tile */
TBX_HD void ami_snap_step(const TbxAcc &S, int m, int x, int y) {
  int tx = ami_floordiv(x, AMI_TW), ty = ami_floordiv(y, AMI_TH);
  if (tx < 0) tx = 0;
  if (tx > TBX_AMI_BW - 1) tx = TBX_AMI_BW - 1;
  if (ty < 0) ty = 0;
  if (ty > TBX_AMI_BH - 1) ty = TBX_AMI_BH - 1;
  S.sti(m + AMI_MW(has_step), 1); S.sti(m + AMI_MW(step_tx), tx); S.sti(m + AMI_MW(step_ty), ty);
}

TBX_HD void ami_step(const TbxAcc &S, const AmiCfg &c, const AmiTable *tables, int in) {
  int lives = S.ldi(TBX_HW(lives));
  if (lives <= 0) return;
  const AmiTable &T = tables[S.ldi(TBX_HW(tbl))];
  int n_enemies = S.ldi(AMI_W(n_enemies));
  if (n_enemies > TBX_AMI_MAX_ENEMIES) n_enemies = TBX_AMI_MAX_ENEMIES;
  int score = S.ldi(TBX_HW(score));
  /* timers */
  int jump_timer = S.ldi(AMI_W(jump_timer)), chase_timer = S.ldi(AMI_W(chase_timer));
  if (jump_timer > 0) jump_timer -= 1;
  if (chase_timer > 0) {
    chase_timer -= 1;
    if (chase_timer == 0)
      for (int i = 0; i < n_enemies; i++)
        if (S.ldi(AMI_ENEMY(i) + AMI_MW(caught))) ami_reset_enemy(S, c, AMI_ENEMY(i));
  }
  S.sti(AMI_W(chase_timer), chase_timer);
  /* jump */
  if ((in & TBX_IN_BUTTON1) && jump_timer == 0) {
    int jumps = S.ldi(AMI_W(jumps));
    if (jumps > 0) { S.sti(AMI_W(jumps), jumps - 1); jump_timer = c.jump_time; }
  }
  /* player */
  {
    const int m = AMI_PLAYER;
    int has_step = S.ldi(m + AMI_MW(has_step));
    if (!has_step) {
      int x = S.ldi(m + AMI_MW(x)), y = S.ldi(m + AMI_MW(y));
      if (!ami_aligned(x, y)) { ami_snap_step(S, m, x, y); has_step = 1; }
      else {
        int tx = x / AMI_TW, ty = y / AMI_TH;
        for (int d = 0; d < 4; d++) { /* tried in the order Up, Down, Left, Right */
          int bit = d == TBX_DIR_UP ? TBX_IN_UP : d == TBX_DIR_DOWN ? TBX_IN_DOWN : d == TBX_DIR_LEFT ? TBX_IN_LEFT : TBX_IN_RIGHT;
          if ((in & bit) && ami_walkable(S, tx + ami_dx(d), ty + ami_dy(d))) {
            S.sti(m + AMI_MW(has_step), 1); S.sti(m + AMI_MW(step_tx), tx + ami_dx(d)); S.sti(m + AMI_MW(step_ty), ty + ami_dy(d));
            has_step = 1;
            break;
          }
        }
      }
    }
    if (has_step) {
      int tx = S.ldi(m + AMI_MW(step_tx)), ty = S.ldi(m + AMI_MW(step_ty));
      if (ami_advance(S, m)) ami_player_arrived(S, c, T, tx, ty, score);
    }
  }
  chase_timer = S.ldi(AMI_W(chase_timer)); /* painting the last corner box starts the chase this very frame */
  /* enemies */
  for (int i = 0; i < n_enemies; i++) {
    const int m = AMI_ENEMY(i);
    if (S.ldi(m + AMI_MW(caught))) continue;
    int has_step = S.ldi(m + AMI_MW(has_step));
    if (!has_step) {
      int x = S.ldi(m + AMI_MW(x)), y = S.ldi(m + AMI_MW(y));
      if (!ami_aligned(x, y)) { ami_snap_step(S, m, x, y); has_step = 1; }
      else {
        int tx = x / AMI_TW, ty = y / AMI_TH;
        int d = ami_enemy_choose(S, c, m, tx, ty);
        if (d >= 0) {
          S.sti(m + AMI_MW(has_step), 1); S.sti(m + AMI_MW(step_tx), tx + ami_dx(d)); S.sti(m + AMI_MW(step_ty), ty + ami_dy(d));
          has_step = 1;
        }
      }
    }
    if (has_step) ami_advance(S, m);
  }
  /* contact */
  {
    int px = S.ldi(AMI_PLAYER + AMI_MW(x)), py = S.ldi(AMI_PLAYER + AMI_MW(y));
    for (int i = 0; i < n_enemies; i++) {
      const int m = AMI_ENEMY(i);
      if (S.ldi(m + AMI_MW(caught))) continue;
      int dx = ami_abs(S.ldi(m + AMI_MW(x)) - px), dy = ami_abs(S.ldi(m + AMI_MW(y)) - py);
      if (dx > AMI_TW / 2 || dy > AMI_TH / 2) continue;
      if (jump_timer > 0) continue;
      if (chase_timer > 0) { S.sti(m + AMI_MW(caught), 1); score += c.chase_score_bonus; continue; }
      lives -= 1;
      S.sti(TBX_HW(lives), lives);
      jump_timer = 0;
      ami_reset_player(S, c, T);
      for (int k = 0; k < n_enemies; k++) ami_reset_enemy(S, c, AMI_ENEMY(k));
      break;
    }
  }
  /* level complete: every box painted */
  if (T.n_boxes > 0 && (T.all_boxes & ~S.ld(AMI_W(box_painted))) == 0) {
    S.sti(TBX_HW(level), S.ldi(TBX_HW(level)) + 1);
    for (int ty = 0; ty < TBX_AMI_BH; ty++) { S.st(AMI_W(tiles) + 2 * ty, c.board[ty][0]); S.st(AMI_W(tiles) + 2 * ty + 1, c.board[ty][1]); }
    S.st(AMI_W(box_painted), 0);
    S.sti(AMI_W(jumps), c.start_jumps);
    jump_timer = 0;
    S.sti(AMI_W(chase_timer), 0);
    ami_reset_player(S, c, T);
    for (int k = 0; k < n_enemies; k++) ami_reset_enemy(S, c, AMI_ENEMY(k));
  }
  S.sti(AMI_W(jump_timer), jump_timer);
  S.sti(TBX_HW(score), score);
}

/* draw list in slot order: tiles, painted box interiors, enemies, player, score, lives, jumps */
TBX_HD TbxPrim ami_prim(const uint32_t *R, const AmiCfg &c, const AmiTable *tables, int slot) {
  TbxAcc S; S.p = const_cast<uint32_t *>(R); S.stride = 1;
  if (slot < AMI_SLOT_BOXES) {
    int tx = slot & 31, ty = slot >> 5;
    int t = ami_tile(S, tx, ty);
    if (t == TBX_TILE_EMPTY) return tbx_prim_none();
    return tbx_prim_rect(t == TBX_TILE_PAINTED ? c.painted_color : c.unpainted_color, AMI_OFF_X + 4 * tx, AMI_OFF_Y + 5 * ty, 4, 5);
  }
  if (slot < AMI_SLOT_ENEMIES) {
    int i = slot - AMI_SLOT_BOXES;
    const AmiTable &T = tables[S.ldi(TBX_HW(tbl))];
    if (i >= T.n_boxes || !((S.ld(AMI_W(box_painted)) >> i) & 1u)) return tbx_prim_none();
    return tbx_prim_rect(c.inner_painted_color, AMI_OFF_X + 4 * (T.tl_tx[i] + 1), AMI_OFF_Y + 5 * (T.tl_ty[i] + 1),
                         4 * (T.br_tx[i] - T.tl_tx[i] - 1), 5 * (T.br_ty[i] - T.tl_ty[i] - 1));
  }
  if (slot <= AMI_SLOT_PLAYER) {
    int m;
    uint32_t col;
    if (slot == AMI_SLOT_PLAYER) { m = AMI_PLAYER; col = c.player_color; }
    else {
      int i = slot - AMI_SLOT_ENEMIES;
      if (i >= S.ldi(AMI_W(n_enemies))) return tbx_prim_none();
      m = AMI_ENEMY(i); col = c.enemy_color;
      if (S.ldi(m + AMI_MW(caught))) return tbx_prim_none();
    }
    return tbx_prim_rect(col, AMI_OFF_X + ami_floordiv(S.ldi(m + AMI_MW(x)), 16) - 1, AMI_OFF_Y + ami_floordiv(S.ldi(m + AMI_MW(y)), 16) - 1, 6, 7);
  }
  if (slot < AMI_SLOT_LIVES) return tbx_prim_digit(c.player_color, 100, 205, S.ldi(TBX_HW(score)), 2, 2, slot - AMI_SLOT_SCORE);
  if (slot < AMI_SLOT_JUMPS) return tbx_prim_digit(c.player_color, 124, 205, S.ldi(TBX_HW(lives)), 2, 2, slot - AMI_SLOT_LIVES);
  return tbx_prim_digit(c.painted_color, 144, 205, S.ldi(AMI_W(jumps)), 2, 2, slot - AMI_SLOT_JUMPS);
}

/* Delta rendering.  Base frame 1 holds the config's board (every tile as new_game lays it out); the tile grid is
 * fixed and sits on the background, so an env is base 1 plus the tiles whose LOOK (empty / unpainted / painted)
 * differs from the config board, painted in their current look. */
TBX_HD int ami_tile_look(int t) { return t == TBX_TILE_EMPTY ? 0 : t == TBX_TILE_PAINTED ? 2 : 1; }
TBX_HD int ami_base_id(const uint32_t *, const AmiCfg &, const AmiTable *) { return 1; }
/* the look of tile (tx, ty) if it differs from the config board's, else -1 */
TBX_HD int ami_delta_look(const uint32_t *R, const AmiCfg &c, int tx, int ty) {
  const int look = ami_tile_look((int)((R[AMI_W(tiles) + 2 * ty + (tx >> 4)] >> (2 * (tx & 15))) & 3u));
  const int ref = ami_tile_look((int)((c.board[ty][tx >> 4] >> (2 * (tx & 15))) & 3u));
  return look == ref ? -1 : look;
}
TBX_HD TbxPrim ami_prim_delta(const uint32_t *R, const AmiCfg &c, const AmiTable *tables, int slot, int base) {
  if (base == 1 && slot < AMI_SLOT_BOXES) {
    const int tx = slot & 31, ty = slot >> 5;
    const int look = ami_delta_look(R, c, tx, ty);
    if (look < 0) return tbx_prim_none();
    /* Horizontally adjacent tiles that differ from the config board in the same way are ONE rectangle, emitted by the
     * leftmost of the run (slots are row-major): painted corridors are long lines, not hundreds of 4 x 5 cells. */
    if (tx > 0 && ami_delta_look(R, c, tx - 1, ty) == look) return tbx_prim_none();
    int len = 1;
    while (tx + len < TBX_AMI_BW && ami_delta_look(R, c, tx + len, ty) == look) len++;
    return tbx_prim_rect(look == 0 ? c.bg_color : look == 2 ? c.painted_color : c.unpainted_color, AMI_OFF_X + 4 * tx, AMI_OFF_Y + 5 * ty, 4 * len, 5);
  }
  return ami_prim(R, c, tables, slot);
}

#endif
