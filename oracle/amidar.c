/* amidar.c -- CPU ORACLE (test infrastructure, not product code).
 *
 * Restates the Amidar engine behind Toybox('amidar') (reference call sites
 * toybox/envs/atari/base.py:126,109; state schema toybox/interventions/amidar.py:22-34 Amidar,
 * :83-166 MovementAI (5 protocols), :171 Enemy, :195 Player, :216 Board, :300 Box, :316 TilePoint;
 * queries tile_to_world/world_to_tile amidar.py:508-518; constants and initial state from
 * toybox/interventions/defaults/amidar_*_default.json).
 *
 * Pinned by the fixtures ([FIX]): board 32x31 + tags, 29 boxes (4 chase corners), 60 junctions,
 * player start (31,15)/history [607], enemy start tiles 0,0,7,800,969, tile<->world scale (64,80),
 * RNG lineage (child of the un-advanced config rand), jumps 4 -> 3 after one FIRE.
 * The enemy route tables are NOT in the reference (compiled into ctoybox); the five loops below only
 * honour the pinned start tiles.  Frame rules A1..A7: PARITY UNPINNED restatement.  All i32.
 */
#include "tbo.h"
#include <string.h>

#define TW 64        /* tile width  in world units (4 px * 16) */
#define TH 80        /* tile height in world units (5 px * 16) */
#define OFF_X 16
#define OFF_Y 37

static const char *DEFAULT_BOARD[TBO_AMI_BH] = {
  "c========================c======",
  "=     =   =   =  =   =   =     =",
  "=     =   =   =  =   =   =     =",
  "=     =   =   =  =   =   =     =",
  "=     =   =   =  =   =   =     =",
  "=     =   =   =  =   =   =     =",
  "================================",
  "=   =    =  =      =  =    =   =",
  "=   =    =  =      =  =    =   =",
  "=   =    =  =      =  =    =   =",
  "=   =    =  =      =  =    =   =",
  "=   =    =  =      =  =    =   =",
  "================================",
  "=  =       =        =       =  p",
  "=  =       =        =       =  p",
  "=  =       =        =       =  p",
  "=  =       =        =       =  p",
  "=  =       =        =       =  p",
  "===============================p",
  "=    =        =  =        =    =",
  "=    =        =  =        =    =",
  "=    =        =  =        =    =",
  "=    =        =  =        =    =",
  "=    =        =  =        =    =",
  "c========================c======",
  "=     =     =      =     =     =",
  "=     =     =      =     =     =",
  "=     =     =      =     =     =",
  "=     =     =      =     =     =",
  "=     =     =      =     =     =",
  "================================" };

/* closed loops over full-width rows / full-height side columns; first entries are the pinned starts */
static const int32_t DEFAULT_ROUTES[5][5] = {
  { 0, 31, 223, 192, -1 },          /* (0,0) -> (31,0) -> (31,6) -> (0,6) */
  { 0, 384, 415, 31, -1 },          /* (0,0) -> (0,12) -> (31,12) -> (31,0) */
  { 7, 31, 415, 384, 0 },           /* (7,0) -> (31,0) -> (31,12) -> (0,12) -> (0,0) */
  { 800, 960, 991, 799, 768 },      /* (0,25) -> (0,30) -> (31,30) -> (31,24) -> (0,24) */
  { 969, 991, 607, 576, 960 } };    /* (9,30) -> (31,30) -> (31,18) -> (0,18) -> (0,30) */

void tbo_ami_tile_to_world(int tx, int ty, int *wx, int *wy) { *wx = tx * TW; *wy = ty * TH; }
static int floordiv(int a, int b) { int q = a / b; if ((a % b) != 0 && ((a < 0) != (b < 0))) q--; return q; }
void tbo_ami_world_to_tile(int wx, int wy, int *tx, int *ty) { *tx = floordiv(wx, TW); *ty = floordiv(wy, TH); }

void tbo_ami_default_cfg(tbo_ami_cfg *c) {
  memset(c, 0, sizeof *c);
  c->bg_color = (tbo_color){0, 0, 0, 255};
  c->player_color = (tbo_color){255, 255, 153, 255};
  c->unpainted_color = (tbo_color){148, 0, 211, 255};
  c->painted_color = (tbo_color){255, 255, 30, 255};
  c->enemy_color = (tbo_color){255, 50, 100, 255};
  c->inner_painted_color = (tbo_color){255, 255, 0, 255};
  c->start_lives = 3; c->start_jumps = 4; c->chase_time = 300; c->chase_score_bonus = 100;
  c->jump_time = 75; c->box_bonus = 50; c->render_images = 1; c->default_board_bugs = 1;
  c->player_start_tx = 31; c->player_start_ty = 15;
  for (int y = 0; y < TBO_AMI_BH; y++)
    for (int x = 0; x < TBO_AMI_BW; x++) {
      char ch = DEFAULT_BOARD[y][x];
      c->board[y][x] = ch == '=' ? TBO_TILE_UNPAINTED : ch == 'p' ? TBO_TILE_PAINTED : ch == 'c' ? TBO_TILE_CHASE : TBO_TILE_EMPTY;
    }
  c->n_enemies = 5;
  for (int i = 0; i < 5; i++) { c->enemies[i].kind = TBO_AI_LOOKUP; c->enemies[i].next = 0; c->enemies[i].default_route_index = i; }
  c->n_routes = 5;
  for (int i = 0; i < 5; i++) {
    int n = 0;
    while (n < 5 && DEFAULT_ROUTES[i][n] >= 0) { c->routes[i][n] = DEFAULT_ROUTES[i][n]; n++; }
    c->route_len[i] = n;
  }
  tbo_rng_seed(&c->rand, 13);
}

static int walkable(const tbo_ami_state *s, int tx, int ty) {
  return tx >= 0 && tx < TBO_AMI_BW && ty >= 0 && ty < TBO_AMI_BH && s->tiles[ty][tx] != TBO_TILE_EMPTY;
}
/* a junction is a walkable tile that has both a horizontal and a vertical walkable neighbour */
static int junction_tile(const tbo_ami_state *s, int tx, int ty) {
  if (!walkable(s, tx, ty)) return 0;
  int h = walkable(s, tx - 1, ty) || walkable(s, tx + 1, ty);
  int v = walkable(s, tx, ty - 1) || walkable(s, tx, ty + 1);
  return h && v;
}
static int is_junction(const tbo_ami_state *s, int id) {
  for (int i = 0; i < s->n_junctions; i++) if (s->junctions[i] == id) return 1;
  return 0;
}
static const int DX[4] = { 0, 0, -1, 1 }, DY[4] = { -1, 1, 0, 0 };   /* Up Down Left Right */
static int opposite(int d) { return d ^ 1; }

static void enemy_start_tile(const tbo_ami_cfg *c, const tbo_ami_ai *ai, int *tx, int *ty) {
  if (ai->kind == TBO_AI_LOOKUP) {
    int id = 0;
    if (c->n_routes > 0) {
      int r = ai->default_route_index % c->n_routes; if (r < 0) r += c->n_routes;
      if (c->route_len[r] > 0) id = c->routes[r][0];
    }
    *tx = id % TBO_AMI_BW; *ty = id / TBO_AMI_BW;
  } else { *tx = ai->start_tx; *ty = ai->start_ty; }
}
static void reset_ai(tbo_ami_ai *ai) {
  switch (ai->kind) {
    case TBO_AI_LOOKUP: ai->next = 0; break;
    case TBO_AI_AMIDAR: ai->vert = ai->start_vert; ai->horiz = ai->start_horiz; break;
    case TBO_AI_TARGET: ai->dir = ai->start_dir; ai->has_seen = 0; ai->seen_tx = 0; ai->seen_ty = 0; break;
    case TBO_AI_RANDOM: ai->dir = ai->start_dir; break;
    default: break;
  }
}
static void place_mob(tbo_ami_mob *m, int tx, int ty) {
  m->x = tx * TW; m->y = ty * TH; m->has_step = 0; m->step_tx = 0; m->step_ty = 0; m->caught = 0;
}
static void reset_player(const tbo_ami_cfg *c, tbo_ami_state *s) {
  place_mob(&s->player, c->player_start_tx, c->player_start_ty);
  s->player.n_history = 0;
  /* [FIX] fixture history [607]: the first junction below the start tile */
  for (int ty = c->player_start_ty; ty < TBO_AMI_BH && walkable(s, c->player_start_tx, ty); ty++)
    if (is_junction(s, ty * TBO_AMI_BW + c->player_start_tx)) {
      s->player.history[0] = ty * TBO_AMI_BW + c->player_start_tx; s->player.n_history = 1; break;
    }
}
static void reset_enemy(const tbo_ami_cfg *c, tbo_ami_mob *e) {
  int tx, ty;
  reset_ai(&e->ai);
  enemy_start_tile(c, &e->ai, &tx, &ty);
  place_mob(e, tx, ty);
  e->n_history = 0;
}
static void reset_board(const tbo_ami_cfg *c, tbo_ami_state *s) {
  memcpy(s->tiles, c->board, sizeof s->tiles);
  for (int i = 0; i < s->n_boxes; i++) s->boxes[i].painted = 0;
}

/* boxes: maximal empty rectangles; the board derives them, junctions and chase junctions at new_game */
static void derive_board(const tbo_ami_cfg *c, tbo_ami_state *s) {
  memcpy(s->tiles, c->board, sizeof s->tiles);
  s->n_junctions = 0; s->n_chase_junctions = 0; s->n_boxes = 0;
  for (int ty = 0; ty < TBO_AMI_BH; ty++)
    for (int tx = 0; tx < TBO_AMI_BW; tx++) {
      if (junction_tile(s, tx, ty) && s->n_junctions < TBO_AMI_MAX_JUNCTIONS) s->junctions[s->n_junctions++] = ty * TBO_AMI_BW + tx;
      if (s->tiles[ty][tx] == TBO_TILE_CHASE && s->n_chase_junctions < 4) s->chase_junctions[s->n_chase_junctions++] = ty * TBO_AMI_BW + tx;
    }
  /* a box's top-left is a walkable tile whose right and lower neighbours are walkable and whose
   * diagonal (tx+1,ty+1) is empty; extend right/down along the empty interior */
  for (int ty = 0; ty + 1 < TBO_AMI_BH; ty++)
    for (int tx = 0; tx + 1 < TBO_AMI_BW; tx++) {
      if (!(walkable(s, tx, ty) && walkable(s, tx + 1, ty) && walkable(s, tx, ty + 1) && !walkable(s, tx + 1, ty + 1))) continue;
      int bx = tx + 1, by = ty + 1;
      while (bx < TBO_AMI_BW && !walkable(s, bx, ty + 1)) bx++;
      while (by < TBO_AMI_BH && !walkable(s, tx + 1, by)) by++;
      if (bx >= TBO_AMI_BW || by >= TBO_AMI_BH || s->n_boxes >= TBO_AMI_MAX_BOXES) continue;
      tbo_ami_box *b = &s->boxes[s->n_boxes++];
      b->tl_tx = tx; b->tl_ty = ty; b->br_tx = bx; b->br_ty = by; b->painted = 0;
      b->triggers_chase = s->tiles[ty][tx] == TBO_TILE_CHASE;
    }
}

void tbo_ami_new_game(tbo_ami_cfg *c, tbo_ami_state *s) {
  tbo_rng copy = c->rand;
  memset(s, 0, sizeof *s);
  s->rand = tbo_rng_child(&copy);                 /* [FIX] App. A.2: config rand is not advanced */
  s->score = 0; s->lives = c->start_lives; s->level = 1; s->jumps = c->start_jumps;
  derive_board(c, s);
  s->player.speed = 8; s->player.ai.kind = TBO_AI_PLAYER;
  reset_player(c, s);
  s->n_enemies = c->n_enemies;
  for (int i = 0; i < s->n_enemies; i++) {
    s->enemies[i].ai = c->enemies[i]; s->enemies[i].speed = 8;
    reset_enemy(c, &s->enemies[i]);
  }
}

/* advance a mob toward its step tile; returns 1 on arrival */
static int advance(tbo_ami_mob *m) {
  int gx = m->step_tx * TW, gy = m->step_ty * TH;
  int dx = gx - m->x, dy = gy - m->y, sp = m->speed < 0 ? 0 : m->speed;
  if (dx > sp) dx = sp; if (dx < -sp) dx = -sp;
  if (dy > sp) dy = sp; if (dy < -sp) dy = -sp;
  m->x += dx; m->y += dy;
  if (m->x == gx && m->y == gy) { m->has_step = 0; return 1; }
  return 0;
}
static int aligned(const tbo_ami_mob *m) { return (m->x % TW) == 0 && (m->y % TH) == 0; }

static void check_boxes(const tbo_ami_cfg *c, tbo_ami_state *s) {
  int newly_chase = 0;
  for (int i = 0; i < s->n_boxes; i++) {
    tbo_ami_box *b = &s->boxes[i];
    if (b->painted) continue;
    int ok = 1;
    for (int tx = b->tl_tx; tx <= b->br_tx && ok; tx++)
      ok = s->tiles[b->tl_ty][tx] == TBO_TILE_PAINTED && s->tiles[b->br_ty][tx] == TBO_TILE_PAINTED;
    for (int ty = b->tl_ty; ty <= b->br_ty && ok; ty++)
      ok = s->tiles[ty][b->tl_tx] == TBO_TILE_PAINTED && s->tiles[ty][b->br_tx] == TBO_TILE_PAINTED;
    if (!ok) continue;
    b->painted = 1; s->score += c->box_bonus;
    if (b->triggers_chase) newly_chase = 1;
  }
  if (newly_chase) {
    int all = 1;
    for (int i = 0; i < s->n_boxes; i++) if (s->boxes[i].triggers_chase && !s->boxes[i].painted) all = 0;
    if (all) s->chase_timer = c->chase_time;
  }
}

static void player_arrived(const tbo_ami_cfg *c, tbo_ami_state *s, int tx, int ty) {
  tbo_ami_mob *p = &s->player;
  int id = ty * TBO_AMI_BW + tx;
  if (!is_junction(s, id)) return;
  if (p->n_history > 0 && p->history[0] != id) {
    int px = p->history[0] % TBO_AMI_BW, py = p->history[0] / TBO_AMI_BW, ok = (px == tx || py == ty);
    int sx = px < tx ? 1 : px > tx ? -1 : 0, sy = py < ty ? 1 : py > ty ? -1 : 0;
    if (ok) for (int x = px, y = py; ; x += sx, y += sy) { if (!walkable(s, x, y)) ok = 0; if (x == tx && y == ty) break; }
    if (ok) {
      int newly = 0;
      for (int x = px, y = py; ; x += sx, y += sy) {
        if (s->tiles[y][x] != TBO_TILE_PAINTED) { s->tiles[y][x] = TBO_TILE_PAINTED; newly++; }
        if (x == tx && y == ty) break;
      }
      s->score += newly;
      if (newly > 0) check_boxes(c, s);
    }
  }
  if (p->n_history == 0 || p->history[0] != id) {
    int n = p->n_history < TBO_AMI_HIST ? p->n_history : TBO_AMI_HIST - 1;
    for (int i = n; i > 0; i--) p->history[i] = p->history[i - 1];
    p->history[0] = id; p->n_history = n + 1;
  }
}

/* head for tile (gx,gy): along the shared row/column when there is one, else x first then y */
static int toward(const tbo_ami_state *s, int tx, int ty, int gx, int gy) {
  int dh = gx > tx ? TBO_DIR_RIGHT : TBO_DIR_LEFT, dv = gy > ty ? TBO_DIR_DOWN : TBO_DIR_UP;
  if (gy == ty && gx != tx) return walkable(s, tx + DX[dh], ty) ? dh : -1;
  if (gx == tx && gy != ty) return walkable(s, tx, ty + DY[dv]) ? dv : -1;
  if (gx != tx && walkable(s, tx + DX[dh], ty)) return dh;
  if (gy != ty && walkable(s, tx, ty + DY[dv])) return dv;
  return -1;
}
/* options = walkable neighbours in Up,Down,Left,Right order, without the reverse of `dir` unless dead end */
static int options(const tbo_ami_state *s, int tx, int ty, int dir, int *out) {
  int n = 0;
  for (int d = 0; d < 4; d++) if (d != opposite(dir) && walkable(s, tx + DX[d], ty + DY[d])) out[n++] = d;
  if (n == 0 && walkable(s, tx + DX[opposite(dir)], ty + DY[opposite(dir)])) out[n++] = opposite(dir);
  return n;
}
static int random_dir(tbo_ami_state *s, int tx, int ty, int dir) {
  int opt[4], n = options(s, tx, ty, dir, opt);
  if (n == 0) return -1;
  if (!junction_tile(s, tx, ty) || n == 1) {
    for (int i = 0; i < n; i++) if (opt[i] == dir) return dir;
    return opt[0];
  }
  return opt[tbo_rng_index(&s->rand, (uint32_t)n)];
}

/* A4: pick the next tile for enemy e standing (aligned) on (tx,ty); returns direction or -1 */
static int enemy_choose(const tbo_ami_cfg *c, tbo_ami_state *s, tbo_ami_mob *e, int tx, int ty) {
  tbo_ami_ai *ai = &e->ai;
  switch (ai->kind) {
    case TBO_AI_LOOKUP: {
      if (c->n_routes <= 0) return -1;
      int r = ai->default_route_index % c->n_routes; if (r < 0) r += c->n_routes;
      int len = c->route_len[r];
      if (len <= 0) return -1;
      int nx = ai->next % len; if (nx < 0) nx += len;
      if (c->routes[r][nx] == ty * TBO_AMI_BW + tx) nx = (nx + 1) % len;
      ai->next = nx;
      return toward(s, tx, ty, c->routes[r][nx] % TBO_AMI_BW, c->routes[r][nx] / TBO_AMI_BW);
    }
    case TBO_AI_PERIMETER: {
      int d = -1;
      if (ty == 0 && tx < TBO_AMI_BW - 1) d = TBO_DIR_RIGHT;
      else if (tx == TBO_AMI_BW - 1 && ty < TBO_AMI_BH - 1) d = TBO_DIR_DOWN;
      else if (ty == TBO_AMI_BH - 1 && tx > 0) d = TBO_DIR_LEFT;
      else if (tx == 0 && ty > 0) d = TBO_DIR_UP;
      if (d >= 0 && walkable(s, tx + DX[d], ty + DY[d])) return d;
      return toward(s, tx, ty, ai->start_tx, ai->start_ty);
    }
    case TBO_AI_AMIDAR: {
      if (ai->vert != TBO_DIR_UP && ai->vert != TBO_DIR_DOWN) ai->vert = TBO_DIR_DOWN;
      if (ai->horiz != TBO_DIR_LEFT && ai->horiz != TBO_DIR_RIGHT) ai->horiz = TBO_DIR_RIGHT;
      if (walkable(s, tx, ty + DY[ai->vert])) return ai->vert;
      if (ty == 0 || ty == TBO_AMI_BH - 1) ai->vert = opposite(ai->vert);
      if (walkable(s, tx + DX[ai->horiz], ty)) return ai->horiz;
      ai->horiz = opposite(ai->horiz);
      if (walkable(s, tx + DX[ai->horiz], ty)) return ai->horiz;
      ai->vert = opposite(ai->vert);
      if (walkable(s, tx, ty + DY[ai->vert])) return ai->vert;
      return -1;
    }
    case TBO_AI_TARGET: {
      int ptx, pty, d;
      tbo_ami_world_to_tile(s->player.x, s->player.y, &ptx, &pty);
      int dist = (ptx > tx ? ptx - tx : tx - ptx) + (pty > ty ? pty - ty : ty - pty);
      if (ai->dir < 0 || ai->dir > 3) ai->dir = TBO_DIR_UP;
      if (dist <= ai->vision_distance) { ai->has_seen = 1; ai->seen_tx = ptx; ai->seen_ty = pty; }
      if (ai->has_seen && ai->seen_tx == tx && ai->seen_ty == ty) { ai->has_seen = 0; ai->seen_tx = 0; ai->seen_ty = 0; }
      if (ai->has_seen) {
        int opt[4], n = options(s, tx, ty, ai->dir, opt), best = 0x7fffffff; d = -1;
        for (int i = 0; i < n; i++) {
          int nx = tx + DX[opt[i]], ny = ty + DY[opt[i]];
          int m = (ai->seen_tx > nx ? ai->seen_tx - nx : nx - ai->seen_tx) + (ai->seen_ty > ny ? ai->seen_ty - ny : ny - ai->seen_ty);
          if (m < best) { best = m; d = opt[i]; }
        }
      } else d = random_dir(s, tx, ty, ai->dir);
      if (d >= 0) ai->dir = d;
      return d;
    }
    case TBO_AI_RANDOM: {
      if (ai->dir < 0 || ai->dir > 3) ai->dir = TBO_DIR_UP;
      int d = random_dir(s, tx, ty, ai->dir);
      if (d >= 0) ai->dir = d;
      return d;
    }
    default: return -1;
  }
}

static void snap_step(tbo_ami_mob *m) {      /* off-grid mob (after an intervention): walk to the containing tile */
  int tx, ty; tbo_ami_world_to_tile(m->x, m->y, &tx, &ty);
  if (tx < 0) tx = 0; if (tx > TBO_AMI_BW - 1) tx = TBO_AMI_BW - 1;
  if (ty < 0) ty = 0; if (ty > TBO_AMI_BH - 1) ty = TBO_AMI_BH - 1;
  m->has_step = 1; m->step_tx = tx; m->step_ty = ty;
}

void tbo_ami_step(const tbo_ami_cfg *c, tbo_ami_state *s, int in) {
  if (s->lives <= 0) return;
  /* A1 timers */
  if (s->jump_timer > 0) s->jump_timer -= 1;
  if (s->chase_timer > 0) {
    s->chase_timer -= 1;
    if (s->chase_timer == 0)
      for (int i = 0; i < s->n_enemies; i++) if (s->enemies[i].caught) reset_enemy(c, &s->enemies[i]);
  }
  /* A2 jump */
  if ((in & TBO_IN_BUTTON1) && s->jumps > 0 && s->jump_timer == 0) { s->jumps -= 1; s->jump_timer = c->jump_time; }
  /* A3 player */
  {
    tbo_ami_mob *p = &s->player;
    if (!p->has_step) {
      if (!aligned(p)) snap_step(p);
      else {
        int tx = p->x / TW, ty = p->y / TH;
        static const int order[4] = { TBO_DIR_UP, TBO_DIR_DOWN, TBO_DIR_LEFT, TBO_DIR_RIGHT };
        static const int bit[4] = { TBO_IN_UP, TBO_IN_DOWN, TBO_IN_LEFT, TBO_IN_RIGHT };
        for (int k = 0; k < 4; k++) {
          int d = order[k];
          if ((in & bit[k]) && walkable(s, tx + DX[d], ty + DY[d])) { p->has_step = 1; p->step_tx = tx + DX[d]; p->step_ty = ty + DY[d]; break; }
        }
      }
    }
    if (p->has_step) { int tx = p->step_tx, ty = p->step_ty; if (advance(p)) player_arrived(c, s, tx, ty); }
  }
  /* A4 enemies */
  for (int i = 0; i < s->n_enemies; i++) {
    tbo_ami_mob *e = &s->enemies[i];
    if (e->caught) continue;
    if (!e->has_step) {
      if (!aligned(e)) snap_step(e);
      else {
        int tx = e->x / TW, ty = e->y / TH;
        int d = enemy_choose(c, s, e, tx, ty);
        if (d >= 0) { e->has_step = 1; e->step_tx = tx + DX[d]; e->step_ty = ty + DY[d]; }
      }
    }
    if (e->has_step) advance(e);
  }
  /* A5 contact */
  for (int i = 0; i < s->n_enemies; i++) {
    tbo_ami_mob *e = &s->enemies[i];
    if (e->caught) continue;
    int dx = e->x - s->player.x, dy = e->y - s->player.y;
    if (dx < 0) dx = -dx; if (dy < 0) dy = -dy;
    if (dx > TW / 2 || dy > TH / 2) continue;
    if (s->jump_timer > 0) continue;
    if (s->chase_timer > 0) { e->caught = 1; s->score += c->chase_score_bonus; continue; }
    s->lives -= 1; s->jump_timer = 0;
    reset_player(c, s);
    for (int k = 0; k < s->n_enemies; k++) reset_enemy(c, &s->enemies[k]);
    break;
  }
  /* A6 level complete: every box painted */
  {
    int all = s->n_boxes > 0;
    for (int i = 0; i < s->n_boxes; i++) if (!s->boxes[i].painted) all = 0;
    if (all) {
      s->level += 1; reset_board(c, s); s->jumps = c->start_jumps; s->jump_timer = 0; s->chase_timer = 0;
      reset_player(c, s);
      for (int k = 0; k < s->n_enemies; k++) reset_enemy(c, &s->enemies[k]);
    }
  }
}

/* A7 draw list */
void tbo_ami_render(const tbo_ami_cfg *c, const tbo_ami_state *s, uint8_t *rgba) {
  tbo_canvas cv = { TBO_AMI_W, TBO_AMI_H, rgba };
  tbo_clear(&cv, c->bg_color);
  for (int ty = 0; ty < TBO_AMI_BH; ty++)
    for (int tx = 0; tx < TBO_AMI_BW; tx++) {
      uint8_t t = s->tiles[ty][tx];
      if (t == TBO_TILE_EMPTY) continue;
      tbo_rect(&cv, t == TBO_TILE_PAINTED ? c->painted_color : c->unpainted_color, OFF_X + 4 * tx, OFF_Y + 5 * ty, 4, 5);
    }
  for (int i = 0; i < s->n_boxes; i++) {
    const tbo_ami_box *b = &s->boxes[i];
    if (!b->painted) continue;
    tbo_rect(&cv, c->inner_painted_color, OFF_X + 4 * (b->tl_tx + 1), OFF_Y + 5 * (b->tl_ty + 1),
             4 * (b->br_tx - b->tl_tx - 1), 5 * (b->br_ty - b->tl_ty - 1));
  }
  for (int i = 0; i < s->n_enemies; i++) {
    const tbo_ami_mob *e = &s->enemies[i];
    if (e->caught) continue;
    tbo_rect(&cv, c->enemy_color, OFF_X + floordiv(e->x, 16) - 1, OFF_Y + floordiv(e->y, 16) - 1, 6, 7);
  }
  tbo_rect(&cv, c->player_color, OFF_X + floordiv(s->player.x, 16) - 1, OFF_Y + floordiv(s->player.y, 16) - 1, 6, 7);
  tbo_digits(&cv, c->player_color, 100, 205, s->score, 2, 2);
  tbo_digits(&cv, c->player_color, 124, 205, s->lives, 2, 2);
  tbo_digits(&cv, c->painted_color, 144, 205, s->jumps, 2, 2);
}
