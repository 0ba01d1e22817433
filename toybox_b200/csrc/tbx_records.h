/* tbx_records.h -- per-env state records, pool configs and shared geometry tables.
 *
 * One record = every mutable field of one environment, laid out as 32-bit words (f64/u64 fields are
 * 8-byte aligned and occupy two consecutive words).  On the device a pool stores the TRANSPOSE of an
 * array of records: plane[w * n_pad + env] (word-major, env-minor), so a thread-per-env kernel reads
 * word w of 32 consecutive envs as one 128-byte line.  The same struct is the AoS form used by the
 * gather/scatter kernels behind JSON import/export.
 *
 * Field lists follow the reference's JSON state schema:
 *   Breakout       toybox/interventions/breakout.py:49-68 (:132 Paddle, :198 Brick, :276 Ball)
 *   Space Invaders toybox/interventions/space_invaders.py:16-32 (:38 Player, :60 Laser, :101 Ufo, :116 Enemy, :146)
 *   Amidar         toybox/interventions/amidar.py:22-34 (:83-166 MovementAI, :171 Enemy, :195 Player, :216 Board, :300 Box)
 * Per-brick / per-box geometry that new_game derives from the config lives in shared tables
 * (BrkTable / AmiTable); an env carries only an index into them (hdr.tbl) plus its dynamic bits.
 */
#ifndef TBX_RECORDS_H
#define TBX_RECORDS_H
#include <stdint.h>
#include <stddef.h>

enum { TBX_BREAKOUT = 0, TBX_AMIDAR = 1, TBX_SPACE_INVADERS = 2 };
enum { TBX_IN_LEFT = 1, TBX_IN_RIGHT = 2, TBX_IN_UP = 4, TBX_IN_DOWN = 8, TBX_IN_BUTTON1 = 16, TBX_IN_BUTTON2 = 32 };
enum { TBX_DIR_UP = 0, TBX_DIR_DOWN = 1, TBX_DIR_LEFT = 2, TBX_DIR_RIGHT = 3 };
#define TBX_NONE (-2147483647 - 1) /* Option<i32>::None */

/* words 0..15 of every record */
typedef struct {
  uint64_t rand[2];     /* state.rand (in-episode stream) */
  uint64_t sim_rand[2]; /* simulator/config rand of this env (episode seeding), SURVEY App. A.2 */
  int32_t lives, score, level;
  int32_t prev_score; /* toybox/envs/atari/base.py:136-138 reward bookkeeping */
  int32_t ep_len, ep_return;
  int32_t tbl; /* index into the pool's geometry tables */
  int32_t _pad;
} TbxHdr;
#define TBX_HDR_WORDS 16

/* ------------------------------------------------------------------ Breakout (f64) */
#define TBX_BRK_W 240
#define TBX_BRK_H 160
#define TBX_BRK_MAX_BRICKS 144
#define TBX_BRK_MAX_BALLS 4
#define TBX_BRK_MAX_ROWS 8
#define TBX_BRK_MAX_STARTS 8
#define TBX_BRK_MAX_SEGS 16

typedef struct {
  TbxHdr hdr;
  double paddle_px, paddle_py, paddle_vx, paddle_vy;
  double ball[TBX_BRK_MAX_BALLS][4]; /* px, py, vx, vy */
  double paddle_width, paddle_speed, ball_radius;
  int32_t n_balls, is_dead, reset, _pad;
  uint32_t alive[5]; /* bit i = bricks[i].alive */
  int32_t _pad2;
} BrkRec;

typedef struct {
  int32_t n_bricks, disjoint; /* disjoint: no two brick rects overlap -> they may be painted in any order */
  int32_t hud_clear;          /* no brick rect reaches into the HUD digit rows (y < 12) */
  int32_t delta_ok;           /* disjoint, HUD-clear and every brick lies on pure background: base frame 1 may hold them */
  double bb_x0, bb_y0, bb_x1, bb_y1; /* union of all brick boxes (collision early-out) */
  /* the bricks form a regular column-major grid (index = col * g_nrows + row, brick (col, row) at (gx0 + col * gw, gy0 + row * gh)):
   * the collision test visits only the few cells around the ball instead of every alive brick (same first hit in index order) */
  int32_t grid, g_ncols, g_nrows, _padg;
  double gx0, gy0, ginv_w, ginv_h;
  double px[TBX_BRK_MAX_BRICKS], py[TBX_BRK_MAX_BRICKS], sx[TBX_BRK_MAX_BRICKS], sy[TBX_BRK_MAX_BRICKS];
  double x1[TBX_BRK_MAX_BRICKS], y1[TBX_BRK_MAX_BRICKS]; /* px+sx, py+sy (same IEEE add the step does) */
  int32_t points[TBX_BRK_MAX_BRICKS], depth[TBX_BRK_MAX_BRICKS], row[TBX_BRK_MAX_BRICKS], col[TBX_BRK_MAX_BRICKS];
  uint32_t color[TBX_BRK_MAX_BRICKS]; /* r | g<<8 | b<<16 | a<<24 */
  int32_t ix[TBX_BRK_MAX_BRICKS], iy[TBX_BRK_MAX_BRICKS], iw[TBX_BRK_MAX_BRICKS], ih[TBX_BRK_MAX_BRICKS]; /* `as i32` of the f64s */
  uint32_t destructible[5];
  uint32_t all_mask[5]; /* low n_bricks bits set */
} BrkTable;

typedef struct {
  uint32_t bg_color, frame_color, paddle_color, ball_color;
  int32_t n_rows;
  uint32_t row_colors[TBX_BRK_MAX_ROWS];
  int32_t row_scores[TBX_BRK_MAX_ROWS];
  int32_t start_lives, paddle_discrete_segments, ball_speed_row_depth;
  double ball_speed_slow, ball_speed_fast;
  int32_t n_starts, _pad;
  double start_x[TBX_BRK_MAX_STARTS], start_y[TBX_BRK_MAX_STARTS], start_angle[TBX_BRK_MAX_STARTS];
  /* host-evaluated trig (glibc), so host oracle and device use identical bits (SURVEY 7 hard parts) */
  double start_cos[TBX_BRK_MAX_STARTS], start_sin[TBX_BRK_MAX_STARTS];
  double seg_cos[TBX_BRK_MAX_SEGS], seg_sin[TBX_BRK_MAX_SEGS];
  uint64_t rand[2];
  int32_t default_tbl, _pad2;
} BrkCfg;

/* ------------------------------------------------------------------ Space Invaders (i32) */
#define TBX_SI_W 320
#define TBX_SI_H 210
#define TBX_SI_N_ENEMIES 36
#define TBX_SI_MAX_LASERS 4
#define TBX_SI_N_SHIELDS 3
#define TBX_SI_SHIELD_W 16
#define TBX_SI_SHIELD_H 18

typedef struct { int32_t x, y, w, h, t, movement, speed; uint32_t color; } SiLaser;
typedef struct {
  TbxHdr hdr;
  int32_t ship_x, ship_y, ship_w, ship_h, ship_speed, ship_death_counter, ship_alive, ship_death_hit_1;
  uint32_t ship_color;
  int32_t has_ship_laser;
  SiLaser ship_laser;
  int32_t n_enemy_lasers, _pad0;
  SiLaser enemy_lasers[TBX_SI_MAX_LASERS];
  int32_t move_counter, move_dir, visual_orientation;
  int32_t ufo_x, ufo_y, ufo_appearance_counter, ufo_death_counter;
  int32_t life_display_timer, enemy_shot_delay, _pad1;
  int32_t shield_x[TBX_SI_N_SHIELDS], shield_y[TBX_SI_N_SHIELDS];
  uint32_t shield_rows[TBX_SI_N_SHIELDS][TBX_SI_SHIELD_H]; /* bit (15-c) of row r = opaque */
  int32_t en_x[TBX_SI_N_ENEMIES], en_y[TBX_SI_N_ENEMIES], en_death[TBX_SI_N_ENEMIES];
  int32_t en_row[TBX_SI_N_ENEMIES], en_col[TBX_SI_N_ENEMIES], en_id[TBX_SI_N_ENEMIES], en_points[TBX_SI_N_ENEMIES];
  uint32_t en_alive[2]; /* bit i = enemies[i].alive */
} SiRec;

typedef struct {
  double jitter;
  int32_t enemy_protocol, start_lives;
  int32_t shields[TBX_SI_N_SHIELDS][2];
  int32_t row_scores[6];
  uint64_t rand[2];
} SiCfg;

/* ------------------------------------------------------------------ Amidar (i32) */
#define TBX_AMI_W 160
#define TBX_AMI_H 250
#define TBX_AMI_BW 32
#define TBX_AMI_BH 31
#define TBX_AMI_MAX_ENEMIES 8
#define TBX_AMI_MAX_BOXES 32
#define TBX_AMI_MAX_JUNCTIONS 64
#define TBX_AMI_HIST 8
#define TBX_AMI_MAX_ROUTES 16
#define TBX_AMI_MAX_ROUTE_LEN 64
enum { TBX_TILE_EMPTY = 0, TBX_TILE_UNPAINTED = 1, TBX_TILE_CHASE = 2, TBX_TILE_PAINTED = 3 };
enum { TBX_AI_PLAYER = 0, TBX_AI_LOOKUP = 1, TBX_AI_PERIMETER = 2, TBX_AI_AMIDAR = 3, TBX_AI_TARGET = 4, TBX_AI_RANDOM = 5 };

typedef struct {
  int32_t kind, next, default_route_index, start_tx, start_ty, vert, horiz, start_vert, start_horiz;
  int32_t start_dir, dir, vision_distance, seen_tx, seen_ty, has_seen;
} AmiAi; /* 15 words */
typedef struct {
  int32_t x, y, has_step, step_tx, step_ty, n_history;
  int32_t history[TBX_AMI_HIST];
  int32_t caught, speed;
  AmiAi ai;
} AmiMob; /* 31 words */
typedef struct {
  TbxHdr hdr;
  int32_t jumps, jump_timer, chase_timer, n_enemies;
  uint32_t box_painted; /* bit i = boxes[i].painted */
  int32_t _pad[3];
  uint32_t tiles[TBX_AMI_BH][2]; /* 2 bits per tile, tile tx at bits (2*tx) of the 64-bit row */
  AmiMob player;
  AmiMob enemies[TBX_AMI_MAX_ENEMIES];
  int32_t _pad2;
} AmiRec;

typedef struct {
  int32_t n_boxes, n_junctions, n_chase_junctions, _pad;
  int32_t tl_tx[TBX_AMI_MAX_BOXES], tl_ty[TBX_AMI_MAX_BOXES], br_tx[TBX_AMI_MAX_BOXES], br_ty[TBX_AMI_MAX_BOXES];
  uint32_t triggers_chase; /* bit i */
  uint32_t all_boxes;      /* low n_boxes bits */
  int32_t junctions[TBX_AMI_MAX_JUNCTIONS];
  int32_t chase_junctions[4];
  uint32_t junction_bits[TBX_AMI_BH]; /* bit tx of row ty: id ty*32+tx is in `junctions` (ids in range only) */
} AmiTable;

typedef struct {
  uint32_t bg_color, player_color, unpainted_color, painted_color, enemy_color, inner_painted_color;
  int32_t start_lives, start_jumps, chase_time, chase_score_bonus, jump_time, box_bonus;
  int32_t render_images, default_board_bugs, player_start_tx, player_start_ty;
  uint32_t board[TBX_AMI_BH][2]; /* packed like AmiRec.tiles */
  int32_t n_enemies, n_routes;
  AmiAi enemies[TBX_AMI_MAX_ENEMIES];
  int32_t route_len[TBX_AMI_MAX_ROUTES];
  int32_t routes[TBX_AMI_MAX_ROUTES][TBX_AMI_MAX_ROUTE_LEN];
  uint64_t rand[2];
  int32_t default_tbl, _pad;
} AmiCfg;

/* ------------------------------------------------------------------ INTER_AREA tap tables (host-built) */
#define TBX_RS_MAX_DST 256
#define TBX_RS_MAX_TAPS 768
typedef struct {
  int32_t ssize, dsize, ntaps, max_taps;
  uint16_t start[TBX_RS_MAX_DST + 1]; /* taps of destination index d are [start[d], start[d+1]) */
  uint16_t si[TBX_RS_MAX_TAPS];
  float alpha[TBX_RS_MAX_TAPS];
} TbxResizeAxis;
typedef struct { TbxResizeAxis x, y; } TbxResizeTab;

/* The same taps repacked for the fused render kernel: every destination index has exactly `tx` / `ty` taps
 * on consecutive source indices (weights zero-padded; x + 0*b == x exactly for the non-negative sums here),
 * plus the inverse maps "which destination indices does source index s feed". */
#define TBX_AREA_MAX_DST 128
#define TBX_AREA_MAX_TAPS 8
#define TBX_AREA_MAX_SRC 320
typedef struct {
  int32_t sw, sh, dw, dh, tx, ty, _pad[2];
  uint16_t xs0[TBX_AREA_MAX_DST], ys0[TBX_AREA_MAX_DST];  /* first source index of each destination index */
  uint16_t yn[TBX_AREA_MAX_DST];                          /* real tap count per destination row */
  float xalpha[TBX_AREA_MAX_TAPS][TBX_AREA_MAX_DST];      /* [tap][dx]: conflict-free for lanes over dx */
  float yalpha[TBX_AREA_MAX_TAPS][TBX_AREA_MAX_DST];
  uint8_t xdlo[TBX_AREA_MAX_SRC], xdhi[TBX_AREA_MAX_SRC]; /* destination columns fed by source column s: [xdlo[s], xdhi[s]] */
  uint8_t ydlo[TBX_AREA_MAX_SRC], ydhi[TBX_AREA_MAX_SRC];
} TbxAreaPlan;

/* Pre-resolved output patch of ONE HUD digit on the base frame (tbx_render_area.cuh): the bytes of output rectangle
 * [x0, x0+w) x [y0, y0+h) when the digit sprite is the only thing that differs from the base there.  w == 0: none. */
#define TBX_DP_MAX 128
#define TBX_DP_SLOTS 32 /* digit slots per game (score / lives / jumps fields, 10 slots each) */
typedef struct { uint8_t x0, y0, w, h; uint8_t px[TBX_DP_MAX]; } TbxDigitPatch;

#define TBX_WORDS(T) ((int)(sizeof(T) / 4))
#define TBX_W(T, field) ((int)(offsetof(T, field) / 4))

#endif
