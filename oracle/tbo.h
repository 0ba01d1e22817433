/* tbo.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar, one-env-at-a-time, array-of-structs restatement of the hot path that
 * the reference drives through the third-party `ctoybox==0.5.0` Rust library
 * (reference call sites: toybox/envs/atari/base.py:109,126,136,145,153 and
 * toybox/interventions/base.py:390-391,402-406).  The Rust source is NOT under
 * /root/reference and cannot be built here (no rustc, no network), so:
 *
 *   PARITY STATUS: everything the reference's fixtures/tests pin is reproduced
 *   and checked in tests/ (RNG algorithm + seeding, new_game child-RNG rule,
 *   ball-start draw, initial states of all three games, post-FIRE known
 *   answers).  The per-frame transition and raster rules beyond that are a
 *   *restatement from the published behaviour* => "parity unpinned" for them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (toybox_b200/)
 * never links or imports it.
 *
 * Compile with:  gcc -O2 -std=c99 -ffp-contract=off -fno-fast-math
 */
#ifndef TBO_H
#define TBO_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- input: the six booleans of ctoybox.Input (toybox/envs/atari/constants.py:3-13) */
enum {
  TBO_IN_LEFT = 1, TBO_IN_RIGHT = 2, TBO_IN_UP = 4, TBO_IN_DOWN = 8,
  TBO_IN_BUTTON1 = 16, TBO_IN_BUTTON2 = 32
};
/* ALE action id (constants.py:16-35) -> input bitmask; returns -1 for an invalid id */
int tbo_ale_action_to_input(int ale_action);

/* ---- RNG: xoroshiro128+ (55,14,36), SURVEY App. A.1 */
typedef struct { uint64_t s[2]; } tbo_rng;
void     tbo_rng_seed(tbo_rng *g, uint32_t seed);    /* [FIX] 0x193a6754a8a7d469^seed, 0x97830e05113ba7bb */
uint64_t tbo_rng_next_u64(tbo_rng *g);
uint32_t tbo_rng_next_u32(tbo_rng *g);               /* [FIX] high 32 bits of next_u64 */
tbo_rng  tbo_rng_child(tbo_rng *parent);             /* [FIX] state = two draws of the parent */
uint32_t tbo_rng_index(tbo_rng *g, uint32_t n);      /* [FIX] widening-multiply + rejection zone */
double   tbo_rng_f64(tbo_rng *g);                    /* (next_u64 >> 11) * 2^-53 */

/* ---- graphics */
typedef struct { uint8_t r, g, b, a; } tbo_color;
typedef struct { int w, h; uint8_t *rgba; } tbo_canvas;    /* row-major, 4 bytes per pixel */
uint8_t tbo_luma(tbo_color c);                              /* (0.299 r + 0.587 g + 0.114 b) as u8, f64 */
void tbo_clear(tbo_canvas *cv, tbo_color c);
void tbo_rect(tbo_canvas *cv, tbo_color c, int x, int y, int w, int h);          /* clipped */
/* 1-bit sprite: rows[] hold `w` bits each, MSB-first within the low `w` bits; scale = integer zoom */
void tbo_sprite1(tbo_canvas *cv, tbo_color c, int x, int y, int w, int h, const uint32_t *rows, int sx, int sy);
void tbo_digits(tbo_canvas *cv, tbo_color c, int x_right, int y, int value, int sx, int sy); /* right-aligned 3x5 font */
void tbo_rgba_to_gray(const uint8_t *rgba, int npix, uint8_t *gray);
void tbo_rgba_to_rgb(const uint8_t *rgba, int npix, uint8_t *rgb);
/* cv2.resize(..., interpolation=INTER_AREA) for uint8, `cn` interleaved channels, non-integer scale path
 * (baselines/baselines/common/atari_wrappers.py:243).  Pinned against cv2 in tests. */
void tbo_resize_area_u8(const uint8_t *src, int sw, int sh, int cn, uint8_t *dst, int dw, int dh);
extern const uint32_t TBO_FONT3X5[10][5];

/* =========================================================================================
 * Breakout (f64)   schema: toybox/interventions/breakout.py:49-68,132,198,276
 * ========================================================================================= */
#define TBO_BRK_W 240
#define TBO_BRK_H 160
#define TBO_BRK_MAX_BRICKS 144
#define TBO_BRK_MAX_BALLS 4
#define TBO_BRK_MAX_ROWS 8
#define TBO_BRK_MAX_STARTS 8
#define TBO_BRK_MAX_SEGS 16

typedef struct { double x, y; } tbo_vec2;
typedef struct { double x, y, angle_degrees; } tbo_ball_start;
typedef struct {
  tbo_color bg_color, frame_color, paddle_color, ball_color;
  int32_t n_rows; tbo_color row_colors[TBO_BRK_MAX_ROWS]; int32_t row_scores[TBO_BRK_MAX_ROWS];
  int32_t start_lives, paddle_discrete_segments, ball_speed_row_depth;
  double ball_speed_slow, ball_speed_fast;
  int32_t n_starts; tbo_ball_start ball_start_positions[TBO_BRK_MAX_STARTS];
  tbo_rng rand;
} tbo_brk_cfg;

typedef struct {
  tbo_vec2 position, size; tbo_color color;
  int32_t points, depth, row, col; uint8_t alive, destructible;
} tbo_brk_brick;
typedef struct { tbo_vec2 position, velocity; } tbo_brk_body;
typedef struct {
  tbo_rng rand;
  tbo_brk_body paddle;
  int32_t n_balls; tbo_brk_body balls[TBO_BRK_MAX_BALLS];
  int32_t n_bricks; tbo_brk_brick bricks[TBO_BRK_MAX_BRICKS];
  double paddle_width, paddle_speed, ball_radius;
  int32_t lives, score, level; uint8_t is_dead, reset;
} tbo_brk_state;

void tbo_brk_default_cfg(tbo_brk_cfg *c);
void tbo_brk_new_game(tbo_brk_cfg *c, tbo_brk_state *s);          /* advances c->rand by 2 draws */
void tbo_brk_step(const tbo_brk_cfg *c, tbo_brk_state *s, int input);
void tbo_brk_render(const tbo_brk_cfg *c, const tbo_brk_state *s, uint8_t *rgba);

/* =========================================================================================
 * Space Invaders (i32)   schema: toybox/interventions/space_invaders.py:16-32,38,60,101,116,146
 * ========================================================================================= */
#define TBO_SI_W 320
#define TBO_SI_H 210
#define TBO_SI_N_ENEMIES 36
#define TBO_SI_MAX_LASERS 4
#define TBO_SI_N_SHIELDS 3
#define TBO_SI_SHIELD_W 16
#define TBO_SI_SHIELD_H 18
#define TBO_NONE (-2147483647 - 1)     /* Option<i32>::None */
enum { TBO_DIR_UP = 0, TBO_DIR_DOWN = 1, TBO_DIR_LEFT = 2, TBO_DIR_RIGHT = 3 };
enum { TBO_SI_PROTO_TARGET_PLAYER = 0, TBO_SI_PROTO_RANDOM = 1 };

typedef struct {
  double jitter; int32_t enemy_protocol; int32_t start_lives;
  int32_t shields[TBO_SI_N_SHIELDS][2]; int32_t row_scores[6];
  tbo_rng rand;
} tbo_si_cfg;
typedef struct { int32_t x, y, w, h, t, movement, speed; tbo_color color; } tbo_si_laser;
typedef struct { int32_t x, y, row, col, id, points, death_counter; uint8_t alive; } tbo_si_enemy;
typedef struct {
  tbo_rng rand;
  struct { int32_t x, y, w, h, speed, death_counter; uint8_t alive, death_hit_1; tbo_color color; } ship;
  uint8_t has_ship_laser; tbo_si_laser ship_laser;
  tbo_si_enemy enemies[TBO_SI_N_ENEMIES];
  struct { int32_t move_counter, move_dir; uint8_t visual_orientation; } enemies_movement;
  int32_t n_enemy_lasers; tbo_si_laser enemy_lasers[TBO_SI_MAX_LASERS];
  struct { int32_t x, y; uint16_t rows[TBO_SI_SHIELD_H]; } shields[TBO_SI_N_SHIELDS];  /* bit (15-c) of rows[r] = opaque */
  struct { int32_t x, y, appearance_counter, death_counter; } ufo;
  int32_t life_display_timer, enemy_shot_delay, score, lives, level;
} tbo_si_state;

void tbo_si_default_cfg(tbo_si_cfg *c);
void tbo_si_new_game(tbo_si_cfg *c, tbo_si_state *s);
void tbo_si_step(const tbo_si_cfg *c, tbo_si_state *s, int input);
void tbo_si_render(const tbo_si_cfg *c, const tbo_si_state *s, uint8_t *rgba);
extern const uint16_t TBO_SI_SHIELD_ROWS[TBO_SI_SHIELD_H];
extern const tbo_color TBO_SI_SHIELD_COLOR;

/* =========================================================================================
 * Amidar (i32)   schema: toybox/interventions/amidar.py:22-34,171,195,216,300,316
 * ========================================================================================= */
#define TBO_AMI_W 160
#define TBO_AMI_H 250
#define TBO_AMI_BW 32
#define TBO_AMI_BH 31
#define TBO_AMI_MAX_ENEMIES 8
#define TBO_AMI_MAX_BOXES 32
#define TBO_AMI_MAX_JUNCTIONS 64
#define TBO_AMI_HIST 8
#define TBO_AMI_MAX_ROUTES 16
#define TBO_AMI_MAX_ROUTE_LEN 64
enum { TBO_TILE_EMPTY = 0, TBO_TILE_UNPAINTED = 1, TBO_TILE_CHASE = 2, TBO_TILE_PAINTED = 3 };
enum { TBO_AI_PLAYER = 0, TBO_AI_LOOKUP = 1, TBO_AI_PERIMETER = 2, TBO_AI_AMIDAR = 3, TBO_AI_TARGET = 4, TBO_AI_RANDOM = 5 };

typedef struct {
  int32_t kind;
  int32_t next, default_route_index;             /* EnemyLookupAI */
  int32_t start_tx, start_ty;                    /* Perimeter / Amidar / Target / Random */
  int32_t vert, horiz, start_vert, start_horiz;  /* EnemyAmidarMvmt */
  int32_t start_dir, dir;                        /* Target / Random */
  int32_t vision_distance;                       /* Target */
  int32_t seen_tx, seen_ty; uint8_t has_seen;    /* Target: player_seen Option<TilePoint> */
} tbo_ami_ai;
typedef struct {
  int32_t x, y;                                  /* WorldPoint */
  uint8_t has_step; int32_t step_tx, step_ty;    /* Option<TilePoint> */
  int32_t n_history; int32_t history[TBO_AMI_HIST];   /* most recent first */
  uint8_t caught; int32_t speed; tbo_ami_ai ai;
} tbo_ami_mob;
typedef struct { int32_t tl_tx, tl_ty, br_tx, br_ty; uint8_t painted, triggers_chase; } tbo_ami_box;
typedef struct {
  tbo_color bg_color, player_color, unpainted_color, painted_color, enemy_color, inner_painted_color;
  int32_t start_lives, start_jumps, chase_time, chase_score_bonus, jump_time, box_bonus;
  uint8_t render_images, default_board_bugs;
  int32_t player_start_tx, player_start_ty;
  uint8_t board[TBO_AMI_BH][TBO_AMI_BW];         /* tile tags as in the config ASCII ('p' => PAINTED) */
  int32_t n_enemies; tbo_ami_ai enemies[TBO_AMI_MAX_ENEMIES];
  int32_t n_routes; int32_t route_len[TBO_AMI_MAX_ROUTES]; int32_t routes[TBO_AMI_MAX_ROUTES][TBO_AMI_MAX_ROUTE_LEN];
  tbo_rng rand;
} tbo_ami_cfg;
typedef struct {
  tbo_rng rand;
  int32_t score, lives, level, jumps, jump_timer, chase_timer;
  tbo_ami_mob player;
  int32_t n_enemies; tbo_ami_mob enemies[TBO_AMI_MAX_ENEMIES];
  uint8_t tiles[TBO_AMI_BH][TBO_AMI_BW];
  int32_t n_boxes; tbo_ami_box boxes[TBO_AMI_MAX_BOXES];
  int32_t n_junctions; int32_t junctions[TBO_AMI_MAX_JUNCTIONS];
  int32_t chase_junctions[4]; int32_t n_chase_junctions;
} tbo_ami_state;

void tbo_ami_default_cfg(tbo_ami_cfg *c);
void tbo_ami_new_game(tbo_ami_cfg *c, tbo_ami_state *s);      /* [FIX] does NOT advance c->rand */
void tbo_ami_step(const tbo_ami_cfg *c, tbo_ami_state *s, int input);
void tbo_ami_render(const tbo_ami_cfg *c, const tbo_ami_state *s, uint8_t *rgba);
void tbo_ami_tile_to_world(int tx, int ty, int *wx, int *wy);
void tbo_ami_world_to_tile(int wx, int wy, int *tx, int *ty);

/* sizes, so the Python ctypes mirrors can be checked */
size_t tbo_sizeof(int which);   /* 0 brk_cfg 1 brk_state 2 si_cfg 3 si_state 4 ami_cfg 5 ami_state */

/* ---- synthetic action stream shared with the product's bench/tests:
 * action index for (seed, env, t) uniform over n_legal, counter based */
uint32_t tbo_action_index(uint64_t seed, uint64_t env, uint64_t t, uint32_t n_legal);

#ifdef __cplusplus
}
#endif
#endif
