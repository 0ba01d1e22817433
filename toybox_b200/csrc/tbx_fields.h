/* tbx_fields.h -- property paths of the JSON state schema compiled to SoA word offsets.
 *
 * The reference edits one env at a time through a JSON round trip (toybox/interventions/base.py:387-408) and reads
 * single values with get_property('a.b[3].c') (toybox/interventions/core.py:271-304).  Every scalar of the schema is
 * one (or two) 32-bit words of the env's record, so at batch size N a property is a strided plane:
 * tbx_field_lookup() turns a path into (word, kind, bit) and the get/set kernels touch that plane for all envs.
 * Paths follow the JSON the library exports (tbx_state_to_json).  Host code only.
 */
#ifndef TBX_FIELDS_H
#define TBX_FIELDS_H
#include <stdio.h>
#include <string.h>
#include "tbx_records.h"

enum { TBX_F_I32 = 0, TBX_F_F64 = 1, TBX_F_BOOL = 2, TBX_F_BIT = 3, TBX_F_OPT_I32 = 4 };

namespace tbxfields {

struct Field { int word, kind, bit; };

/* "name[i].rest": matches prefix `name[`, parses i, expects "]." + rest or "]" (rest empty) */
static inline bool indexed(const char *path, const char *name, int *i, const char **rest) {
  size_t n = strlen(name);
  if (strncmp(path, name, n) != 0 || path[n] != '[') return false;
  int k = 0, used = 0;
  if (sscanf(path + n + 1, "%d%n", &k, &used) != 1 || used <= 0 || path[n + 1 + used] != ']') return false;
  const char *r = path + n + 2 + used;
  if (*r == '.') r++;
  *i = k;
  *rest = r;
  return true;
}
static inline bool xy(const char *rest, const char *prefix, int *which) { /* "<prefix>.x" / "<prefix>.y" */
  size_t n = strlen(prefix);
  if (strncmp(rest, prefix, n) != 0 || rest[n] != '.') return false;
  if (!strcmp(rest + n + 1, "x")) { *which = 0; return true; }
  if (!strcmp(rest + n + 1, "y")) { *which = 1; return true; }
  return false;
}
#define TBX_FW(T, f) ((int)(offsetof(T, f) / 4))

static inline bool lookup(int game, const char *p, Field *f) {
  f->bit = -1;
  struct S { const char *name; int word, kind; };
  static const S common[] = {{"lives", TBX_FW(TbxHdr, lives), TBX_F_I32}, {"score", TBX_FW(TbxHdr, score), TBX_F_I32}, {"level", TBX_FW(TbxHdr, level), TBX_F_I32}};
  for (const S &s : common) if (!strcmp(p, s.name)) { f->word = s.word; f->kind = s.kind; return true; }
  int i = 0, w = 0;
  const char *rest = 0;
  if (game == TBX_BREAKOUT) {
    static const S flat[] = {
        {"paddle.position.x", TBX_FW(BrkRec, paddle_px), TBX_F_F64}, {"paddle.position.y", TBX_FW(BrkRec, paddle_py), TBX_F_F64},
        {"paddle.velocity.x", TBX_FW(BrkRec, paddle_vx), TBX_F_F64}, {"paddle.velocity.y", TBX_FW(BrkRec, paddle_vy), TBX_F_F64},
        {"paddle_width", TBX_FW(BrkRec, paddle_width), TBX_F_F64}, {"paddle_speed", TBX_FW(BrkRec, paddle_speed), TBX_F_F64},
        {"ball_radius", TBX_FW(BrkRec, ball_radius), TBX_F_F64}, {"is_dead", TBX_FW(BrkRec, is_dead), TBX_F_BOOL}, {"reset", TBX_FW(BrkRec, reset), TBX_F_BOOL}};
    for (const S &s : flat) if (!strcmp(p, s.name)) { f->word = s.word; f->kind = s.kind; return true; }
    if (indexed(p, "balls", &i, &rest) && i >= 0 && i < TBX_BRK_MAX_BALLS) {
      if (xy(rest, "position", &w)) { f->word = TBX_FW(BrkRec, ball) + 8 * i + 2 * w; f->kind = TBX_F_F64; return true; }
      if (xy(rest, "velocity", &w)) { f->word = TBX_FW(BrkRec, ball) + 8 * i + 4 + 2 * w; f->kind = TBX_F_F64; return true; }
    }
    if (indexed(p, "bricks", &i, &rest) && i >= 0 && i < TBX_BRK_MAX_BRICKS && !strcmp(rest, "alive")) {
      f->word = TBX_FW(BrkRec, alive) + (i >> 5); f->kind = TBX_F_BIT; f->bit = i & 31; return true;
    }
    return false;
  }
  if (game == TBX_SPACE_INVADERS) {
    static const S flat[] = {
        {"ship.x", TBX_FW(SiRec, ship_x), TBX_F_I32}, {"ship.y", TBX_FW(SiRec, ship_y), TBX_F_I32}, {"ship.w", TBX_FW(SiRec, ship_w), TBX_F_I32},
        {"ship.h", TBX_FW(SiRec, ship_h), TBX_F_I32}, {"ship.speed", TBX_FW(SiRec, ship_speed), TBX_F_I32},
        {"ship.death_counter", TBX_FW(SiRec, ship_death_counter), TBX_F_OPT_I32}, {"ship.alive", TBX_FW(SiRec, ship_alive), TBX_F_BOOL},
        {"ship.death_hit_1", TBX_FW(SiRec, ship_death_hit_1), TBX_F_BOOL},
        {"enemies_movement.move_counter", TBX_FW(SiRec, move_counter), TBX_F_I32},
        {"enemies_movement.visual_orientation", TBX_FW(SiRec, visual_orientation), TBX_F_BOOL},
        {"ufo.x", TBX_FW(SiRec, ufo_x), TBX_F_I32}, {"ufo.y", TBX_FW(SiRec, ufo_y), TBX_F_I32},
        {"ufo.appearance_counter", TBX_FW(SiRec, ufo_appearance_counter), TBX_F_OPT_I32}, {"ufo.death_counter", TBX_FW(SiRec, ufo_death_counter), TBX_F_OPT_I32},
        {"life_display_timer", TBX_FW(SiRec, life_display_timer), TBX_F_I32}, {"enemy_shot_delay", TBX_FW(SiRec, enemy_shot_delay), TBX_F_I32}};
    for (const S &s : flat) if (!strcmp(p, s.name)) { f->word = s.word; f->kind = s.kind; return true; }
    if (indexed(p, "enemies", &i, &rest) && i >= 0 && i < TBX_SI_N_ENEMIES) {
      static const S per[] = {{"x", TBX_FW(SiRec, en_x), TBX_F_I32}, {"y", TBX_FW(SiRec, en_y), TBX_F_I32}, {"row", TBX_FW(SiRec, en_row), TBX_F_I32},
                              {"col", TBX_FW(SiRec, en_col), TBX_F_I32}, {"id", TBX_FW(SiRec, en_id), TBX_F_I32}, {"points", TBX_FW(SiRec, en_points), TBX_F_I32},
                              {"death_counter", TBX_FW(SiRec, en_death), TBX_F_OPT_I32}};
      for (const S &s : per) if (!strcmp(rest, s.name)) { f->word = s.word + i; f->kind = s.kind; return true; }
      if (!strcmp(rest, "alive")) { f->word = TBX_FW(SiRec, en_alive) + (i >> 5); f->kind = TBX_F_BIT; f->bit = i & 31; return true; }
    }
    if (indexed(p, "shields", &i, &rest) && i >= 0 && i < TBX_SI_N_SHIELDS) {
      if (!strcmp(rest, "x")) { f->word = TBX_FW(SiRec, shield_x) + i; f->kind = TBX_F_I32; return true; }
      if (!strcmp(rest, "y")) { f->word = TBX_FW(SiRec, shield_y) + i; f->kind = TBX_F_I32; return true; }
    }
    return false;
  }
  /* Amidar */
  static const S flat[] = {{"jumps", TBX_FW(AmiRec, jumps), TBX_F_I32}, {"jump_timer", TBX_FW(AmiRec, jump_timer), TBX_F_I32},
                           {"chase_timer", TBX_FW(AmiRec, chase_timer), TBX_F_I32}};
  for (const S &s : flat) if (!strcmp(p, s.name)) { f->word = s.word; f->kind = s.kind; return true; }
  int mob = -1;
  if (!strncmp(p, "player.", 7)) { mob = TBX_FW(AmiRec, player); rest = p + 7; }
  else if (indexed(p, "enemies", &i, &rest) && i >= 0 && i < TBX_AMI_MAX_ENEMIES) mob = TBX_FW(AmiRec, enemies) + i * (int)(sizeof(AmiMob) / 4);
  if (mob >= 0) {
    if (xy(rest, "position", &w)) { f->word = mob + (w ? TBX_FW(AmiMob, y) : TBX_FW(AmiMob, x)); f->kind = TBX_F_I32; return true; }
    if (!strcmp(rest, "caught")) { f->word = mob + TBX_FW(AmiMob, caught); f->kind = TBX_F_BOOL; return true; }
    if (!strcmp(rest, "speed")) { f->word = mob + TBX_FW(AmiMob, speed); f->kind = TBX_F_I32; return true; }
  }
  if (indexed(p, "board.boxes", &i, &rest) && i >= 0 && i < TBX_AMI_MAX_BOXES && !strcmp(rest, "painted")) {
    f->word = TBX_FW(AmiRec, box_painted); f->kind = TBX_F_BIT; f->bit = i; return true;
  }
  return false;
}

} /* namespace tbxfields */
#endif
