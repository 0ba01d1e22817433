"""CPU ORACLE python binding (test infrastructure, NOT product code).

ctypes mirrors of oracle/tbo.h plus the JSON state/config codec in the reference's schema
(toybox/interventions/{breakout,amidar,space_invaders}.py `expected_keys`).  `OracleToybox` has the
method surface of ctoybox.Toybox that the reference consumes so the reference's own
python (toybox.interventions, test/interventions/*) can be run against the oracle in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Parity status: see oracle/tbo.h ("parity unpinned" beyond the fixtures' known answers).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GAMES = {"breakout": 0, "amidar": 1, "space_invaders": 2}
DIMS = {"breakout": (240, 160), "amidar": (160, 250), "space_invaders": (320, 210)}   # (W, H)
LEGAL = {"breakout": [0, 1, 3, 4], "amidar": [0, 1, 2, 3, 4, 5, 10, 11, 12, 13], "space_invaders": [0, 1, 3, 4, 11, 12]}
MODES = {"rgba": 0, "rgb": 1, "gray": 2, "gray84": 3}
NONE = -2147483648
DIRS = ["Up", "Down", "Left", "Right"]
TILES = ["Empty", "Unpainted", "ChaseMarker", "Painted"]


def build(force=False):
    so = os.path.join(_HERE, "libtbo.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.tbo_sizeof.restype = C.c_size_t
        _LIB.tbo_rng_next_u64.restype = C.c_uint64
        _LIB.tbo_rng_next_u32.restype = C.c_uint32
        _LIB.tbo_rng_index.restype = C.c_uint32
        _LIB.tbo_rng_f64.restype = C.c_double
        _LIB.tbo_action_index.restype = C.c_uint32
        _LIB.tbo_action_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        _LIB.tbo_frame_bytes.restype = C.c_size_t
        _LIB.tbo_batch_rollout.restype = C.c_double
        _LIB.tbo_batch_rollout.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                           C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p]
        for i, st in enumerate([BrkCfg, BrkState, SiCfg, SiState, AmiCfg, AmiState]):
            assert _LIB.tbo_sizeof(i) == C.sizeof(st), (i, _LIB.tbo_sizeof(i), C.sizeof(st))
    return _LIB


# ------------------------------------------------------------------------------------------------
# struct mirrors
class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 2)]


class Color(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


class Vec2(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double)]


class BallStart(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("angle_degrees", C.c_double)]


class BrkCfg(C.Structure):
    _fields_ = [("bg_color", Color), ("frame_color", Color), ("paddle_color", Color), ("ball_color", Color),
                ("n_rows", C.c_int32), ("row_colors", Color * 8), ("row_scores", C.c_int32 * 8),
                ("start_lives", C.c_int32), ("paddle_discrete_segments", C.c_int32), ("ball_speed_row_depth", C.c_int32),
                ("ball_speed_slow", C.c_double), ("ball_speed_fast", C.c_double),
                ("n_starts", C.c_int32), ("ball_start_positions", BallStart * 8), ("rand", Rng)]


class BrkBrick(C.Structure):
    _fields_ = [("position", Vec2), ("size", Vec2), ("color", Color), ("points", C.c_int32), ("depth", C.c_int32),
                ("row", C.c_int32), ("col", C.c_int32), ("alive", C.c_uint8), ("destructible", C.c_uint8)]


class BrkBody(C.Structure):
    _fields_ = [("position", Vec2), ("velocity", Vec2)]


class BrkState(C.Structure):
    _fields_ = [("rand", Rng), ("paddle", BrkBody), ("n_balls", C.c_int32), ("balls", BrkBody * 4),
                ("n_bricks", C.c_int32), ("bricks", BrkBrick * 144),
                ("paddle_width", C.c_double), ("paddle_speed", C.c_double), ("ball_radius", C.c_double),
                ("lives", C.c_int32), ("score", C.c_int32), ("level", C.c_int32), ("is_dead", C.c_uint8), ("reset", C.c_uint8)]


class SiCfg(C.Structure):
    _fields_ = [("jitter", C.c_double), ("enemy_protocol", C.c_int32), ("start_lives", C.c_int32),
                ("shields", (C.c_int32 * 2) * 3), ("row_scores", C.c_int32 * 6), ("rand", Rng)]


class SiLaser(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("t", C.c_int32),
                ("movement", C.c_int32), ("speed", C.c_int32), ("color", Color)]


class SiEnemy(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("row", C.c_int32), ("col", C.c_int32), ("id", C.c_int32),
                ("points", C.c_int32), ("death_counter", C.c_int32), ("alive", C.c_uint8)]


class SiShip(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("speed", C.c_int32),
                ("death_counter", C.c_int32), ("alive", C.c_uint8), ("death_hit_1", C.c_uint8), ("color", Color)]


class SiMove(C.Structure):
    _fields_ = [("move_counter", C.c_int32), ("move_dir", C.c_int32), ("visual_orientation", C.c_uint8)]


class SiShield(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("rows", C.c_uint16 * 18)]


class SiUfo(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("appearance_counter", C.c_int32), ("death_counter", C.c_int32)]


class SiState(C.Structure):
    _fields_ = [("rand", Rng), ("ship", SiShip), ("has_ship_laser", C.c_uint8), ("ship_laser", SiLaser),
                ("enemies", SiEnemy * 36), ("enemies_movement", SiMove),
                ("n_enemy_lasers", C.c_int32), ("enemy_lasers", SiLaser * 4), ("shields", SiShield * 3), ("ufo", SiUfo),
                ("life_display_timer", C.c_int32), ("enemy_shot_delay", C.c_int32), ("score", C.c_int32),
                ("lives", C.c_int32), ("level", C.c_int32)]


class AmiAi(C.Structure):
    _fields_ = [("kind", C.c_int32), ("next", C.c_int32), ("default_route_index", C.c_int32),
                ("start_tx", C.c_int32), ("start_ty", C.c_int32),
                ("vert", C.c_int32), ("horiz", C.c_int32), ("start_vert", C.c_int32), ("start_horiz", C.c_int32),
                ("start_dir", C.c_int32), ("dir", C.c_int32), ("vision_distance", C.c_int32),
                ("seen_tx", C.c_int32), ("seen_ty", C.c_int32), ("has_seen", C.c_uint8)]


class AmiMob(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("has_step", C.c_uint8), ("step_tx", C.c_int32), ("step_ty", C.c_int32),
                ("n_history", C.c_int32), ("history", C.c_int32 * 8), ("caught", C.c_uint8), ("speed", C.c_int32), ("ai", AmiAi)]


class AmiBox(C.Structure):
    _fields_ = [("tl_tx", C.c_int32), ("tl_ty", C.c_int32), ("br_tx", C.c_int32), ("br_ty", C.c_int32),
                ("painted", C.c_uint8), ("triggers_chase", C.c_uint8)]


class AmiCfg(C.Structure):
    _fields_ = [("bg_color", Color), ("player_color", Color), ("unpainted_color", Color), ("painted_color", Color),
                ("enemy_color", Color), ("inner_painted_color", Color),
                ("start_lives", C.c_int32), ("start_jumps", C.c_int32), ("chase_time", C.c_int32),
                ("chase_score_bonus", C.c_int32), ("jump_time", C.c_int32), ("box_bonus", C.c_int32),
                ("render_images", C.c_uint8), ("default_board_bugs", C.c_uint8),
                ("player_start_tx", C.c_int32), ("player_start_ty", C.c_int32),
                ("board", (C.c_uint8 * 32) * 31),
                ("n_enemies", C.c_int32), ("enemies", AmiAi * 8),
                ("n_routes", C.c_int32), ("route_len", C.c_int32 * 16), ("routes", (C.c_int32 * 64) * 16), ("rand", Rng)]


class AmiState(C.Structure):
    _fields_ = [("rand", Rng), ("score", C.c_int32), ("lives", C.c_int32), ("level", C.c_int32), ("jumps", C.c_int32),
                ("jump_timer", C.c_int32), ("chase_timer", C.c_int32), ("player", AmiMob),
                ("n_enemies", C.c_int32), ("enemies", AmiMob * 8), ("tiles", (C.c_uint8 * 32) * 31),
                ("n_boxes", C.c_int32), ("boxes", AmiBox * 32), ("n_junctions", C.c_int32), ("junctions", C.c_int32 * 64),
                ("chase_junctions", C.c_int32 * 4), ("n_chase_junctions", C.c_int32)]


CFG = {"breakout": BrkCfg, "amidar": AmiCfg, "space_invaders": SiCfg}
STATE = {"breakout": BrkState, "amidar": AmiState, "space_invaders": SiState}
PREFIX = {"breakout": "brk", "amidar": "ami", "space_invaders": "si"}


# ------------------------------------------------------------------------------------------------
# JSON codec (reference schema)
def _col(c):
    return {"r": c.r, "g": c.g, "b": c.b, "a": c.a}


def _setcol(c, d):
    c.r, c.g, c.b, c.a = (max(0, min(255, int(d[k]))) for k in "rgba")


def _vec(v):
    return {"x": v.x, "y": v.y}


def _opt(v):
    return None if v == NONE else v


def _unopt(v):
    return NONE if v is None else int(v)


def _rand(r):
    return {"state": [int(r.s[0]), int(r.s[1])]}


def _setrand(r, d):
    r.s[0], r.s[1] = int(d["state"][0]), int(d["state"][1])


def brk_state_to_json(s):
    return {
        "score": s.score, "lives": s.lives, "rand": _rand(s.rand), "level": s.level,
        "paddle": {"velocity": _vec(s.paddle.velocity), "position": _vec(s.paddle.position)},
        "paddle_width": s.paddle_width, "paddle_speed": s.paddle_speed, "ball_radius": s.ball_radius,
        "balls": [{"position": _vec(s.balls[i].position), "velocity": _vec(s.balls[i].velocity)} for i in range(s.n_balls)],
        "bricks": [{"destructible": bool(b.destructible), "depth": b.depth, "color": _col(b.color), "alive": bool(b.alive),
                    "points": b.points, "size": _vec(b.size), "position": _vec(b.position), "row": b.row, "col": b.col}
                   for b in (s.bricks[i] for i in range(s.n_bricks))],
        "reset": bool(s.reset), "is_dead": bool(s.is_dead)}


def brk_state_from_json(s, d):
    s.score = int(d["score"] if "score" in d else d["points"])     # fixture era: `points`
    s.lives = int(d["lives"]); s.level = int(d.get("level", 1)); _setrand(s.rand, d["rand"])
    for name in ("velocity", "position"):
        getattr(s.paddle, name).x = float(d["paddle"][name]["x"]); getattr(s.paddle, name).y = float(d["paddle"][name]["y"])
    s.paddle_width = float(d["paddle_width"]); s.paddle_speed = float(d["paddle_speed"]); s.ball_radius = float(d["ball_radius"])
    if len(d["balls"]) > 4 or len(d["bricks"]) > 144:
        raise ValueError("too many balls/bricks")
    s.n_balls = len(d["balls"])
    for i, b in enumerate(d["balls"]):
        s.balls[i].position.x, s.balls[i].position.y = float(b["position"]["x"]), float(b["position"]["y"])
        s.balls[i].velocity.x, s.balls[i].velocity.y = float(b["velocity"]["x"]), float(b["velocity"]["y"])
    s.n_bricks = len(d["bricks"])
    for i, b in enumerate(d["bricks"]):
        k = s.bricks[i]
        k.destructible = bool(b["destructible"]); k.depth = int(b["depth"]); _setcol(k.color, b["color"]); k.alive = bool(b["alive"])
        k.points = int(b["points"]); k.size.x, k.size.y = float(b["size"]["x"]), float(b["size"]["y"])
        k.position.x, k.position.y = float(b["position"]["x"]), float(b["position"]["y"]); k.row = int(b["row"]); k.col = int(b["col"])
    s.reset = bool(d.get("reset") or False); s.is_dead = bool(d["is_dead"])


def brk_cfg_to_json(c):
    return {"paddle_discrete_segments": c.paddle_discrete_segments,
            "ball_start_positions": [{"angle_degrees": p.angle_degrees, "y": p.y, "x": p.x}
                                     for p in (c.ball_start_positions[i] for i in range(c.n_starts))],
            "start_lives": c.start_lives, "row_scores": [c.row_scores[i] for i in range(c.n_rows)],
            "ball_speed_row_depth": c.ball_speed_row_depth, "bg_color": _col(c.bg_color), "rand": _rand(c.rand),
            "row_colors": [_col(c.row_colors[i]) for i in range(c.n_rows)], "frame_color": _col(c.frame_color),
            "paddle_color": _col(c.paddle_color), "ball_color": _col(c.ball_color),
            "ball_speed_fast": c.ball_speed_fast, "ball_speed_slow": c.ball_speed_slow}


def brk_cfg_from_json(c, d):
    c.paddle_discrete_segments = int(d["paddle_discrete_segments"])
    c.n_starts = len(d["ball_start_positions"])
    for i, p in enumerate(d["ball_start_positions"]):
        q = c.ball_start_positions[i]; q.x, q.y, q.angle_degrees = float(p["x"]), float(p["y"]), float(p["angle_degrees"])
    c.start_lives = int(d["start_lives"]); c.n_rows = len(d["row_scores"])
    for i, v in enumerate(d["row_scores"]):
        c.row_scores[i] = int(v)
    for i, v in enumerate(d["row_colors"]):
        _setcol(c.row_colors[i], v)
    c.ball_speed_row_depth = int(d["ball_speed_row_depth"])
    for k in ("bg_color", "frame_color", "paddle_color", "ball_color"):
        _setcol(getattr(c, k), d[k])
    _setrand(c.rand, d["rand"]); c.ball_speed_fast = float(d["ball_speed_fast"]); c.ball_speed_slow = float(d["ball_speed_slow"])


def _laser(l):
    return {"x": l.x, "y": l.y, "w": l.w, "h": l.h, "t": l.t, "movement": DIRS[l.movement], "speed": l.speed, "color": _col(l.color)}


def _setlaser(l, d):
    l.x, l.y, l.w, l.h, l.t, l.speed = (int(d[k]) for k in ("x", "y", "w", "h", "t", "speed"))
    l.movement = DIRS.index(d["movement"]); _setcol(l.color, d["color"])


def si_state_to_json(s):
    sc = (172, 80, 48, 255)
    def shield(sh):
        return {"x": sh.x, "y": sh.y,
                "data": [[({"r": sc[0], "g": sc[1], "b": sc[2], "a": sc[3]} if (sh.rows[r] >> (15 - q)) & 1
                           else {"r": 0, "g": 0, "b": 0, "a": 0}) for q in range(16)] for r in range(18)]}
    return {
        "score": s.score, "lives": s.lives, "rand": _rand(s.rand), "level": s.level,
        "ship": {"x": s.ship.x, "y": s.ship.y, "w": s.ship.w, "h": s.ship.h, "speed": s.ship.speed, "color": _col(s.ship.color),
                 "alive": bool(s.ship.alive), "death_counter": _opt(s.ship.death_counter), "death_hit_1": bool(s.ship.death_hit_1)},
        "ship_laser": _laser(s.ship_laser) if s.has_ship_laser else None,
        "enemies": [{"x": e.x, "y": e.y, "row": e.row, "col": e.col, "id": e.id, "alive": bool(e.alive), "points": e.points,
                     "death_counter": _opt(e.death_counter)} for e in s.enemies],
        "enemies_movement": {"move_counter": s.enemies_movement.move_counter, "move_dir": DIRS[s.enemies_movement.move_dir],
                             "visual_orientation": bool(s.enemies_movement.visual_orientation)},
        "enemy_lasers": [_laser(s.enemy_lasers[i]) for i in range(s.n_enemy_lasers)],
        "shields": [shield(sh) for sh in s.shields],
        "ufo": {"x": s.ufo.x, "y": s.ufo.y, "appearance_counter": _opt(s.ufo.appearance_counter), "death_counter": _opt(s.ufo.death_counter)},
        "life_display_timer": s.life_display_timer, "enemy_shot_delay": s.enemy_shot_delay}


def si_state_from_json(s, d):
    s.score = int(d["score"]); s.lives = int(d["lives"]); _setrand(s.rand, d["rand"])
    s.level = int(d["level"]) if "level" in d else int(d.get("levels_completed", 0)) + 1
    sh = d["ship"]
    s.ship.x, s.ship.y, s.ship.w, s.ship.h, s.ship.speed = (int(sh[k]) for k in ("x", "y", "w", "h", "speed"))
    _setcol(s.ship.color, sh["color"]); s.ship.alive = bool(sh["alive"]); s.ship.death_counter = _unopt(sh["death_counter"])
    s.ship.death_hit_1 = bool(sh["death_hit_1"])
    s.has_ship_laser = d["ship_laser"] is not None
    if s.has_ship_laser:
        _setlaser(s.ship_laser, d["ship_laser"])
    if len(d["enemies"]) != 36 or len(d["enemy_lasers"]) > 4 or len(d["shields"]) != 3:
        raise ValueError("unsupported collection size")
    for i, e in enumerate(d["enemies"]):
        k = s.enemies[i]
        k.x, k.y, k.row, k.col, k.id, k.points = (int(e[f]) for f in ("x", "y", "row", "col", "id", "points"))
        k.alive = bool(e["alive"]); k.death_counter = _unopt(e["death_counter"])
    if "enemies_movement" in d:
        m = d["enemies_movement"]
        s.enemies_movement.move_counter = int(m["move_counter"]); s.enemies_movement.move_dir = DIRS.index(m["move_dir"])
        s.enemies_movement.visual_orientation = bool(m["visual_orientation"])
    else:                                                                  # fixture era: per-enemy fields
        e0 = d["enemies"][0]
        s.enemies_movement.move_counter = int(e0["move_counter"]); s.enemies_movement.move_dir = 3 if e0["move_right"] else 2
        s.enemies_movement.visual_orientation = bool(e0["orientation_init"])
    s.n_enemy_lasers = len(d["enemy_lasers"])
    for i, l in enumerate(d["enemy_lasers"]):
        _setlaser(s.enemy_lasers[i], l)
    for i, shd in enumerate(d["shields"]):
        s.shields[i].x = int(shd["x"]); s.shields[i].y = int(shd["y"])
        if len(shd["data"]) != 18 or any(len(r) != 16 for r in shd["data"]):
            raise ValueError("shield sprite must be 18x16")
        for r in range(18):
            bits = 0
            for q in range(16):
                if int(shd["data"][r][q]["a"]) != 0:
                    bits |= 1 << (15 - q)
            s.shields[i].rows[r] = bits
    u = d["ufo"]
    s.ufo.x = int(u["x"]); s.ufo.y = int(u["y"]); s.ufo.appearance_counter = _unopt(u["appearance_counter"])
    s.ufo.death_counter = _unopt(u["death_counter"])
    s.life_display_timer = int(d["life_display_timer"]); s.enemy_shot_delay = int(d["enemy_shot_delay"])


SI_PROTOCOLS = ["TargetPlayer", "Random"]


def si_cfg_to_json(c):
    return {"jitter": c.jitter, "shields": [[c.shields[i][0], c.shields[i][1]] for i in range(3)], "rand": _rand(c.rand),
            "row_scores": list(c.row_scores), "enemy_protocol": SI_PROTOCOLS[c.enemy_protocol], "start_lives": c.start_lives}


def si_cfg_from_json(c, d):
    c.jitter = float(d["jitter"]); c.enemy_protocol = SI_PROTOCOLS.index(d["enemy_protocol"]); c.start_lives = int(d["start_lives"])
    for i in range(3):
        c.shields[i][0], c.shields[i][1] = int(d["shields"][i][0]), int(d["shields"][i][1])
    for i in range(6):
        c.row_scores[i] = int(d["row_scores"][i])
    _setrand(c.rand, d["rand"])


AI_NAMES = {1: "EnemyLookupAI", 2: "EnemyPerimeterAI", 3: "EnemyAmidarMvmt", 4: "EnemyTargetPlayer", 5: "EnemyRandomMvmt"}


def _ai(a):
    st = {"tx": a.start_tx, "ty": a.start_ty}
    if a.kind == 0:
        return "Player"
    if a.kind == 1:
        return {"EnemyLookupAI": {"next": a.next, "default_route_index": a.default_route_index}}
    if a.kind == 2:
        return {"EnemyPerimeterAI": {"start": st}}
    if a.kind == 3:
        return {"EnemyAmidarMvmt": {"vert": DIRS[a.vert], "horiz": DIRS[a.horiz], "start_vert": DIRS[a.start_vert],
                                    "start_horiz": DIRS[a.start_horiz], "start": st}}
    if a.kind == 4:
        return {"EnemyTargetPlayer": {"start": st, "start_dir": DIRS[a.start_dir], "vision_distance": a.vision_distance,
                                      "dir": DIRS[a.dir], "player_seen": {"tx": a.seen_tx, "ty": a.seen_ty} if a.has_seen else None}}
    return {"EnemyRandomMvmt": {"start": st, "start_dir": DIRS[a.start_dir], "dir": DIRS[a.dir]}}


def _setai(a, d):
    C.memset(C.byref(a), 0, C.sizeof(a))
    if d == "Player":
        a.kind = 0
        return
    (name, kw), = d.items()
    a.kind = {v: k for k, v in AI_NAMES.items()}[name]
    if "start" in kw:
        a.start_tx, a.start_ty = int(kw["start"]["tx"]), int(kw["start"]["ty"])
    for k in ("next", "default_route_index", "vision_distance"):
        if k in kw:
            setattr(a, k, int(kw[k]))
    for k in ("vert", "horiz", "start_vert", "start_horiz", "start_dir", "dir"):
        if k in kw:
            setattr(a, k, DIRS.index(kw[k]))
    if kw.get("player_seen") is not None:
        a.has_seen = 1; a.seen_tx, a.seen_ty = int(kw["player_seen"]["tx"]), int(kw["player_seen"]["ty"])


def _mob(m):
    return {"history": [m.history[i] for i in range(m.n_history)],
            "step": {"tx": m.step_tx, "ty": m.step_ty} if m.has_step else None,
            "position": {"x": m.x, "y": m.y}, "caught": bool(m.caught), "speed": m.speed, "ai": _ai(m.ai)}


def _setmob(m, d):
    if len(d["history"]) > 8:
        raise ValueError("history longer than 8")
    m.n_history = len(d["history"])
    for i, h in enumerate(d["history"]):
        m.history[i] = int(h)
    m.has_step = d["step"] is not None
    m.step_tx, m.step_ty = (int(d["step"]["tx"]), int(d["step"]["ty"])) if m.has_step else (0, 0)
    m.x, m.y = int(d["position"]["x"]), int(d["position"]["y"]); m.caught = bool(d["caught"]); m.speed = int(d["speed"])
    _setai(m.ai, d["ai"])


def ami_state_to_json(s):
    return {
        "score": s.score, "lives": s.lives, "rand": _rand(s.rand), "level": s.level,
        "enemies": [_mob(s.enemies[i]) for i in range(s.n_enemies)], "player": _mob(s.player),
        "jumps": s.jumps, "jump_timer": s.jump_timer, "chase_timer": s.chase_timer,
        "board": {"boxes": [{"triggers_chase": bool(b.triggers_chase), "top_left": {"tx": b.tl_tx, "ty": b.tl_ty},
                             "bottom_right": {"tx": b.br_tx, "ty": b.br_ty}, "painted": bool(b.painted)}
                            for b in (s.boxes[i] for i in range(s.n_boxes))],
                  "tiles": [[TILES[s.tiles[y][x]] for x in range(32)] for y in range(31)],
                  "height": 31, "chase_junctions": [s.chase_junctions[i] for i in range(s.n_chase_junctions)], "width": 32,
                  "junctions": [s.junctions[i] for i in range(s.n_junctions)]}}


def ami_state_from_json(s, d):
    s.score = int(d["score"]); s.lives = int(d["lives"]); _setrand(s.rand, d["rand"]); s.level = int(d.get("level", 1))
    if len(d["enemies"]) > 8:
        raise ValueError("more than 8 enemies")
    s.n_enemies = len(d["enemies"])
    for i, e in enumerate(d["enemies"]):
        _setmob(s.enemies[i], e)
    _setmob(s.player, d["player"])
    s.jumps = int(d["jumps"]); s.jump_timer = int(d["jump_timer"]); s.chase_timer = int(d["chase_timer"])
    b = d["board"]
    if b["width"] != 32 or b["height"] != 31 or len(b["boxes"]) > 32 or len(b["junctions"]) > 64 or len(b["chase_junctions"]) > 4:
        raise ValueError("unsupported board")
    s.n_boxes = len(b["boxes"])
    for i, x in enumerate(b["boxes"]):
        k = s.boxes[i]
        k.tl_tx, k.tl_ty, k.br_tx, k.br_ty = x["top_left"]["tx"], x["top_left"]["ty"], x["bottom_right"]["tx"], x["bottom_right"]["ty"]
        k.painted = bool(x["painted"]); k.triggers_chase = bool(x["triggers_chase"])
    for y in range(31):
        for x in range(32):
            s.tiles[y][x] = TILES.index(b["tiles"][y][x])
    s.n_junctions = len(b["junctions"])
    for i, j in enumerate(b["junctions"]):
        s.junctions[i] = int(j)
    s.n_chase_junctions = len(b["chase_junctions"])
    for i, j in enumerate(b["chase_junctions"]):
        s.chase_junctions[i] = int(j)


_BOARD_CH = {0: " ", 1: "=", 2: "c", 3: "p"}


def ami_cfg_to_json(c):
    return {"box_bonus": c.box_bonus, "inner_painted_color": _col(c.inner_painted_color), "jump_time": c.jump_time,
            "render_images": bool(c.render_images), "board": ["".join(_BOARD_CH[c.board[y][x]] for x in range(32)) for y in range(31)],
            "enemy_color": _col(c.enemy_color), "chase_time": c.chase_time, "rand": _rand(c.rand), "painted_color": _col(c.painted_color),
            "enemies": [_ai(c.enemies[i]) for i in range(c.n_enemies)], "start_lives": c.start_lives,
            "player_start": {"tx": c.player_start_tx, "ty": c.player_start_ty}, "start_jumps": c.start_jumps,
            "default_board_bugs": bool(c.default_board_bugs), "player_color": _col(c.player_color), "bg_color": _col(c.bg_color),
            "chase_score_bonus": c.chase_score_bonus, "unpainted_color": _col(c.unpainted_color)}


def ami_cfg_from_json(c, d):
    for k in ("box_bonus", "jump_time", "chase_time", "start_lives", "start_jumps", "chase_score_bonus"):
        setattr(c, k, int(d[k]))
    for k in ("inner_painted_color", "enemy_color", "painted_color", "player_color", "bg_color", "unpainted_color"):
        _setcol(getattr(c, k), d[k])
    c.render_images = bool(d["render_images"]); c.default_board_bugs = bool(d["default_board_bugs"])
    if len(d["board"]) != 31 or any(len(r) != 32 for r in d["board"]) or len(d["enemies"]) > 8:
        raise ValueError("unsupported board/enemies")
    inv = {v: k for k, v in _BOARD_CH.items()}
    for y in range(31):
        for x in range(32):
            c.board[y][x] = inv[d["board"][y][x]]
    c.n_enemies = len(d["enemies"])
    for i, e in enumerate(d["enemies"]):
        _setai(c.enemies[i], e)
    c.player_start_tx, c.player_start_ty = int(d["player_start"]["tx"]), int(d["player_start"]["ty"])
    _setrand(c.rand, d["rand"])


CODEC = {"breakout": (brk_state_to_json, brk_state_from_json, brk_cfg_to_json, brk_cfg_from_json),
         "amidar": (ami_state_to_json, ami_state_from_json, ami_cfg_to_json, ami_cfg_from_json),
         "space_invaders": (si_state_to_json, si_state_from_json, si_cfg_to_json, si_cfg_from_json)}
NEW_GAMES_AT_CTOR = {"breakout": 2, "amidar": 1, "space_invaders": 2}    # [FIX] SURVEY App. A.2 lineages


class Input:
    """ctoybox.Input look-alike (toybox/envs/atari/constants.py:3-13 reads the class constants)."""
    _LEFT, _RIGHT, _UP, _DOWN, _BUTTON1, _BUTTON2, _NOOP = "left", "right", "up", "down", "button1", "button2", "noop"

    def __init__(self):
        self.left = self.right = self.up = self.down = self.button1 = self.button2 = False

    def mask(self):
        return (1 * bool(self.left) | 2 * bool(self.right) | 4 * bool(self.up) | 8 * bool(self.down)
                | 16 * bool(self.button1) | 32 * bool(self.button2))


class OracleToybox:
    """Single-env oracle with ctoybox.Toybox's consumed surface."""

    def __init__(self, game_name, grayscale=True, frameskip=0, seed=None):
        self.game_name = game_name
        self.grayscale = grayscale
        self.L = lib()
        self.g = GAMES[game_name]
        self.p = PREFIX[game_name]
        self.cfg = CFG[game_name]()
        self.state = STATE[game_name]()
        getattr(self.L, "tbo_%s_default_cfg" % self.p)(C.byref(self.cfg))
        if seed is not None:
            self.L.tbo_rng_seed(C.byref(self.cfg.rand), C.c_uint32(seed))
        for _ in range(NEW_GAMES_AT_CTOR[game_name]):
            self.new_game()
        self.rstate = self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def new_game(self):
        getattr(self.L, "tbo_%s_new_game" % self.p)(C.byref(self.cfg), C.byref(self.state))

    def set_seed(self, seed):
        self.L.tbo_rng_seed(C.byref(self.cfg.rand), C.c_uint32(int(seed) & 0xFFFFFFFF))

    def get_width(self):
        return DIMS[self.game_name][0]

    def get_height(self):
        return DIMS[self.game_name][1]

    def get_legal_action_set(self):
        return list(LEGAL[self.game_name])

    def apply_ale_action(self, a):
        m = self.L.tbo_ale_action_to_input(int(a))
        if m < 0:
            raise ValueError("Expected to apply action, but failed: {0}".format(a))
        getattr(self.L, "tbo_%s_step" % self.p)(C.byref(self.cfg), C.byref(self.state), m)

    def apply_action(self, inp):
        getattr(self.L, "tbo_%s_step" % self.p)(C.byref(self.cfg), C.byref(self.state), inp.mask())

    def _rgba(self):
        w, h = DIMS[self.game_name]
        buf = np.empty((h, w, 4), np.uint8)
        getattr(self.L, "tbo_%s_render" % self.p)(C.byref(self.cfg), C.byref(self.state), buf.ctypes.data_as(C.c_void_p))
        return buf

    def frame(self, mode, out_w=84, out_h=84):
        w, h = DIMS[self.game_name]
        shape = {"rgba": (h, w, 4), "rgb": (h, w, 3), "gray": (h, w, 1), "gray84": (out_h, out_w, 1)}[mode]
        out = np.empty(shape, np.uint8)
        rgba = self._rgba()
        self.L.tbo_frame_convert(self.g, rgba.ctypes.data_as(C.c_void_p), MODES[mode], out.ctypes.data_as(C.c_void_p), out_w, out_h)
        return out

    def get_state(self):
        return self.frame("gray") if self.grayscale else self.frame("rgba")

    def get_rgb_frame(self):
        return self.frame("rgb")

    def get_score(self):
        return int(self.state.score)

    def get_lives(self):
        return int(self.state.lives)

    def get_level(self):
        return int(self.state.level)

    def game_over(self):
        return self.state.lives <= 0

    def to_state_json(self):
        return CODEC[self.game_name][0](self.state)

    state_to_json = to_state_json

    def write_state_json(self, js):
        new = STATE[self.game_name]()
        CODEC[self.game_name][1](new, js)
        self.state = new

    def config_to_json(self):
        return CODEC[self.game_name][2](self.cfg)

    def write_config_json(self, js):
        CODEC[self.game_name][3](self.cfg, js)

    def query_state_json(self, query, args="null"):
        if self.game_name == "amidar" and query == "tile_to_world":
            return [args["tx"] * 64, args["ty"] * 80]
        if self.game_name == "amidar" and query == "world_to_tile":
            return [args["x"] // 64, args["y"] // 80]
        if self.game_name == "breakout" and query == "bricks_remaining":
            return sum(1 for i in range(self.state.n_bricks) if self.state.bricks[i].alive)
        raise ValueError("unknown query %s" % query)


class OracleBatch:
    """n independent oracle envs with the product's env-level semantics (auto-reset etc.)."""

    def __init__(self, game_name, n, seeds=None, cfg_json=None):
        self.game_name, self.n = game_name, n
        self.L = lib(); self.g = GAMES[game_name]; self.p = PREFIX[game_name]
        self.cfg = CFG[game_name]()
        getattr(self.L, "tbo_%s_default_cfg" % self.p)(C.byref(self.cfg))
        if cfg_json is not None:
            CODEC[game_name][3](self.cfg, cfg_json)
        self.states = (STATE[game_name] * n)()
        self.sim = (Rng * n)()
        for i in range(n):
            self.sim[i] = self.cfg.rand
        self.prev_score = np.zeros(n, np.int32)
        if seeds is not None:
            self.seed(seeds)
        else:
            for _ in range(NEW_GAMES_AT_CTOR[game_name]):
                self.new_game()

    def seed(self, seeds):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        self.L.tbo_batch_seed(self.sim, seeds.ctypes.data_as(C.c_void_p), self.n)
        self.new_game()

    def new_game(self):
        self.L.tbo_batch_new_game(self.g, C.byref(self.cfg), self.states, self.sim, self.n)
        self.prev_score[:] = 0

    def step(self, ale_actions, auto_reset=True):
        a = np.ascontiguousarray(ale_actions, np.int32)
        reward = np.empty(self.n, np.int32); done = np.empty(self.n, np.uint8)
        score = np.empty(self.n, np.int32); lives = np.empty(self.n, np.int32)
        rc = self.L.tbo_batch_step(self.g, C.byref(self.cfg), self.states, self.sim, self.n, a.ctypes.data_as(C.c_void_p),
                                   int(auto_reset), self.prev_score.ctypes.data_as(C.c_void_p), reward.ctypes.data_as(C.c_void_p),
                                   done.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p), lives.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise ValueError("invalid ALE action id")
        return reward, done.astype(bool), score, lives

    def render(self, mode, out_w=84, out_h=84):
        w, h = DIMS[self.game_name]
        shape = {"rgba": (h, w, 4), "rgb": (h, w, 3), "gray": (h, w), "gray84": (out_h, out_w)}[mode]
        out = np.empty((self.n,) + shape, np.uint8)
        self.L.tbo_batch_render(self.g, C.byref(self.cfg), self.states, self.n, MODES[mode], out.ctypes.data_as(C.c_void_p), out_w, out_h)
        return out

    def state_json(self, i):
        return CODEC[self.game_name][0](self.states[i])

    def write_state_json(self, i, js):
        new = STATE[self.game_name]()
        CODEC[self.game_name][1](new, js)
        self.states[i] = new
        self.prev_score[i] = new.score

    def rollout(self, steps, t0, action_seed, env0=0, render_mode=None):
        legal = np.asarray(LEGAL[self.game_name], np.int32)
        frames = None
        mode = -1
        if render_mode is not None:
            mode = MODES[render_mode]
            frames = np.empty(self.n * self.L.tbo_frame_bytes(self.g, mode, 84, 84), np.uint8)
        eps = C.c_int64(0)
        sec = self.L.tbo_batch_rollout(self.g, C.addressof(self.cfg), C.addressof(self.states), C.addressof(self.sim), self.n, steps, t0,
                                       action_seed, env0, legal.ctypes.data, len(legal), mode,
                                       frames.ctypes.data if frames is not None else None, 84, 84,
                                       self.prev_score.ctypes.data, C.addressof(eps))
        return sec, eps.value, frames


def action_index(seed, env, t, n_legal):
    return lib().tbo_action_index(seed, env, t, n_legal)
