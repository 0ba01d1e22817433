"""The callers either side of the hot path, at batch size N on one GPU.

`DeepmindToybox` is the reference's wrapper stack
    make_atari:    MaxAndSkipEnv(NoopResetEnv(env, 30), 4)             baselines/baselines/common/atari_wrappers.py:323-333
    wrap_deepmind: FrameStack(ClipRewardEnv(WarpFrame(FireResetEnv(EpisodicLifeEnv(env)))), 4)          :345-360
fused into two kernel launches per agent step (tbx_wrap_step: csrc/tbx_wrap.cuh + the dual mode of
csrc/tbx_render_area.cuh).  `ToyboxVecEnv` gives it the VecEnv surface baselines' trainers consume
(baselines/baselines/common/vec_env/__init__.py:26-131: reset / step_async / step_wait / step / close, num_envs,
observation_space, action_space; auto-reset and `infos[i]['episode']` as vec_env/subproc_vec_env.py:11-15 +
bench/monitor.py:58-76 provide them).
"""
import ctypes as C
import json
import os
import time

import numpy as np
import torch

from . import _lib
from .pool import BatchedToybox, _ptr, _stream


class DeepmindToybox:
    """N wrapped envs.  step(actions) takes gym action INDICES (0..len(legal)-1, envs/atari/base.py:126) as an int32
    tensor on the pool's device and returns (obs, reward, done, info):
      obs    uint8[N, k, out_h, out_w]: a RING of the last k observations per env -- the newest is obs[:, self.slot],
             FrameStack order (oldest first) is slots (self.slot + 1 .. self.slot + k) % k; `stacked()` returns that
             order as the reference's uint8[N, out_h, out_w, k] (a copy);
      reward int32[N] (sign of the summed reward when clip_rewards), done uint8[N] (game over, or a lost life when
      episode_life), info = {'lives', 'score', 'real_done'} tensors (values before the reset of a finished env).
    A finished env is reset inside the same call and `obs` then holds its reset observation (VecEnv semantics)."""

    def __init__(self, game, n_envs, device=None, seeds=None, config=None, frame_skip=4, noop_max=30, episode_life=True,
                 fire_reset=True, clip_rewards=True, frame_stack=4, size=(84, 84), noop_seed=0, env0=0, stack_reset="fill"):
        """stack_reset: what a reset observation does to the other k-1 ring slots -- "fill" (FrameStack.reset,
        atari_wrappers.py:262-266: wrap_deepmind(frame_stack=True), the deepq path) or "zero" (VecFrameStack,
        vec_env/vec_frame_stack.py:17-30: make_vec_env + VecFrameStack, the ppo2 / a2c path of run.py:123-125)."""
        if stack_reset not in ("fill", "zero"):
            raise ValueError("stack_reset is 'fill' (FrameStack) or 'zero' (VecFrameStack)")
        self.pool = BatchedToybox(game, n_envs, device=device, obs=("gray_area", size[0], size[1]), config=config, seeds=seeds)
        self.L = self.pool.L
        self.n_envs, self.device = self.pool.n_envs, self.pool.device
        self.k, self.out_w, self.out_h = int(frame_stack), int(size[0]), int(size[1])
        h = C.c_void_p()
        _lib.check(self.L.tbx_wrap_create(self.pool._h, int(frame_skip), int(noop_max), int(bool(episode_life)), int(bool(fire_reset)),
                                          int(bool(clip_rewards)), self.k, self.out_w, self.out_h, int(noop_seed), int(env0), C.byref(h)))
        self._w = h
        _lib.check(self.L.tbx_wrap_set_stack_mode(self._w, 1 if stack_reset == "zero" else 0))
        n, dev = self.n_envs, self.device
        # Monitor's episode record (bench/monitor.py:58-76) of the episode that ended in the last agent step, where real_done
        self.ep_return = torch.zeros(n_envs, dtype=torch.int32, device=self.device)
        self.ep_length = torch.zeros(n_envs, dtype=torch.int32, device=self.device)
        _lib.check(self.L.tbx_wrap_set_episode_outputs(self._w, _ptr(self.ep_return), _ptr(self.ep_length)))
        self.obs = torch.zeros((n, self.k, self.out_h, self.out_w), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros(n, dtype=torch.int32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.real_done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.score = torch.zeros(n, dtype=torch.int32, device=dev)
        self.lives = torch.zeros(n, dtype=torch.int32, device=dev)
        self.slot = 0
        self.n_actions = len(self.pool.get_legal_action_set())

    def close(self):
        if getattr(self, "_w", None):
            self.L.tbx_wrap_destroy(self._w)
            self._w = None
        self.pool.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, actions, render=True):
        slot = C.c_int(0)
        _lib.check(self.L.tbx_wrap_step(self._w, _ptr(actions), _ptr(self.obs) if render else None, _ptr(self.reward), _ptr(self.done),
                                        _ptr(self.real_done), _ptr(self.score), _ptr(self.lives), C.byref(slot), _stream(self.device)))
        self.slot = slot.value

    def reset(self):
        """env.reset() of every wrapped env (a new game only where the last one is over, as EpisodicLifeEnv.reset)."""
        self._call(None)
        return self.obs

    def step(self, actions, render=True):
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.ascontiguousarray(actions, dtype=np.int32), device=self.device)
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous() or actions.numel() != self.n_envs:
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
            if actions.numel() != self.n_envs:
                raise ValueError("expected %d actions" % self.n_envs)
        self._call(actions, render)
        return self.obs, self.reward, self.done, {"lives": self.lives, "score": self.score, "real_done": self.real_done,
                                                    "ep_return": self.ep_return, "ep_length": self.ep_length}

    def check(self):
        self.pool.check()

    def order(self):
        """ring slots, oldest observation first"""
        return [(self.slot + 1 + i) % self.k for i in range(self.k)]

    def stacked(self):
        """FrameStack's observation (atari_wrappers.py:246-275) for every env: uint8[N, out_h, out_w, k], oldest first."""
        idx = torch.as_tensor(self.order(), device=self.device)
        return self.obs.index_select(1, idx).permute(0, 2, 3, 1).contiguous()


class _Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class _Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)


class ToyboxVecEnv:
    """The VecEnv surface of baselines (vec_env/__init__.py:26-131) over a DeepmindToybox: numpy in, numpy out, one
    process, one GPU.  It stands for `VecFrameStack(make_vec_env(env_id, 'atari', n, seed), 4)` (run.py:123-125): the stack
    of a finished env is zeroed and holds its reset observation only (vec_frame_stack.py:17-30), and
    `infos[i]['episode'] = {'r', 'l', 't'}` appears when env i's game ends, exactly as the Monitor between make_atari and
    wrap_deepmind reports it (bench/monitor.py:58-76: r = sum of raw rewards, l = MaxAndSkipEnv steps incl. those of
    FireResetEnv / EpisodicLifeEnv resets, t = seconds since start) -- both counters are kept by the step kernel.
    monitor_file: path of a Monitor-compatible CSV (bench/monitor.py:22-28,70-71,96-126: a '#{json header}' line, then the
    columns r,l,t), one row per finished episode."""

    def __init__(self, game, num_envs, monitor_file=None, **kw):
        kw.setdefault("stack_reset", "zero")
        self.env = DeepmindToybox(game, num_envs, **kw)
        self.num_envs = int(num_envs)
        e = self.env
        self.observation_space = _Box(0, 255, (e.out_h, e.out_w, e.k), np.uint8)
        self.action_space = _Discrete(e.n_actions)
        self._actions = None
        self._t0 = time.time()
        self.closed = False
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []
        self._csv = None
        if monitor_file is not None:
            if not monitor_file.endswith("monitor.csv"):
                monitor_file = monitor_file + ".monitor.csv" if not os.path.isdir(monitor_file) else os.path.join(monitor_file, "monitor.csv")
            self._csv = open(monitor_file, "wt")
            self._csv.write("#%s\n" % json.dumps({"t_start": self._t0, "env_id": "%sToyboxNoFrameskip-v4" % game.title().replace("_", "")}))
            self._csv.write("r,l,t\n")
            self._csv.flush()

    def reset(self):
        self.env.reset()
        return self.env.stacked().cpu().numpy()

    def step_async(self, actions):
        self._actions = np.ascontiguousarray(actions, dtype=np.int32)

    def step_wait(self):
        e = self.env
        e.step(self._actions)
        obs = e.stacked().cpu().numpy()
        rew = e.reward.cpu().numpy().astype(np.float32)
        done = e.done.cpu().numpy().astype(bool)
        real = e.real_done.cpu().numpy().astype(bool)
        score = e.score.cpu().numpy().astype(np.int64)
        lives = e.lives.cpu().numpy()
        infos = [{"lives": int(lives[i]), "score": int(score[i])} for i in range(self.num_envs)]
        if real.any():
            ep_r, ep_l = e.ep_return.cpu().numpy(), e.ep_length.cpu().numpy()
            t = round(time.time() - self._t0, 6)
            for i in np.flatnonzero(real):
                ep = {"r": round(float(ep_r[i]), 6), "l": int(ep_l[i]), "t": t}
                infos[i]["episode"] = ep
                self.episode_rewards.append(ep["r"])
                self.episode_lengths.append(ep["l"])
                self.episode_times.append(t)
                if self._csv is not None:
                    self._csv.write("%s,%d,%s\n" % (ep["r"], ep["l"], ep["t"]))
            if self._csv is not None:
                self._csv.flush()
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if not self.closed:
            self.env.close()
            self.closed = True
            if self._csv is not None:
                self._csv.close()

    def get_episode_rewards(self):           # the Monitor accessors of bench/monitor.py:84-94
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths

    def get_episode_times(self):
        return self.episode_times

    def get_images(self):
        return self.env.pool.render(obs="rgb").cpu().numpy()

    def render(self, mode="rgb_array"):
        return self.get_images()

    @property
    def unwrapped(self):
        return self
