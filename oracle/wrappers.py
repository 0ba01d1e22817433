"""TEST INFRASTRUCTURE (like everything under oracle/): a one-env-at-a-time CPU restatement of the reference's wrapper
chain over the oracle engine, the checker for toybox_b200.wrappers / tbx_wrap_step.  Each class follows the class of
the same name in baselines/baselines/common/atari_wrappers.py (line ranges cited per class); the env underneath
follows ToyboxBaseEnv (toybox/envs/atari/base.py:115-156).  gym itself is not importable here, so the classes are
plain Python with the same step/reset protocol.

Parity status: pinned to the reference by construction only (the wrappers are short Python; no reference test runs
them on Toybox envs); the no-op count uses the library's counter-based generator instead of gym's np_random and gym's
TimeLimit is not applied."""
import ctypes as C
from collections import deque

import numpy as np

from . import oracle as O


class BaseEnv:
    """ToyboxBaseEnv (toybox/envs/atari/base.py:37-160) over one oracle env, grayscale observations."""

    def __init__(self, game, seed):
        self.b = O.OracleBatch(game, 1, seeds=np.asarray([seed], np.uint32))
        self.legal = list(O.LEGAL[game])
        self.game = game

    def _obs(self):
        return self.b.render("gray")[0]

    def step(self, action_index):                       # base.py:115-149
        r, d, s, l = self.b.step(np.asarray([self.legal[action_index]], np.int32), auto_reset=False)
        return self._obs(), int(r[0]), bool(d[0]), {"lives": int(l[0]), "score": int(s[0])}

    def reset(self):                                    # base.py:151-156
        self.b.new_game()
        return self._obs()

    def lives(self):
        return int(self.b.states[0].lives)

    def score(self):
        return int(self.b.states[0].score)


class NoopResetEnv:                                     # atari_wrappers.py:107-134
    def __init__(self, env, noop_max, noop_seed, env_id):
        self.env, self.noop_max, self.noop_seed, self.env_id, self.resets = env, noop_max, noop_seed, env_id, 0

    def reset(self):
        obs = self.env.reset()
        if self.noop_max > 0:
            noops = 1 + O.action_index(self.noop_seed, self.env_id, self.resets, self.noop_max)
            for _ in range(noops):
                obs, _, done, _ = self.env.step(0)
                if done:
                    obs = self.env.reset()
        self.resets += 1
        return obs

    def step(self, ac):
        return self.env.step(ac)


class MaxAndSkipEnv:                                    # atari_wrappers.py:186-209
    def __init__(self, env, skip, shape):
        self.env, self._skip = env, skip
        self._obs_buffer = np.zeros((2,) + shape, dtype=np.uint8)
        self.fresh = True                               # False when the buffer was not (fully) refreshed by this step

    def step(self, action):
        total_reward, done, info = 0, None, None
        seen = 0
        for i in range(self._skip):
            obs, reward, done, info = self.env.step(action)
            if i == self._skip - 2:
                self._obs_buffer[0] = obs
                seen |= 1
            if i == self._skip - 1:
                self._obs_buffer[1] = obs
                seen |= 2
            total_reward += reward
            if done:
                break
        self.fresh = seen == (3 if self._skip > 1 else 2)
        if self._skip == 1:
            self._obs_buffer[0] = self._obs_buffer[1]
        return self._obs_buffer.max(axis=0), total_reward, done, info

    def reset(self):
        return self.env.reset()


class Monitor:                                          # baselines/baselines/bench/monitor.py:36-76 (between make_atari and wrap_deepmind,
    def __init__(self, env):                            # common/cmd_util.py:30-36, run.py:116-119)
        self.env, self.rewards, self.needs_reset = env, None, True
        self.episode_rewards, self.episode_lengths = [], []

    def reset(self):
        self.rewards, self.needs_reset = [], False
        return self.env.reset()

    def step(self, action):
        assert not self.needs_reset, "Tried to step environment that needs reset"
        ob, rew, done, info = self.env.step(action)
        self.rewards.append(rew)
        if done:
            self.needs_reset = True
            info = dict(info)
            info["episode"] = {"r": round(sum(self.rewards), 6), "l": len(self.rewards)}
            self.episode_rewards.append(sum(self.rewards))
            self.episode_lengths.append(len(self.rewards))
        return ob, rew, done, info


class EpisodicLifeEnv:                                  # atari_wrappers.py:153-184
    def __init__(self, env, base):
        self.env, self.base, self.lives, self.was_real_done = env, base, 0, True

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.was_real_done = done
        lives = self.base.lives()
        if lives < self.lives and lives > 0:
            done = True
        self.lives = lives
        return obs, reward, done, info

    def reset(self):
        if self.was_real_done:
            obs = self.env.reset()
        else:
            obs, _, _, _ = self.env.step(0)
        self.lives = self.base.lives()
        return obs


class FireResetEnv:                                     # atari_wrappers.py:136-151
    def __init__(self, env):
        self.env = env

    def reset(self):
        self.env.reset()
        obs, _, done, _ = self.env.step(1)
        if done:
            self.env.reset()
        obs, _, done, _ = self.env.step(2)
        if done:
            self.env.reset()
        return obs

    def step(self, ac):
        return self.env.step(ac)


def warp(frame, w, h):                                  # WarpFrame.observation, atari_wrappers.py:238-244 (Toybox branch: no cvtColor)
    src = np.ascontiguousarray(frame.reshape(frame.shape[0], frame.shape[1]), np.uint8)
    dst = np.empty((h, w), np.uint8)
    O.lib().tbo_resize_area_u8(src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], 1, dst.ctypes.data_as(C.c_void_p), w, h)
    return dst


class WrappedEnv:
    """wrap_deepmind(make_atari(env), frame_stack=True) (atari_wrappers.py:323-360) with the VecEnv worker's
    reset-on-done (vec_env/subproc_vec_env.py:11-15).  step() returns (frames oldest-first [k,h,w], reward, done, info)."""

    def __init__(self, game, seed, env_id=0, frame_skip=4, noop_max=30, episode_life=True, fire_reset=True, clip_rewards=True,
                 frame_stack=4, size=(84, 84), noop_seed=0, stack_reset="fill"):
        """stack_reset "fill": wrap_deepmind(frame_stack=True) -- FrameStack (:246-275) owns the stack; "zero": wrap_deepmind()
        under VecFrameStack (vec_env/vec_frame_stack.py:17-30) -- the stack of a finished env is zeroed, then holds the reset
        observation only."""
        self.base = BaseEnv(game, seed)
        dims = O.DIMS[game]
        env = NoopResetEnv(self.base, noop_max, noop_seed, env_id) if noop_max > 0 else _PlainReset(self.base)
        self.skipper = env = MaxAndSkipEnv(env, frame_skip, (dims[1], dims[0]))
        self.monitor = env = Monitor(env)
        self.stack_reset = stack_reset
        if episode_life:
            env = EpisodicLifeEnv(env, self.base)
        if fire_reset and len(self.base.legal) >= 3:
            env = FireResetEnv(env)
        self.env, self.clip, self.k, self.size = env, clip_rewards, frame_stack, size
        self.frames = deque([], maxlen=frame_stack)

    def _warp(self, obs):
        return warp(obs, self.size[0], self.size[1])

    def reset(self):                                    # FrameStack.reset, :262-266 / VecFrameStack.reset, vec_frame_stack.py:27-31
        ob = self._warp(self.env.reset())
        for _ in range(self.k):
            self.frames.append(np.zeros_like(ob) if self.stack_reset == "zero" else ob)
        self.frames.append(ob)
        return np.stack(self.frames)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        fresh = self.skipper.fresh
        score, lives = self.base.score(), self.base.lives()
        real_done = self.base.lives() <= 0
        if self.clip:
            reward = int(np.sign(reward))               # ClipRewardEnv, :211-218
        self.frames.append(self._warp(obs))             # FrameStack.step, :268-271
        if done:
            stacked = self.reset()                      # the VecEnv worker: "if done: ob = env.reset()"
        else:
            stacked = np.stack(self.frames)
        out = {"score": score, "lives": lives, "real_done": real_done, "fresh": fresh}
        if "episode" in info:
            out["episode"] = info["episode"]
        return stacked, reward, done, out


class _PlainReset:
    def __init__(self, env):
        self.env = env

    def reset(self):
        return self.env.reset()

    def step(self, ac):
        return self.env.step(ac)
