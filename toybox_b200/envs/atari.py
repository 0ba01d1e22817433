"""The reference's environment layer (toybox/envs/atari/base.py:15-173, breakout.py, amidar.py, space_invaders.py,
constants.py:16-37) on the B200 library, in two shapes:

* `ToyboxBaseEnv` / `BreakoutEnv` / `AmidarEnv` / `SpaceInvadersEnv`: one environment with the reference's exact
  step/reset/seed/render semantics and `MockALE` (plus `getScreenRGB`, which north_star names).  gym is optional:
  when it is importable the classes are `gym.Env`s and the three ids of toybox/__init__.py:8-24 are registered.
* `BatchedToyboxEnv`: the same semantics at batch N on device tensors -- the vectorised driver that replaces the
  reference's process-per-env SubprocVecEnv (auto-reset on done, subproc_vec_env.py:11-15).
"""
import hashlib

import numpy as np
import torch

from ..ctoybox import Toybox
from ..pool import BatchedToybox

try:                                   # optional
    import gym
    _EnvBase = gym.Env
except Exception:                      # pragma: no cover - gym is not in this image
    gym = None
    _EnvBase = object

ACTION_MEANING = {
    0: "NOOP", 1: "FIRE", 2: "UP", 3: "RIGHT", 4: "LEFT", 5: "DOWN", 6: "UPRIGHT", 7: "UPLEFT", 8: "DOWNRIGHT", 9: "DOWNLEFT",
    10: "UPFIRE", 11: "RIGHTFIRE", 12: "LEFTFIRE", 13: "DOWNFIRE", 14: "UPRIGHTFIRE", 15: "UPLEFTFIRE", 16: "DOWNRIGHTFIRE",
    17: "DOWNLEFTFIRE",
}
ACTION_LOOKUP = {v: k for (k, v) in ACTION_MEANING.items()}


def hash_seed(seed, max_bytes=8):
    """gym.utils.seeding.hash_seed: decorrelates nearby seeds (used by ToyboxBaseEnv.seed, base.py:94)."""
    h = hashlib.sha512(str(seed).encode("utf8")).digest()
    return int.from_bytes(h[:max_bytes], "little")


class Discrete:
    def __init__(self, n):
        self.n = n
        self._rng = np.random.RandomState()

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class MockALE:
    """toybox/envs/atari/base.py:15-35, plus the screen getters of the real ALE interface."""

    def __init__(self, toybox):
        self.toybox = toybox

    def lives(self):
        return self.toybox.get_lives()

    def get_score(self):
        return self.toybox.get_score()

    def game_over(self):
        return self.toybox.get_lives() <= 0

    def getScreenRGB(self):
        return self.toybox.get_rgb_frame()

    def getScreenGrayscale(self):
        g = self.toybox.grayscale
        self.toybox.grayscale = True
        try:
            return self.toybox.get_state()
        finally:
            self.toybox.grayscale = g

    def saveScreenPNG(self, name):
        name = name.decode("utf-8") if isinstance(name, bytes) else name
        grayscale = self.toybox.grayscale
        self.toybox.grayscale = False
        self.toybox.save_frame_image(name)
        self.toybox.grayscale = grayscale


class ToyboxBaseEnv(_EnvBase):
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, toybox, game, frameskip=(2, 5), repeat_action_probability=0., grayscale=True, alpha=False, actions=None):
        assert toybox.rstate
        self.toybox = toybox
        self.game = game
        self.cached_state = None
        self.score = self.toybox.get_score()
        self.viewer = None
        self._np_random = None
        self.ale = MockALE(toybox)
        if actions is None:
            actions = toybox.get_legal_action_set()
        self._action_set = list(actions)
        self._obs_type = "image"
        self._rgba = 1 if grayscale else 4 if alpha else 3
        self._height = self.toybox.get_height()
        self._width = self.toybox.get_width()
        self._dim = (self._height, self._width, self._rgba)
        self.reward_range = (0, float("inf"))
        spaces = gym.spaces if gym is not None else None
        self.action_space = spaces.Discrete(len(self._action_set)) if spaces else Discrete(len(self._action_set))
        self.observation_space = (spaces.Box(low=0, high=255, shape=self._dim, dtype="uint8") if spaces
                                  else Box(0, 255, self._dim, "uint8"))

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        seed1 = int(seed) if seed is not None else int.from_bytes(np.random.bytes(8), "little")
        self._np_random = np.random.RandomState(hash_seed(seed1) % 2 ** 32)
        seed2 = hash_seed(seed1 + 1) % 2 ** 31
        self.toybox.set_seed(seed2)
        self.toybox.new_game()          # start a new game so the seed gets used (base.py:96-97)
        return [seed1, seed2]

    def get_action_meanings(self):
        return list(ACTION_MEANING.values())

    def _get_obs(self):
        obs = self.toybox.get_state()
        if self._rgba == 3:
            obs = obs[:, :, :-1]
        return obs

    def step(self, action_index):
        info = {}
        assert action_index < len(self._action_set)
        self.toybox.apply_ale_action(self._action_set[action_index])
        if self.ale.game_over():
            info["cached_state"] = self.toybox.to_state_json()
        obs = self._get_obs()
        score = self.toybox.get_score()
        reward = max(score - self.score, 0)
        self.score = score
        done = self.ale.game_over()
        info["lives"] = self.toybox.get_lives()
        info["score"] = 0 if done else self.score
        return obs, reward, done, info

    def reset(self):
        self.cached_state = self.toybox.to_state_json()
        self.toybox.new_game()
        self.score = self.toybox.get_score()
        return self._get_obs()

    def render(self, mode="human", close=False):
        if mode == "rgb_array":
            return self.toybox.get_rgb_frame()
        if self.viewer is None and gym is not None:
            from gym.envs.classic_control.rendering import SimpleImageViewer
            self.viewer = SimpleImageViewer()
        if self.viewer is not None:
            self.viewer.imshow(self.toybox.get_rgb_frame())
            return self.viewer.isopen
        return False

    def close(self):
        if self.viewer is not None:
            self.viewer.close()
        if self.toybox is not None:
            self.toybox.close()
        self.toybox = None


class BreakoutEnv(ToyboxBaseEnv):
    def __init__(self, frameskip=(2, 5), repeat_action_probability=0., grayscale=True, alpha=False, device=None):
        super().__init__(Toybox("breakout", grayscale, device=device), "breakout", frameskip, repeat_action_probability,
                         grayscale=grayscale, alpha=alpha)


class AmidarEnv(ToyboxBaseEnv):
    def __init__(self, frameskip=(2, 5), repeat_action_probability=0., grayscale=True, alpha=False, device=None):
        super().__init__(Toybox("amidar", grayscale, device=device), "amidar", frameskip, repeat_action_probability,
                         grayscale=grayscale, alpha=alpha)


class SpaceInvadersEnv(ToyboxBaseEnv):
    def __init__(self, frameskip=(2, 5), repeat_action_probability=0., grayscale=True, alpha=False, device=None):
        super().__init__(Toybox("space_invaders", grayscale, device=device), "space_invaders", frameskip, repeat_action_probability,
                         grayscale=grayscale, alpha=alpha)


ENV_IDS = {"BreakoutToyboxNoFrameskip-v4": BreakoutEnv, "AmidarToyboxNoFrameskip-v4": AmidarEnv,
           "SpaceInvadersToyboxNoFrameskip-v4": SpaceInvadersEnv}


def make(env_id, **kwargs):
    """gym.make for the three ids the reference registers (toybox/__init__.py:8-24)."""
    return ENV_IDS[env_id](**kwargs)


if gym is not None:                    # pragma: no cover
    try:
        from gym.envs.registration import register
        for _id, _cls in ENV_IDS.items():
            register(id=_id, entry_point="toybox_b200.envs.atari:%s" % _cls.__name__, nondeterministic=False)
    except Exception:
        pass


class BatchedToyboxEnv:
    """N environments stepped together on one GPU; tensors in, tensors out.

    step(action_indices) -> (obs uint8[N,*obs_shape], reward int32[N], done bool[N], info) with the reference's
    semantics per env (reward = max(score - previous score, 0); done = lives <= 0; info['score'] is 0 when done) and
    the vectorised driver's auto-reset: a finished env starts a new game inside the same call and its returned
    observation is the first frame of the new episode (subproc_vec_env.py:11-15).
    """

    def __init__(self, game, n_envs, device=None, obs="gray84", seed=None, auto_reset=True, config=None):
        self.toybox = BatchedToybox(game, n_envs, device=device, obs=obs, config=config)
        self.game, self.n_envs, self.auto_reset = game, n_envs, auto_reset
        self.device = self.toybox.device
        self._action_set = self.toybox.get_legal_action_set()
        self._action_lut = torch.tensor(self._action_set, dtype=torch.int32, device=self.device)
        self.action_space = Discrete(len(self._action_set))
        self.observation_space = Box(0, 255, self.toybox.obs_shape, "uint8")
        self.ale = self          # .ale.lives() etc. at batch N
        if seed is not None:
            self.seed(seed)

    # MockALE surface, batched
    def lives(self):
        return self.toybox.get_lives()

    def get_score(self):
        return self.toybox.get_score()

    def game_over(self):
        return self.toybox.game_over()

    def getScreenRGB(self):
        return self.toybox.get_rgb_frame()

    def get_action_meanings(self):
        return list(ACTION_MEANING.values())

    def seed(self, seed=None):
        """Env i gets ToyboxBaseEnv.seed(seed + i)'s derived simulator seed, then a new game."""
        seed1 = int(seed) if seed is not None else int.from_bytes(np.random.bytes(4), "little")
        seeds2 = np.asarray([hash_seed(seed1 + i + 1) % 2 ** 31 for i in range(self.n_envs)], dtype=np.uint32)
        self.toybox.set_seed(seeds2)
        self.toybox.new_game()
        return [seed1, seeds2]

    def reset(self):
        self.toybox.new_game()
        return self.toybox.render()

    def step(self, action_indices):
        idx = torch.as_tensor(action_indices, device=self.device).long()
        ale = self._action_lut[idx]
        self.toybox.apply_ale_action(ale, auto_reset=self.auto_reset)
        obs = self.toybox.render()
        done = self.toybox.done.bool()
        info = {"lives": self.toybox.lives, "score": torch.where(done, torch.zeros_like(self.toybox.score), self.toybox.score)}
        return obs, self.toybox.reward, done, info

    def episode_stats(self, reset=False):
        return self.toybox.episode_stats(reset)

    def close(self):
        self.toybox.close()
