"""gym/ALE-mocking environment surface of the reference (toybox/envs/atari/*) on the B200 library."""
from .atari import (ACTION_LOOKUP, ACTION_MEANING, AmidarEnv, BatchedToyboxEnv, BreakoutEnv, MockALE,  # noqa: F401
                    SpaceInvadersEnv, ToyboxBaseEnv, make)
