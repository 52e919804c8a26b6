"""Minimal stand-in for gym 0.15 (TEST INFRASTRUCTURE: gym is not installed in the build image and is
not part of the step path).  Only what robogym's dactyl envs and wrappers touch."""
from . import spaces  # noqa: F401
from .core import ActionWrapper, Env, GoalEnv, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401
from . import utils, wrappers  # noqa: F401

__version__ = "0.15.3-stub"
