"""Minimal behavioural stub of gym 0.18 (absent from this image) -- TEST INFRASTRUCTURE ONLY.

Just enough surface for the unmodified reference (kevslinger/DTQN, /root/reference) to import and
run its CarFlag / Memory hot path so golden vectors can be generated from it.
"""
from gym.core import Env, Wrapper
from gym import spaces, error, envs, wrappers, utils
from gym.envs.registration import make, register

__version__ = "0.18.0-stub"
