"""TEST INFRASTRUCTURE ONLY -- the synthetic checkpoint / batch generators moved to synth/ckpt.py (data only, shared with
bench.py and smoke()); this module keeps the old import path for the golden generators and the tests."""
from synth.ckpt import *  # noqa: F401,F403
from synth.ckpt import _rng, _adds_into_a_sum  # noqa: F401
