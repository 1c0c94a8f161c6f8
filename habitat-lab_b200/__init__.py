"""hb200 -- B200-native DD-PPO learner hot path behind habitat-baselines' registry API.

The compute path is libhb200.so (hand-written sm_100a CUDA, C ABI in include/hb200.h);
this package is the thin Python host side that mirrors the reference's
Policy / Updater / Storage interfaces.  There is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401
from ._lib import Hb200Error, load  # noqa: F401

__version__ = "0.1.0"


def smoke() -> None:
    from .smoke import run
    run()
