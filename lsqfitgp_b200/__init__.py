"""lsqfitgp_b200: B200-native (sm_100a) GP-fitting hot path with the lsqfitgp API."""

__version__ = '0.1.0'

from . import _lib  # noqa: F401
