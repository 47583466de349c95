"""Importable alias of the product package directory `mpc-ilqr-mujoco_b200/` (a hyphen is not a legal
Python identifier, so this shim extends its search path to that directory)."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mpc-ilqr-mujoco_b200"))
from .api import *  # noqa: F401,F403,E402
