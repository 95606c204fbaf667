"""Importable alias for the package directory ``gym-swarm_b200/`` (a hyphen is
not a valid Python identifier).  ``import gym_swarm_b200`` loads the real package
from ``../gym-swarm_b200`` and registers it under this name, so that
``gym_swarm_b200.envs`` etc. resolve into that directory."""
import importlib.util as _u
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gym-swarm_b200")
_spec = _u.spec_from_file_location("gym_swarm_b200", _os.path.join(_root, "__init__.py"),
                                   submodule_search_locations=[_root])
_mod = _u.module_from_spec(_spec)
_sys.modules["gym_swarm_b200"] = _mod
_spec.loader.exec_module(_mod)
