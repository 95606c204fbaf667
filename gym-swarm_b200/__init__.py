"""swarm-b200: B200-native batched implementation of the gym-swarm
``DiscreteSwarmEnv`` transition (reference: gym_swarm/envs/swarm_discrete_env.py).

Only the hot path lives here: ``csrc/`` (sm_100a CUDA kernels + the C-ABI),
``_lib`` (ctypes loader), ``envs`` (host mirror of the reference Gym surface) and
``tapes`` (explicit random streams)."""
from . import tapes  # noqa: F401

__all__ = ["tapes"]
