"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Executes the UNMODIFIED reference file
``/root/reference/gym_swarm/envs/swarm_discrete_env.py`` under an import shim
and feeds every ``np.random.*`` call from explicit tapes, so that the reference
and the CUDA path can be driven lock-step from the same random streams.

This module only works where ``/root/reference`` exists (the authoring
container).  It is used by ``tests/golden/make_golden.py`` to record the golden
trajectories that are committed under ``tests/golden/`` and by the (skipped when
absent) live differential tests.  Nothing from the reference is copied: the file
is loaded by path with ``importlib``.

Why a shim is needed: ``gym`` and ``matplotlib`` are not
installed and ``sklearn.neighbors.DistanceMetric`` moved to ``sklearn.metrics``.

Determinism wrapper
-------------------
The reference's module-global ``np`` is replaced by a proxy:

* ``np.random.randint`` / ``np.random.random`` are served from tapes.  The call
  sites are ``swarm_discrete_env.py:92,101`` (predator placement),
  ``:138`` (10 % retarget roll), ``:202,208`` (agent placement + re-draw rounds),
  ``:216`` (render-only orientation) and ``:250`` (collision re-draw actions).
* ``np.argsort`` is forced to ``kind='stable'`` -- the canonical tie rule for the
  nearest-neighbour target (``swarm_discrete_env.py:121-122``); NumPy's default
  sort is not stable on AVX-512 hosts, so the raw reference is platform dependent
  on ties.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as _np

# The unmodified reference file: where it lies in the authoring container, else the copy `pip install --target
# baseline/_ref` made of it (git-ignored, travels to the GPU box with the snapshot; __graft_entry__.build()).
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference/gym_swarm/envs/swarm_discrete_env.py",
               os.path.join(_REPO, "baseline", "_ref", "gym_swarm", "envs", "swarm_discrete_env.py"))
REFERENCE_FILE = next((c for c in _CANDIDATES if os.path.isfile(c)), _CANDIDATES[0])


def reference_available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


class TapeExhausted(RuntimeError):
    pass


class _TapeRandom:
    """Stands in for ``np.random`` inside the reference module."""

    def __init__(self):
        self.S = None
        self.N = None
        self.reset_feed = None   # dict(place=1d int array, ptape=(KP,2) int array, orient=(N,) int array)
        self.step_feed = None    # dict(slab=1d int array, roll=float)
        self.place_cur = 0
        self.pred_cur = 0
        self.slab_cur = 0
        self.orient_served = False
        self.phase = None

    # -- feeding -----------------------------------------------------------------
    def begin_reset(self, place, ptape, orient):
        self.phase = "reset"
        self.reset_feed = dict(place=_np.asarray(place, dtype=_np.int64),
                               ptape=_np.asarray(ptape, dtype=_np.int64).reshape(-1, 2),
                               orient=_np.asarray(orient, dtype=_np.int64))
        self.place_cur = 0
        self.pred_cur = 0
        self.orient_served = False

    def begin_step(self, slab, roll):
        self.phase = "step"
        self.step_feed = dict(slab=_np.asarray(slab, dtype=_np.int64), roll=float(roll))
        self.slab_cur = 0
        self.roll_served = False

    # -- the two functions the reference calls -----------------------------------
    def randint(self, high, size=None):
        if self.phase == "step":
            # swarm_discrete_env.py:250  np.random.randint(8, size=(1, len(ch_id)))
            assert high == 8 and isinstance(size, tuple) and size[0] == 1, (high, size)
            L = size[1]
            slab = self.step_feed["slab"]
            if self.slab_cur + L > slab.shape[0]:
                raise TapeExhausted("re-draw slab exhausted")
            out = slab[self.slab_cur:self.slab_cur + L].reshape(1, L).copy()
            self.slab_cur += L
            return out
        assert self.phase == "reset"
        if isinstance(size, tuple):
            # swarm_discrete_env.py:202-203 and :208-209  randint(S, size=(2, M))
            assert high == self.S and size[0] == 2, (high, size)
            M = size[1]
            place = self.reset_feed["place"]
            if self.place_cur + 2 * M > place.shape[0]:
                raise TapeExhausted("placement tape exhausted")
            out = place[self.place_cur:self.place_cur + 2 * M].reshape(2, M).copy()
            self.place_cur += 2 * M
            return out
        if not self.orient_served:
            # swarm_discrete_env.py:216  randint(8, size=num_agents)
            assert high == 8 and size == self.N, (high, size)
            self.orient_served = True
            return self.reset_feed["orient"].copy()
        # swarm_discrete_env.py:92 and :101  randint(S, size=2)
        assert high == self.S and size == 2, (high, size)
        pt = self.reset_feed["ptape"]
        if self.pred_cur >= pt.shape[0]:
            raise TapeExhausted("predator placement tape exhausted")
        out = pt[self.pred_cur].copy()
        self.pred_cur += 1
        return out

    def random(self):
        # swarm_discrete_env.py:138  roll = np.random.random()
        assert self.phase == "step" and not self.roll_served
        self.roll_served = True
        return self.step_feed["roll"]


class _NpProxy:
    """numpy, except ``random`` (tapes) and ``argsort`` (stable)."""

    def __init__(self, tape_random, stable_argsort=True):
        self.random = tape_random
        self._stable = stable_argsort

    def argsort(self, a, *args, **kwargs):
        if self._stable:
            kwargs["kind"] = "stable"
        return _np.argsort(a, *args, **kwargs)

    def __getattr__(self, name):
        return getattr(_np, name)


def _install_stub_modules():
    """Inject the minimal stand-ins the reference imports at module scope
    (swarm_discrete_env.py:1-15)."""
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")

        class Discrete(object):
            def __init__(self, n):
                self.n = n

        spaces.Discrete = Discrete
        error = types.ModuleType("gym.error")
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")
        gym.spaces, gym.error, gym.utils = spaces, error, utils
        utils.seeding = seeding
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.error": error,
                            "gym.utils": utils, "gym.utils.seeding": seeding})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        cbook = types.ModuleType("matplotlib.cbook")
        plt.imread = lambda *_a, **_k: _np.zeros((226, 226, 4), dtype=_np.float32)
        cbook.get_sample_data = lambda p, *_a, **_k: p
        mpl.pyplot, mpl.cbook = plt, cbook
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.cbook": cbook})
    import sklearn.metrics
    import sklearn.neighbors
    if not hasattr(sklearn.neighbors, "DistanceMetric"):
        sklearn.neighbors.DistanceMetric = sklearn.metrics.DistanceMetric


_REF_MODULE = None


def load_reference_module():
    """importlib-load the unmodified reference file (cached)."""
    global _REF_MODULE
    if _REF_MODULE is None:
        if not reference_available():
            raise FileNotFoundError(REFERENCE_FILE)
        _install_stub_modules()
        spec = importlib.util.spec_from_file_location("_gym_swarm_reference_sde", REFERENCE_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF_MODULE = mod
    return _REF_MODULE


UNIT_REWARD_TYPE = {"attraction": True, "repulsion": True, "alignment": True,
                    "indiv_rewards": False, "vf_size": None}


class TapedReferenceEnv:
    """One instance of the reference ``DiscreteSwarmEnv`` driven by tapes.

    NOTE: the proxy is installed on the (shared) reference module, so instances
    must be stepped one call at a time (each call re-installs its own proxy).
    """

    def __init__(self, N=4, S=20, att=5, rep=2, eat=-10, stable_argsort=True):
        self.mod = load_reference_module()
        self.env = self.mod.DiscreteSwarmEnv()
        self.env.set_env_parameters(num_agents=N, obs_space_size=S, verbose=False)
        self.env.set_reward_parameters(attraction_thresh=att, repulsion_thresh=rep,
                                       predator_eat_rew=eat, verbose=False)
        self.N, self.S = N, S
        self.tape = _TapeRandom()
        self.tape.S, self.tape.N = S, N
        self.proxy = _NpProxy(self.tape, stable_argsort)

    def _install(self):
        self.mod.np = self.proxy

    def reset(self, place, ptape, orient):
        """Reference reset (swarm_discrete_env.py:197-221). Returns a dict of facts."""
        self._install()
        self.tape.begin_reset(place, ptape, orient)
        obs = self.env.reset()
        return dict(
            x=_np.array([obs[i][0] for i in range(self.N)], dtype=_np.int64),
            y=_np.array([obs[i][1] for i in range(self.N)], dtype=_np.int64),
            pred=_np.array(self.env.predator.current_state, dtype=_np.int64),
            target=int(self.env.predator.current_target),
            place_used=self.tape.place_cur, pred_used=self.tape.pred_cur,
            orientation=_np.array([self.env.orientation[i] for i in range(self.N)]),
        )

    def step(self, actions, slab, roll, reward_type=None):
        """Reference step (swarm_discrete_env.py:223-267). ``roll`` is the value
        ``np.random.random()`` returns at :138."""
        self._install()
        self.tape.begin_step(slab, roll)
        action = {i: int(actions[i]) for i in range(self.N)}
        if reward_type is None:
            obs, reward, done, info = self.env.step(action)
        else:
            obs, reward, done, info = self.env.step(action, reward_type)
        # Integer counts: swarm_reward is a pure function of the committed state
        # (swarm_discrete_env.py:294-380); calling it again with unit weights
        # exposes the integer terms rew_rep / rew_attr exactly.
        unit, _ = self.env.swarm_reward(dict(UNIT_REWARD_TYPE, vf_size=(reward_type or UNIT_REWARD_TYPE)["vf_size"]))
        nn = self.N * (self.N - 1)
        keys = ("survival", "attraction", "repulsion", "alignment", "sum")
        out = dict(
            x=_np.array([obs[i][0] for i in range(self.N)], dtype=_np.int64),
            y=_np.array([obs[i][1] for i in range(self.N)], dtype=_np.int64),
            done=bool(done),
            pred_prev=_np.array(info["predator_state"], dtype=_np.int64),
            pred_next=_np.array(info["predator_next_state"], dtype=_np.int64),
            pred_orient=int(info["predator_orientation"]),
            target=int(self.env.predator.current_target),
            glob=_np.array([reward["global"][k] for k in keys], dtype=_np.float64),
            agent=_np.array([[reward[i][k] for k in keys] for i in range(self.N)], dtype=_np.float64),
            consumed=self.tape.slab_cur,
            eff_move=_np.array([self.env.move[i] for i in range(self.N)], dtype=_np.int64),
        )
        if done:
            out["eaten"] = int(_np.argmax(out["agent"][:, 0] != 0)) if _np.any(out["agent"][:, 0] != 0) else -1
            if out["eaten"] < 0:  # eat reward 0: recover from positions
                hit = (out["x"] == out["pred_next"][0]) & (out["y"] == out["pred_next"][1])
                out["eaten"] = int(_np.argmax(hit))
            out["cnt_rep"] = 0
            out["cnt_att"] = 0
            out["agent_cnt"] = _np.zeros((self.N, 2), dtype=_np.int64)
        else:
            out["eaten"] = -1
            out["cnt_rep"] = int(round(unit["global"]["repulsion"] * nn))
            out["cnt_att"] = int(round(unit["global"]["attraction"] * nn))
            out["agent_cnt"] = _np.array(
                [[int(round(unit[i]["repulsion"] * nn)), int(round(unit[i]["attraction"] * nn))]
                 for i in range(self.N)], dtype=_np.int64)
        return out
