"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product path.

Times the UNMODIFIED reference ``DiscreteSwarmEnv`` (``gym_swarm/envs/swarm_discrete_env.py``, loaded by path through
``oracle/reference_shim.py``'s import stubs) on the host cores: SURVEY.md 8d "CPU baseline" -- raw NumPy, the real
process-global ``np.random``, ``multiprocessing`` with one worker per core, each worker looping ``env.step`` with random
actions (``reset`` on done, as a user's loop would: ``swarm_discrete_env.py:223-267`` / ``:197-221``) over its share of
a SAMPLE of environments.  Used by ``bench.py`` (``cpu_baseline`` and ``--impl reference``) only.
"""
from __future__ import annotations

import os
import time


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _worker(args):
    (wid, n_envs, N, S, att, rep, eat, warmup, steps, reward_type, seed) = args
    os.environ["OMP_NUM_THREADS"] = "1"             # one core per worker: sklearn / BLAS must not oversubscribe
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import reference_shim as shim
    mod = shim.load_reference_module()              # the unmodified file; module-global np untouched (raw NumPy)
    np.random.seed(seed + wid)                      # the reference draws from the process-global generator
    envs = []
    for _ in range(n_envs):
        env = mod.DiscreteSwarmEnv()
        env.set_env_parameters(num_agents=N, obs_space_size=S, verbose=False)
        env.set_reward_parameters(attraction_thresh=att, repulsion_thresh=rep, predator_eat_rew=eat, verbose=False)
        env.reset()
        envs.append(env)
    episodes = 0

    def one_pass():
        nonlocal episodes
        for env in envs:
            acts = np.random.randint(8, size=N)
            action = {i: int(acts[i]) for i in range(N)}
            _, _, done, _ = env.step(action, reward_type) if reward_type is not None else env.step(action)
            if done:
                episodes += 1
                env.reset()

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    return dict(worker=wid, envs=n_envs, seconds=dt, env_steps=n_envs * steps, episodes=episodes)


def time_reference(N, S, sample_envs, steps, warmup=1, att=5, rep=2, eat=-10, reward_type=None, workers=None, seed=1234,
                   timeout_s=600):
    """agent-steps/s of the reference on ``workers`` processes (default: every core this process may use), each stepping
    its share of ``sample_envs`` environments ``steps`` times.  Workers are plain subprocesses of this module (no fork
    of a parent that may hold a CUDA context, no re-import of the caller's main module).  Returns a dict for the bench
    record."""
    import json
    import subprocess
    import sys
    from oracle import reference_shim as shim
    if not shim.reference_available():
        raise FileNotFoundError(shim.REFERENCE_FILE)
    workers = host_cores() if workers is None else int(workers)
    workers = max(1, min(workers, sample_envs))
    base, extra = divmod(int(sample_envs), workers)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", PYTHONPATH=root)
    t0 = time.perf_counter()
    procs = []
    for w in range(workers):
        job = [w, base + (1 if w < extra else 0), N, S, att, rep, eat, warmup, steps, reward_type, seed]
        procs.append(subprocess.Popen([sys.executable, "-m", "oracle.reference_runner", json.dumps(job)], cwd=root, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = []
    try:
        for pr in procs:
            out, err = pr.communicate(timeout=max(1.0, timeout_s - (time.perf_counter() - t0)))
            if pr.returncode != 0:
                raise RuntimeError("reference worker failed: " + err.strip()[-400:])
            res.append(json.loads(out.strip().splitlines()[-1]))
    finally:
        for pr in procs:                             # exact PIDs we started, nothing by pattern
            if pr.poll() is None:
                pr.kill()
    wall = time.perf_counter() - t0
    slowest = max(r["seconds"] for r in res)
    env_steps = sum(r["env_steps"] for r in res)
    per_core = [r["env_steps"] * N / r["seconds"] for r in res]
    return dict(value=env_steps * N / slowest, seconds_per_step=slowest / steps, workers=workers, sample_envs=int(sample_envs),
                steps=int(steps), per_core_mean=sum(per_core) / len(per_core), per_core_min=min(per_core),
                episodes=sum(r["episodes"] for r in res), wall_with_startup=wall, reference_file=shim.REFERENCE_FILE)


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(_worker(tuple(json.loads(sys.argv[1])))), flush=True)
