"""CPU oracle restatement of the two POMDP environments + TimeLimit (test infrastructure only).

Follows:
  envs/car_flag.py:18-133,145-159   CarFlag dynamics / reset (f64, exact operation order, no FMA)
  envs/memory_cards.py:43-116       Memory card game
  envs/__init__.py:31-48            max_episode_steps: Memory-5-v0 = 50, DiscreteCarFlag-v0 = 200
  gym 0.18 TimeLimit (not under /root/reference; pinned gym==0.18.0 in requirements.txt:11):
      elapsed += 1; if elapsed >= max: info["TimeLimit.truncated"] = not done; done = True
Each env owns a ``oracle.pcg64.PCG64`` seeded through numpy's SeedSequence exactly like ``env.seed(seed)``
(car_flag.py:70-74, memory_cards.py:64-68), so its RNG state can be compared with the device's.
"""
import numpy as np

from oracle.pcg64 import PCG64

CARFLAG_MAX_STEPS = 200
MEMORY_MAX_STEPS = 50


class CarFlagOracle:
    """envs/car_flag.py CarFlag(discrete=True).  State (p, v, dir) is float64."""

    obs_dim = 3
    num_actions = 3
    max_episode_steps = CARFLAG_MAX_STEPS
    obs_mask = -5  # utils/env_processing.py:113-118 (Box space)

    def __init__(self, seed: int):
        self.rng = PCG64.from_seed(seed)
        self.heaven = 1.0
        self.state = np.zeros(3)

    def reset(self) -> np.ndarray:
        # car_flag.py:147-152: heaven side; :158 start position
        self.heaven = 1.0 if self.rng.integers(0, 2) == 0 else -1.0
        p0 = self.rng.uniform(-0.2, 0.2)
        self.state = np.array([p0, 0.0, 0.0])
        return self.state.copy()

    def step(self, action: int):
        p, v = float(self.state[0]), float(self.state[1])
        force = action - 1                      # :81
        v = v + force * 0.0015                  # :85 (product rounded, then sum rounded)
        if v > 0.07:                            # :86-89
            v = 0.07
        if v < -0.07:
            v = -0.07
        p = p + v                               # :90
        if p > 1.1:                             # :91-94
            p = 1.1
        if p < -1.1:
            p = -1.1
        if p == -1.1 and v < 0:                 # :95-96
            v = 0.0
        hell = -self.heaven
        hi, lo = max(self.heaven, hell), min(self.heaven, hell)
        done = bool(p >= hi or p <= lo)         # :98-101
        r = 0.0
        if self.heaven > hell:                  # :105-110
            if p >= self.heaven:
                r = 1.0
            if p <= hell:
                r = -1.0
        if self.heaven < hell:                  # :112-117
            if p <= self.heaven:
                r = 1.0
            if p >= hell:
                r = -1.0
        d = 0.0
        if p >= 0.5 - 0.2 and p <= 0.5 + 0.2:   # :119-129
            d = 1.0 if self.heaven > hell else -1.0
        self.state = np.array([p, v, d])
        return self.state.copy(), r, done, {"is_success": r > 0}


class MemoryOracle:
    """envs/memory_cards.py Memory(num_pairs=5).  Observation is float64 holding small integers."""

    obs_dim = 10
    num_actions = 10
    max_episode_steps = MEMORY_MAX_STEPS
    obs_mask = 8  # max(nvec)+1, utils/env_processing.py:111-112

    def __init__(self, seed: int, num_pairs: int = 5):
        self.rng = PCG64.from_seed(seed)
        self.num_pairs = num_pairs
        self.num_cards = 2 * num_pairs
        self.removed = num_pairs + 1
        self.cards = [self.removed] * self.num_cards
        self.obs = np.zeros(self.num_cards)
        self.cur = -1

    def reset(self) -> np.ndarray:
        self.cards = [c for c in range(1, self.num_pairs + 1) for _ in range(2)]  # :72 np.repeat
        self.rng.shuffle(self.cards)                                              # :73
        self.obs = np.zeros(self.num_cards)                                       # :75
        self.cur = self.rng.integers(self.num_cards)                              # :77
        self.obs[self.cur] = self.cards[self.cur]                                 # :78
        return self.obs.copy()

    def step(self, action: int):
        if all(c == self.removed for c in self.cards):                            # :83-84
            raise ValueError("Trying to take step in invalid state. Did you reset?")
        done, info = False, {}
        if action == self.cur:                                                    # :89-91
            self.obs[self.cur] = 0
            r = -1
        elif self.cards[action] == self.obs[self.cur]:                            # :93-103
            self.obs[action] = self.removed
            self.obs[self.cur] = self.removed
            r = 0
            if np.all(self.obs == self.removed):
                done = True
                info["is_success"] = True
        else:                                                                     # :104-106
            self.obs[self.cur] = 0
            r = -1
        if not done:                                                              # :108-114
            self.cur = self.rng.integers(self.num_cards)
            while self.obs[self.cur] == self.removed:
                self.cur = self.rng.integers(self.num_cards)
            self.obs[self.cur] = self.cards[self.cur]
        return self.obs.copy(), r, done, info


class TimeLimitOracle:
    """gym 0.18 TimeLimit around an oracle env; same (obs, r, done, info) contract."""

    def __init__(self, env, max_episode_steps=None):
        self.env = env
        self.max_episode_steps = max_episode_steps or env.max_episode_steps
        self.elapsed = None

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self.elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, r, done, info = self.env.step(action)
        self.elapsed += 1
        if self.elapsed >= self.max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, r, done, info


ENVS = {"DiscreteCarFlag-v0": CarFlagOracle, "Memory-5-v0": MemoryOracle}


def make(env_id: str, seed: int) -> TimeLimitOracle:
    """gym.make(id) + env.seed(seed) (envs/__init__.py:31-48, utils/random.py:26-29)."""
    return TimeLimitOracle(ENVS[env_id](seed))


def rollout(env_id: str, seed: int, actions, auto_reset=True):
    """Roll one env over an action tape with auto-reset on done (run.py:290-296 loop shape).

    Returns per-step arrays: obs_after_step [T,O] (f64), reward, done, truncated, and the obs produced by the
    auto-reset (nan rows where no reset happened), plus the initial obs and the final RNG state tuple.
    """
    env = make(env_id, seed)
    o0 = env.reset()
    T, O = len(actions), env.obs_dim
    obs = np.zeros((T, O)); rew = np.zeros(T); done = np.zeros(T, bool); trunc = np.zeros(T, bool)
    reset_obs = np.full((T, O), np.nan)
    for t, a in enumerate(actions):
        o, r, d, info = env.step(int(a))
        obs[t], rew[t], done[t] = o, r, d
        trunc[t] = bool(info.get("TimeLimit.truncated", False))
        if d and auto_reset:
            reset_obs[t] = env.reset()
    return dict(initial_obs=o0, obs=obs, reward=rew, done=done, truncated=trunc, reset_obs=reset_obs,
                rng_state=np.array(env.rng.as_tuple(), dtype=np.uint64))
