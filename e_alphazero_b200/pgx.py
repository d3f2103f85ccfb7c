"""pgx-compatible surface for the two custom envs of the reference (envs/deep_sea.py, envs/subleq.py).

`Env.init / step / observe` and the `State` leaves keep the names the callers use
(selfplay.py:135, context.py:127, evaluate.py:47,54, main.py:354-369), but are natively
batched: every leaf carries a leading batch axis (what `jax.vmap(env.step)` produces in
the reference), and all transitions run in CUDA through the C ABI.
"""
from __future__ import annotations

from enum import IntEnum

from . import _abi, ops
from ._lib import EazError, require_cuda


class SubleqTask(IntEnum):  # envs/subleq.py:110-124
    NEGATION_POSITIVE = 1
    NEGATION = 2
    IDENTITY = 3
    SUBTRACTION = 4
    ADDITION = 5
    MAXIMUM = 6
    MINIMUM = 7
    COMPARISON = 8
    SORT_2 = 9
    SORT_3 = 10
    SORT_4 = 11
    MULTIPLICATION = 12
    DIVISION = 13
    SUMMATION = 14


solved_or_not = _abi.SUBLEQ_REWARD_SOLVED  # subleq.py:535-537
lowest_bytes = _abi.SUBLEQ_REWARD_LOWEST_BYTES  # subleq.py:540-542

_LEAF_ALIASES = {
    "_step_count": "step_count",
    "_horizontal_position": "col",
    "_memory_state": "memory",
    "_task": "task",
    "_solved": "solved",
    "_example_input_after": "input_after",
    "_example_output_after": "output_after",
}


class State:
    """Batched pgx.State (DeepSeaState deep_sea.py:12-27 / SubleqState subleq.py:545-568)."""

    def __init__(self, env: "Env", leaves: dict):
        self._env = env
        self.leaves = leaves

    def __getattr__(self, name):
        leaves = self.__dict__["leaves"]
        key = _LEAF_ALIASES.get(name, name)
        if key in leaves:
            v = leaves[key]
            return v.bool() if key in ("terminated", "truncated", "solved") else v
        env = self.__dict__["_env"]
        torch = require_cuda()
        B = leaves["step_count"].shape[0]
        if name == "observation":  # pure function of the other leaves
            return ops.env_observe(env.spec, leaves).bool().reshape((B,) + env.spec.obs_shape)
        if name == "legal_action_mask":  # all True for both envs (deep_sea.py:19, subleq.py:626)
            return torch.ones((B, env.num_actions), dtype=torch.bool, device=leaves["step_count"].device)
        if name == "current_player":
            return torch.zeros(B, dtype=torch.int32, device=leaves["step_count"].device)
        if name in ("_example_input", "_example_output", "_test_cases"):
            return env._test_case_leaf(name, leaves)
        raise AttributeError(name)

    @property
    def env_id(self):
        return self._env.id

    @property
    def batch_size(self):
        return self.leaves["step_count"].shape[0]

    def replace(self, **kw):
        leaves = dict(self.leaves)
        for k, v in kw.items():
            leaves[_LEAF_ALIASES.get(k, k)] = v
        return State(self._env, leaves)

    def __getitem__(self, idx):
        return State(self._env, {k: v[idx] for k, v in self.leaves.items()})


class Env:
    spec: ops.EnvSpec

    def init(self, key=None, batch_size: int | None = None) -> State:
        """vmap(env.init)(keys).  `key`: int seed / torch.Generator / explicit int32 task ids [B] for Subleq."""
        raise NotImplementedError

    def step(self, state: State, action, key=None) -> State:
        """vmap(env.step): pgx core semantics (absorbing terminal states, _step_count incremented before _step)."""
        return State(self, ops.env_step(self.spec, state.leaves, action))

    def step_auto_reset(self, state: State, action, key=None) -> State:
        """vmap(auto_reset(env.step, env.init)) with the reference's wrapper (selfplay.py:26-75)."""
        tasks = self._draw_tasks(key, state.batch_size) if isinstance(self, Subleq) else None
        return State(self, ops.env_step(self.spec, state.leaves, action, auto_reset=True, task_ids=tasks))

    def observe(self, state: State, player_id=None):
        return state.observation

    @property
    def num_actions(self) -> int:
        return self.spec.num_actions

    @property
    def num_players(self) -> int:
        return 1

    @property
    def version(self) -> str:
        return "0.0.1"


class DeepSea(Env):
    """envs/deep_sea.py:30-102.  `action_map`: bool [N,N] (the reference draws it once from a key, :51-52)."""

    id = "deep_sea"

    def __init__(self, size_of_grid: int = 4, action_map=None, device="cuda"):
        self.size_of_grid = size_of_grid
        self.spec = ops.deepsea_spec(size_of_grid, action_map, device)
        self.action_map = self.spec.action_map
        self.device = device

    def init(self, key=None, batch_size: int | None = None) -> State:
        if batch_size is None:
            raise EazError("DeepSea.init needs batch_size (the reference vmaps init over a batch of keys)")
        return State(self, ops.env_init(self.spec, batch_size, device=self.device))


class Subleq(Env):
    """envs/subleq.py:571-724."""

    id = "subleq"

    def __init__(self, tasks, word_size: int = 256, reward_fn=solved_or_not, use_binary_encoding: bool = False, device="cuda"):
        if not 16 <= word_size <= 256:  # subleq.py:606
            raise EazError("assert 16 <= word_size <= 256 (subleq.py:606)")
        self.tasks = [int(t) for t in tasks]
        self.word_size = word_size
        self.spec = ops.subleq_spec(word_size, use_binary_encoding, int(reward_fn))
        self.device = device

    def _draw_tasks(self, key, B):
        """jax.random.choice(key, self.tasks) per env (subleq.py:624), drawn here with torch."""
        torch = require_cuda()
        if key is not None and hasattr(key, "shape") and tuple(key.shape) == (B,):
            return key  # explicit pre-drawn task ids
        gen = key if isinstance(key, torch.Generator) else None
        if isinstance(key, int):
            gen = torch.Generator(device=self.device).manual_seed(key)
        tasks = torch.tensor(self.tasks, dtype=torch.int32, device=self.device)
        idx = torch.randint(0, len(self.tasks), (B,), device=self.device, generator=gen)
        return tasks[idx].contiguous()

    def init(self, key=None, batch_size: int | None = None) -> State:
        if batch_size is None:
            if key is not None and hasattr(key, "shape"):
                batch_size = key.shape[0]
            else:
                raise EazError("Subleq.init needs batch_size or explicit task ids")
        return State(self, ops.env_init(self.spec, batch_size, self._draw_tasks(key, batch_size), device=self.device))

    def _test_case_leaf(self, name, leaves):
        torch = require_cuda()
        import numpy as np

        tabs = {t: ops.subleq_test_cases(t, self.word_size) for t in set(leaves["task"].tolist())}
        tin = np.stack([tabs[t][0] for t in leaves["task"].tolist()])
        tout = np.stack([tabs[t][1] for t in leaves["task"].tolist()])
        dev = leaves["task"].device
        if name == "_example_input":
            return torch.as_tensor(tin[:, 0]).to(dev)
        if name == "_example_output":
            return torch.as_tensor(tout[:, 0]).to(dev)
        return torch.as_tensor(tin).to(dev), torch.as_tensor(tout).to(dev)
