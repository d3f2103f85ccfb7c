"""Compact replay ring: the device-side stand-in for the reference's flashbax prioritised flat buffer
(main.py:217-225 `make_prioritised_flat_buffer(..., add_sequences=True, add_batch_size=selfplay_batch_size)`,
main.py:383-385 `buffer_fn.add`, main.py:413 `buffer_fn.sample`).

The reference stores a full `pgx.State` per env per step -- for DeepSea-30 that is 900 observation bytes + bookkeeping,
for DeepSea-100 10 KB.  Here a step is the COMPACT state (`eaz_env_compact`: 4 bytes DeepSea, 40 + ws bytes Subleq) plus
the rewards leaf; `sample()` draws (t, env) pairs and decodes `ExperiencePair(first=state_t, second=state_t+1)` on the
device with `eaz_env_uncompact` (observations are materialised only on request: the fused search and the network kernels
read compact states directly).  Sampling is uniform over the stored consecutive pairs -- the reference never updates
priorities, so its prioritised buffer samples uniformly too.  Multi-GPU: every rank owns the ring of its env shard."""
from __future__ import annotations

from . import ops
from ._lib import EazError, require_cuda
from .reanalyze import ExperiencePair


class CompactReplayRing:
    def __init__(self, env_spec: ops.EnvSpec, max_length_time_axis: int, add_batch_size: int, device="cuda", seed: int = 0):
        torch = require_cuda()
        self.env, self.T, self.B, self.device = env_spec, int(max_length_time_axis), int(add_batch_size), device
        if self.T < 2:
            raise EazError("the ring needs at least two time steps to form (state, next state) pairs")
        self.S = env_spec.compact_bytes
        self.states = torch.zeros((self.T, self.B, self.S), dtype=torch.uint8, device=device)
        self.rewards = torch.zeros((self.T, self.B), dtype=torch.float32, device=device)
        self.cursor = 0   # next time slot to write
        self.length = 0   # number of valid time slots
        self.gen = torch.Generator(device=device).manual_seed(seed)

    @property
    def bytes_per_step(self) -> int:
        return self.B * (self.S + 4)

    def add(self, state: dict) -> None:
        """Append one time step for all `add_batch_size` envs (call once per self-play step, in order)."""
        self.states[self.cursor].copy_(ops.env_compact(self.env, state))
        self.rewards[self.cursor].copy_(state["rewards"].reshape(self.B))
        self.cursor = (self.cursor + 1) % self.T
        self.length = min(self.length + 1, self.T)

    def can_sample(self, min_length: int = 2) -> bool:
        return self.length >= max(2, min_length)

    def sample(self, batch_size: int, with_obs: bool = False) -> ExperiencePair:
        """Uniformly sampled consecutive pairs (state_t, state_t+1) of the same env; t+1 never crosses the write cursor."""
        torch = require_cuda()
        if not self.can_sample():
            raise EazError("not enough steps in the ring to sample a pair")
        oldest = (self.cursor - self.length) % self.T
        k = torch.randint(0, self.length - 1, (batch_size,), device=self.device, generator=self.gen)  # pair index in logical time
        b = torch.randint(0, self.B, (batch_size,), device=self.device, generator=self.gen)
        t0, t1 = (oldest + k) % self.T, (oldest + k + 1) % self.T
        first = ops.env_uncompact(self.env, self.states[t0, b], self.rewards[t0, b], with_obs=with_obs)
        second = ops.env_uncompact(self.env, self.states[t1, b], self.rewards[t1, b], with_obs=with_obs)
        return ExperiencePair(first=first, second=second)
