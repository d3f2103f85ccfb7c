"""Mirror of the reference's context.py:84-157 for the fused path: `get_forward_fn` (the FC network's
forward.apply) and `get_epistemic_recurrent_fn` (env.step + network + glue, executed inside the search
kernels -- the returned object is a descriptor the search recognises, not a Python callback)."""
from __future__ import annotations

from dataclasses import dataclass

from . import _abi, ops
from ._lib import EazError, require_cuda
from .pgx import Env, State


@dataclass(eq=False)
class FusedRecurrentFn:
    """What context.get_epistemic_recurrent_fn(env, forward, batch_size, exploration, discount, two_players_game)
    closes over (context.py:109-116)."""

    env: Env
    batch_size: int
    exploration: bool
    discount: float
    two_players_game: bool
    mlp_mode: int = _abi.MLP_EXACT

    def __call__(self, model, rng_key, action, state: State):
        """Stand-alone evaluation of the recurrent_fn (one env.step + network + glue), for callers that want it
        outside a search; same kernels as the fused path."""
        from .emctx import EpistemicRecurrentFnOutput

        torch = require_cuda()
        net = as_fc_params(model, self)
        nxt = self.env.step(state, action)
        ev = ops.mlp_forward_states(net, self.env.spec, nxt.leaves)
        logits = ev["explore_logits"] if self.exploration else ev["exploit_logits"]
        logits = logits - logits.max(dim=-1, keepdim=True).values  # context.py:135
        term = nxt.terminated
        zero = torch.zeros_like(ev["value"])
        disc = torch.full_like(ev["value"], -self.discount if self.two_players_game else self.discount)
        out = EpistemicRecurrentFnOutput(reward=nxt.rewards[:, 0], reward_epistemic_variance=zero, discount=torch.where(term, zero, disc),
                                         prior_logits=logits, value=torch.where(term, zero, ev["value"]),
                                         value_epistemic_variance=torch.where(term, zero, ev["ube"]))
        return out, nxt


def get_epistemic_recurrent_fn(env: Env, forward, batch_size: int, exploration: bool, discount: float, two_players_game: bool,
                               mlp_mode: int = _abi.MLP_EXACT) -> FusedRecurrentFn:
    if not isinstance(env, Env):
        raise NotImplementedError("the fused recurrent_fn exists for e_alphazero_b200.pgx.DeepSea / Subleq only")
    return FusedRecurrentFn(env, int(batch_size), bool(exploration), float(discount), bool(two_players_game), int(mlp_mode))


_param_cache: dict = {}


def as_fc_params(model, rf: FusedRecurrentFn | None = None, env: Env | None = None) -> ops.FcParams:
    """Accepts an ops.FcParams, or the reference's `model = (params, state)` haiku pytrees (numpy / torch leaves)."""
    if isinstance(model, ops.FcParams):
        return model
    env = env or (rf.env if rf is not None else None)
    if isinstance(model, (tuple, list)) and len(model) == 2 and env is not None:
        key = (id(model[0]), id(model[1]))
        hit = _param_cache.get(key)
        if hit is None:
            if len(_param_cache) > 8:
                _param_cache.clear()
            subleq = env.spec.kind == _abi.ENV_SUBLEQ
            hit = _param_cache[key] = ops.FcParams.from_haiku(model[0], model[1], env.num_actions, hash_io=int(subleq),
                                                             word_size=env.spec.word_size if subleq else 0)
        return hit
    raise EazError("params must be an ops.FcParams or a (haiku params, haiku state) pair")


class ForwardFn:
    """forward.apply(params, state, observation, is_training=False) -> ((exploit, explore, value, ube, reward_var), state)
    (context.py:84-106; output order fully_connected.py:101).  `observation` may be a bool observation batch or a
    pgx State (then the observation is never materialised)."""

    def __init__(self, env: Env):
        self.env = env

    def apply(self, params, state, observation, is_training: bool = False, update_hash: bool = False):
        if is_training or update_hash:
            raise NotImplementedError("training-mode forward / hash update are learner-side (train.py) and out of scope")
        net = as_fc_params(params if state is None else (params, state), env=self.env)
        if isinstance(observation, State):
            ev = ops.mlp_forward_states(net, self.env.spec, observation.leaves)
        else:
            ev = ops.mlp_forward(net, observation)
        return (ev["exploit_logits"], ev["explore_logits"], ev["value"], ev["ube"], ev["novelty"]), state


def get_forward_fn(env: Env, config=None) -> ForwardFn:
    return ForwardFn(env)
