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


class _ParamEntry:
    """Device copy of one haiku (params, state) pair.  Holds STRONG references to the two pytrees, so their id()s cannot be
    reused by other objects while the entry is alive (a key on bare id()s returned stale weights once the learner freed the old
    pytrees: CPython hands the addresses out again)."""

    def __init__(self, model, fc, sources):
        self.model = (model[0], model[1])
        self.fc = fc
        self.sources = sources  # [(device tensor, source leaf, version token or None)]


_param_cache: dict = {}


def _leaf_token(leaf):
    """Change token of a source leaf: torch tensors count their in-place updates (`_version`); numpy / other leaves have no such
    counter and are re-copied on every call."""
    import torch

    if torch.is_tensor(leaf):
        return (leaf.data_ptr(), leaf._version)
    return None


def _haiku_sources(params, state, prefix="fc_az_net"):
    names = [f"{prefix}/linear" + ("" if i == 0 else f"_{i}") for i in range(12)]
    w = [[params[names[h * 3 + l]]["w"] for l in range(3)] for h in range(4)]
    b = [[params[names[h * 3 + l]]["b"] for l in range(3)] for h in range(4)]
    return w, b, state[f"{prefix}/xxhash32"]["binary_set"]


def as_fc_params(model, rf: FusedRecurrentFn | None = None, env: Env | None = None) -> ops.FcParams:
    """Accepts an ops.FcParams, or the reference's `model = (params, state)` haiku pytrees (numpy / torch leaves).

    The device copy is cached per pytree PAIR (strong references, identity-checked) and REFRESHED on every call: a leaf whose
    change token moved -- or that has none (numpy) -- is copied again into the existing device tensor, and `FcParams.version`
    is bumped so that plans rebuild their parameter-derived tables.  New pytrees (a learner update) get a new entry."""
    import torch

    if isinstance(model, ops.FcParams):
        return model
    env = env or (rf.env if rf is not None else None)
    if not (isinstance(model, (tuple, list)) and len(model) == 2 and env is not None):
        raise EazError("params must be an ops.FcParams or a (haiku params, haiku state) pair")
    key = (id(model[0]), id(model[1]), id(env))
    hit = _param_cache.get(key)
    if hit is not None and hit.model[0] is model[0] and hit.model[1] is model[1]:
        w, b, bset = _haiku_sources(model[0], model[1])
        fresh = [w[h][l] for h in range(4) for l in range(3)] + [b[h][l] for h in range(4) for l in range(3)] + [bset]
        changed = False
        for i, (dst, src, tok) in enumerate(hit.sources):
            cur = fresh[i]
            ntok = _leaf_token(cur)
            if cur is src and tok is not None and ntok == tok:
                continue
            dst.copy_(torch.as_tensor(cur).to(dtype=dst.dtype).reshape(dst.shape), non_blocking=False)
            hit.sources[i] = (dst, cur, ntok)
            changed = True
        if changed:
            hit.fc.version += 1
        return hit.fc
    while len(_param_cache) >= 8:  # bounded: drop the oldest entry (dicts keep insertion order)
        _param_cache.pop(next(iter(_param_cache)))
    subleq = env.spec.kind == _abi.ENV_SUBLEQ
    fc = ops.FcParams.from_haiku(model[0], model[1], env.num_actions, hash_io=int(subleq), word_size=env.spec.word_size if subleq else 0)
    w, b, bset = _haiku_sources(model[0], model[1])
    srcs = [w[h][l] for h in range(4) for l in range(3)] + [b[h][l] for h in range(4) for l in range(3)] + [bset]
    dsts = [fc.w[h][l] for h in range(4) for l in range(3)] + [fc.b[h][l] for h in range(4) for l in range(3)] + [fc.binary_set]
    _param_cache[key] = _ParamEntry(model, fc, [(d, s_, _leaf_token(s_)) for d, s_ in zip(dsts, srcs)])
    return fc


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


@dataclass(eq=False)
class BoardEnvSpec:
    """What the convolutional evaluators need to know about a pgx env that is not DeepSea / Subleq (context.py:76-82 dispatches on
    config.env_id): its id, the observation shape [H, W, C] and the number of actions."""

    env_id: str
    observation_shape: tuple
    num_actions: int


class ConvForwardFn:
    """forward.apply for EpistemicResidualAZNet (network/resnet.py) / EpistemicMinatarAZNet (network/minatar.py): the same
    5-tuple as ForwardFn from `eaz_convnet_forward`.  The device copy of the haiku pytrees is cached per (params, state) PAIR with
    strong references (identity-checked, like as_fc_params); call `params_updated()` after in-place changes of numpy leaves."""

    def __init__(self, env: BoardEnvSpec, config=None, mlp_mode: int = _abi.MLP_EXACT):
        self.env, self.mlp_mode = env, int(mlp_mode)
        self.minatar = "minatar" in env.env_id
        g = lambda k, d: getattr(config, k, d) if config is not None else d
        self.kw = dict(hash_bits=24, max_u=1.0, novelty_scale=1.0)
        if self.minatar:  # context.py:52-60
            self.kw.update(num_channels=g("num_channels", 16), hidden=g("linear_layer_size", 64), max_u=g("max_ube", 1.0),
                           novelty_scale=g("max_epistemic_variance_reward", 1.0), discount=g("discount", 0.9997))
        self._cache = None  # (params, state, ConvNetParams)

    def params_updated(self):
        self._cache = None

    def _net(self, params, state):
        c = self._cache
        if c is not None and c[0] is params and c[1] is state:
            return c[2]
        H, W, Cc = self.env.observation_shape
        kind = _abi.CONVNET_MINATAR if self.minatar else _abi.CONVNET_RESNET
        desc = _abi.convnet_description(params, state, kind, H, W, Cc, self.env.num_actions, **self.kw)
        net = ops.ConvNetParams(dict(desc, mlp_mode=self.mlp_mode))
        self._cache = (params, state, net)
        return net

    def apply(self, params, state, observation, is_training: bool = False, update_hash: bool = False):
        if is_training or update_hash:
            raise NotImplementedError("training-mode forward / hash update are learner-side (train.py) and out of scope")
        ev = self._net(params, state).forward(observation)
        return (ev["exploit_logits"], ev["explore_logits"], ev["value"], ev["ube"], ev["novelty"]), state


def get_forward_fn(env, config=None, mlp_mode: int = _abi.MLP_EXACT):
    """context.get_forward_fn (context.py:84-106) with the network dispatch of context.get_network (:40-82): the FC network for
    DeepSea / Subleq, EpistemicMinatarAZNet for MinAtar env ids, EpistemicResidualAZNet for every other pgx env."""
    if isinstance(env, Env):
        return ForwardFn(env)
    if isinstance(env, BoardEnvSpec):
        return ConvForwardFn(env, config, mlp_mode)
    raise NotImplementedError("get_forward_fn: pass an e_alphazero_b200.pgx env (DeepSea / Subleq) or a context.BoardEnvSpec")
