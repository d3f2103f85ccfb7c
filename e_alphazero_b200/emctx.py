"""emctx-compatible surface of the search (the names selfplay.py:100-142, reanalyze.py:70-116 and
evaluate.py:29-47 use), backed by the fused CUDA search in libeaz_b200.so.

    root = emctx.EpistemicRootFnOutput(prior_logits, value, value_epistemic_variance, embedding, beta)
    out  = emctx.epistemic_gumbel_muzero_policy(params=model, rng_key=key, root=root, recurrent_fn=rf,
               num_simulations=n, invalid_actions=~mask,
               qtransform=functools.partial(emctx.epistemic_qtransform_completed_by_mix_value, rescale_values=True))
    out.action, out.action_weights, out.search_tree.epistemic_summary()

`recurrent_fn` must come from `e_alphazero_b200.context.get_epistemic_recurrent_fn` (DeepSea / Subleq
+ FC net): env step, network and the glue of context.py:117-155 are fused into the expand step.
Any other callable raises NotImplementedError -- there is no generic (CPU) fallback.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Any

from . import _abi, ops
from ._lib import EazError, require_cuda


@dataclass
class EpistemicRootFnOutput:
    prior_logits: Any
    value: Any
    value_epistemic_variance: Any
    embedding: Any
    beta: Any


@dataclass
class EpistemicRecurrentFnOutput:
    reward: Any
    reward_epistemic_variance: Any
    discount: Any
    prior_logits: Any
    value: Any
    value_epistemic_variance: Any


@dataclass
class EpistemicSearchSummary:
    visit_counts: Any
    visit_probs: Any
    value: Any
    value_epistemic_std: Any
    qvalues: Any
    qvalues_epistemic_variance: Any


class EpistemicTree:
    """Search result.  Summary arrays are always present; the full struct-of-arrays tree in emctx layout
    ([B,N], [B,N,A]) only if the policy was called with `return_tree=True`."""

    ROOT_INDEX = 0
    NO_PARENT = -1
    UNVISITED = -1

    def __init__(self, arrays: dict, num_actions: int, num_simulations: int):
        self._a = arrays
        self.num_actions = num_actions
        self.num_simulations = num_simulations

    def __getattr__(self, name):
        a = self.__dict__["_a"]
        if name in a and name not in {f[0] for f in _abi.SUMMARY_FIELDS}:
            return a[name]
        raise AttributeError(f"{name} (full tree arrays need return_tree=True)")

    def epistemic_summary(self) -> EpistemicSearchSummary:
        a = self._a
        return EpistemicSearchSummary(a["visit_counts"], a["visit_probs"], a["value"], a["value_epistemic_std"], a["qvalues"],
                                      a["qvalues_epistemic_variance"])

    summary = epistemic_summary


@dataclass
class PolicyOutput:
    action: Any
    action_weights: Any
    search_tree: EpistemicTree


def epistemic_qtransform_completed_by_mix_value(tree=None, node_index=None, *, value_scale=0.1, maxvisit_init=50.0, rescale_values=True,
                                                use_mixed_value=True, epsilon=1e-8):
    """Marker for the qtransform (its arithmetic is inside the search kernels).
    Pass it -- or a functools.partial of it -- as `qtransform=`."""
    raise EazError("the qtransform is evaluated inside the fused search; pass this function (or a partial of it) as qtransform=")


class PreDrawnGumbel:
    """rng_key stand-in carrying pre-drawn standard Gumbel noise [B,A] (parity tests)."""

    def __init__(self, gumbel):
        self.gumbel = gumbel


def _qtransform_kwargs(qtransform) -> dict:
    kw = {}
    fn = qtransform
    while isinstance(fn, functools.partial):
        kw = {**fn.keywords, **kw}
        fn = fn.func
    if fn is not epistemic_qtransform_completed_by_mix_value:
        raise NotImplementedError("only emctx.epistemic_qtransform_completed_by_mix_value is supported by the fused search")
    allowed = {"value_scale", "maxvisit_init", "rescale_values", "use_mixed_value", "epsilon"}
    if set(kw) - allowed:
        raise TypeError(f"unknown qtransform arguments {set(kw) - allowed}")
    return kw


def _draw_gumbel(rng_key, shape, device):
    torch = require_cuda()
    if isinstance(rng_key, PreDrawnGumbel):
        return rng_key.gumbel.to(device=device, dtype=torch.float32).contiguous()
    gen = rng_key if isinstance(rng_key, torch.Generator) else None
    if isinstance(rng_key, int):
        gen = torch.Generator(device=device).manual_seed(rng_key)
    u = torch.rand(shape, device=device, generator=gen).clamp_(1e-20, 1.0 - 1e-7)
    return (-(-u.log()).log()).contiguous()


_plans: dict = {}


def qtransform_by_parent_and_siblings(tree=None, node_index=None, *, epsilon=1e-8):
    """Marker for mctx's PUCT qtransform (evaluated inside the search kernels with EAZ_FLAG_PUCT)."""
    raise EazError("the qtransform is evaluated inside the fused search; pass this function (or a partial of it) as qtransform=")


def epistemic_muzero_policy(params, rng_key, root: EpistemicRootFnOutput, recurrent_fn, num_simulations: int, invalid_actions=None,
                            max_depth=None, *, qtransform=qtransform_by_parent_and_siblings, dirichlet_fraction: float = 0.25,
                            dirichlet_alpha: float = 0.3, pb_c_init: float = 1.25, pb_c_base: float = 19652.0, temperature: float = 1.0,
                            return_tree: bool = False, flags: int = _abi.SEARCH_DEFAULT_FLAGS, mlp_mode: int | None = None,
                            dirichlet_noise=None, noise_seed: int = 0) -> PolicyOutput:
    """emctx.epistemic_muzero_policy (named by the task; the reference itself only calls the Gumbel policy): PUCT selection at the
    root and inside the tree, action sampled in proportion to the visit counts, action_weights = visit_probs (mctx
    policies.muzero_policy).  The Dirichlet root noise is drawn here (or passed as `dirichlet_noise` [B,A]) and mixed into the
    prior logits exactly like mctx `_add_dirichlet_noise` / `_get_logits_from_probs`; the final categorical draw uses Gumbel-max
    with noise from `rng_key` (an int seed, torch.Generator or PreDrawnGumbel)."""
    torch = require_cuda()
    kw = {}
    fn = qtransform
    while isinstance(fn, functools.partial):
        kw = {**fn.keywords, **kw}
        fn = fn.func
    if fn is not qtransform_by_parent_and_siblings:
        raise NotImplementedError("epistemic_muzero_policy supports qtransform_by_parent_and_siblings only")
    probs = torch.softmax(root.prior_logits.to(torch.float32), dim=-1)
    if dirichlet_noise is None:
        gen = rng_key if isinstance(rng_key, torch.Generator) else None
        conc = torch.full_like(probs, float(dirichlet_alpha))
        g = torch._standard_gamma(conc, generator=gen) if gen is not None else torch._standard_gamma(conc)
        dirichlet_noise = g / g.sum(-1, keepdim=True)
    noisy = (1.0 - dirichlet_fraction) * probs + dirichlet_fraction * dirichlet_noise.to(probs)
    noisy_logits = torch.log(torch.clamp_min(noisy, torch.finfo(torch.float32).tiny))
    root = EpistemicRootFnOutput(noisy_logits, root.value, root.value_epistemic_variance, root.embedding, root.beta)
    return _run_policy(params, rng_key, root, recurrent_fn, num_simulations, invalid_actions, max_depth,
                       dict(epsilon=float(kw.get("epsilon", 1e-8))), 16, 1.0, return_tree, int(flags) | _abi.FLAG_PUCT, mlp_mode,
                       dict(pb_c_init=float(pb_c_init), pb_c_base=float(pb_c_base), temperature=float(temperature), noise_seed=int(noise_seed)))


def epistemic_gumbel_muzero_policy(params, rng_key, root: EpistemicRootFnOutput, recurrent_fn, num_simulations: int, invalid_actions=None,
                                   max_depth=None, *, qtransform=epistemic_qtransform_completed_by_mix_value,
                                   max_num_considered_actions: int = 16, gumbel_scale: float = 1.0, return_tree: bool = False,
                                   flags: int = _abi.SEARCH_DEFAULT_FLAGS, mlp_mode: int | None = None) -> PolicyOutput:
    return _run_policy(params, rng_key, root, recurrent_fn, num_simulations, invalid_actions, max_depth, _qtransform_kwargs(qtransform),
                       max_num_considered_actions, gumbel_scale, return_tree, flags, mlp_mode, {})


def _run_policy(params, rng_key, root, recurrent_fn, num_simulations, invalid_actions, max_depth, q, max_num_considered_actions, gumbel_scale,
                return_tree, flags, mlp_mode, extra_cfg) -> PolicyOutput:
    from .context import FusedRecurrentFn, as_fc_params

    if not isinstance(recurrent_fn, FusedRecurrentFn):
        raise NotImplementedError("recurrent_fn must come from e_alphazero_b200.context.get_epistemic_recurrent_fn "
                                  "(DeepSea/Subleq + FC net); arbitrary Python recurrent_fns have no CUDA path and there is no fallback")
    torch = require_cuda()
    rf = recurrent_fn
    net = as_fc_params(params, rf)
    B, A = root.prior_logits.shape
    if A != rf.env.num_actions:
        raise EazError(f"prior_logits has {A} actions, env has {rf.env.num_actions}")
    cfg = _abi.default_search_config(
        batch=B, num_simulations=int(num_simulations), max_depth=0 if max_depth is None else int(max_depth),
        max_num_considered_actions=int(max_num_considered_actions), gumbel_scale=float(gumbel_scale), discount=float(rf.discount),
        two_players_game=int(rf.two_players_game), exploration=int(rf.exploration), value_scale=float(q.get("value_scale", 0.1)),
        maxvisit_init=float(q.get("maxvisit_init", 50.0)), rescale_values=int(q.get("rescale_values", True)),
        use_mixed_value=int(q.get("use_mixed_value", True)), epsilon=float(q.get("epsilon", 1e-8)), flags=int(flags),
        mlp_mode=int(rf.mlp_mode if mlp_mode is None else mlp_mode), **extra_cfg)
    dev = root.prior_logits.device
    # the entry keeps rf and net alive, so their id()s stay unique for as long as the key exists (and are re-checked with `is`)
    key = (id(rf), id(net), B, bytes(cfg), bool(return_tree), str(dev))
    hit = _plans.get(key)
    if hit is None or hit[0] is not rf or hit[1] is not net:
        while len(_plans) >= 16:
            _plans.pop(next(iter(_plans)))
        hit = _plans[key] = (rf, net, ops.SearchPlan(cfg, rf.env.spec, net, want_tree=bool(return_tree), device=str(dev)))
    plan = hit[2]
    f32 = lambda t: t.to(dtype=torch.float32).contiguous()
    emb = root.embedding.leaves if hasattr(root.embedding, "leaves") else root.embedding
    beta = root.beta
    if not torch.is_tensor(beta):
        beta = torch.full((B,), float(beta), device=dev)
    rd = dict(prior_logits=f32(root.prior_logits), value=f32(root.value), value_epistemic_variance=f32(root.value_epistemic_variance),
              beta=f32(beta.reshape(B)), embedding=emb, gumbel=_draw_gumbel(rng_key, (B, A), dev))
    if invalid_actions is not None:
        rd["invalid_actions"] = invalid_actions.to(torch.uint8).contiguous()
    out = plan.run(rd, reuse_prepared=None)  # parameter-derived tables are rebuilt only when net.content_token() moved
    out = {k: v.clone() for k, v in out.items()}  # the plan's buffers are reused by the next call
    tree = EpistemicTree(out, A, int(num_simulations))
    return PolicyOutput(action=out["action"], action_weights=out["action_weights"], search_tree=tree)
