"""Mirror of the reference's reanalyze.py:52-131 on top of the fused CUDA path.

    out = reanalyze(model, config, context, experience_pair, rng_key)

runs, like the reference: root forward on the stored states (exploitation logits, reanalyze.py:67-75), the E-MCTS search
with `context.reanalyze_recurrent_fn`, `config.reanalyze_simulations_per_step` simulations and beta = config.reanalyze_beta
(:70-85), the summary (:86), a second forward on the next observations (:90-92) and the target arithmetic (:87-122, CUDA
kernel `eaz_reanalyze_targets`).  `ReanalyzeRunner` is the pre-allocated, CUDA-graph-capturable form used by the sweeps."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from . import _abi, ops
from ._lib import require_cuda


@dataclass
class ReanalyzeOutput:  # reanalyze.py:43-49
    observation: Any
    next_observation: Any
    value_target: Any
    ube_target: Any
    exploration_policy_target: Any
    exploitation_policy_target: Any


@dataclass
class ExperiencePair:  # flashbax prioritised_flat_buffer.ExperiencePair: .first / .second are pgx States
    first: Any
    second: Any


def _leaves(state):
    return state.leaves if hasattr(state, "leaves") else state


class ReanalyzeRunner:
    """One reanalyze() call for a fixed batch shape: search plan, output buffers and (optionally) a CUDA graph."""

    def __init__(self, env_spec: ops.EnvSpec, net: ops.FcParams, batch: int, num_simulations: int, discount: float, reanalyze_beta: float = 0.0,
                 exploration_beta: float = 0.0, exploration_ube_target: bool = True, temperature: float = 1.0, rescale_values: bool = True,
                 mlp_mode: int = _abi.MLP_EXACT, device="cuda", seed: int = 0, use_graph: bool = False, streams: int = 1):
        torch = require_cuda()
        self.env, self.net, self.B, self.device = env_spec, net, batch, device
        self.A = env_spec.num_actions
        # reanalyze_recurrent_fn: exploration=False (main.py:266-273); root = exploitation logits (reanalyze.py:70-71)
        self.cfg = _abi.default_search_config(batch=batch, num_simulations=num_simulations, discount=discount, exploration=0,
                                              rescale_values=int(rescale_values), mlp_mode=mlp_mode)
        if streams > 1:
            self.cfg.flags |= _abi.flag_streams(streams)
        self.plan = ops.SearchPlan(self.cfg, env_spec, net, want_tree=False, device=device)
        self.beta = torch.full((batch,), float(reanalyze_beta), device=device)  # reanalyze.py:75
        self.tcfg = (float(discount), float(exploration_beta), bool(exploration_ube_target), float(temperature))
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.targets = dict(value_target=torch.empty(batch, device=device), ube_target=torch.empty(batch, device=device),
                            exploration_policy_target=torch.empty((batch, self.A), device=device))
        self.use_graph = use_graph
        self._graphs = {}
        self._static = None
        self._stale = True
        self.launches_per_call = self.plan.num_launches + 3       # + next-state pack & forward, + targets
        self.launches_per_call_reuse = self.plan.num_launches_reuse + 3

    def params_updated(self):
        self._stale = True

    def draw_gumbel(self):
        torch = require_cuda()
        u = torch.rand((self.B, self.A), device=self.device, generator=self.gen).clamp_(1e-20, 1.0 - 1e-7)
        return (-(-u.log()).log()).contiguous()

    def _run(self, first: dict, second: dict, gumbel, invalid, reuse: bool):
        # fused root: forward.apply on the stored states with the exploitation head (cfg.exploration == 0)
        root = dict(beta=self.beta, embedding=first, gumbel=gumbel)
        if invalid is not None:
            root["invalid_actions"] = invalid
        out = self.plan.run(root, reuse_prepared=reuse)
        nxt = ops.mlp_forward_states(self.net, self.env, second)  # reanalyze.py:90-92 (value head of the next observation)
        d, be, ut, tm = self.tcfg
        ops.reanalyze_targets(d, be, ut, tm, out["action"], out["qvalues"], out["qvalues_epistemic_variance"], out["visit_counts"], out["value"],
                              out["value_epistemic_std"], nxt["value"], second["rewards"], second["terminated"], first["terminated"], invalid,
                              out=self.targets)
        return out

    def __call__(self, first: dict, second: dict, gumbel=None, invalid_actions=None):
        """first / second: device state dicts of the sampled transitions.  Returns (targets dict, search outputs)."""
        torch = require_cuda()
        reuse = not self._stale
        if not self.use_graph:
            self._stale = False
            out = self._run(first, second, self.draw_gumbel() if gumbel is None else gumbel, invalid_actions, reuse)
            return self.targets, out
        if self._static is None:
            sf = {k: v.clone() for k, v in first.items()}
            ss = {k: v.clone() for k, v in second.items()}
            sg = torch.zeros((self.B, self.A), dtype=torch.float32, device=self.device)
            si = torch.zeros((self.B, self.A), dtype=torch.uint8, device=self.device)
            self._run(sf, ss, self.draw_gumbel(), si, False)  # warm-up
            torch.cuda.synchronize()
            self._static = (sf, ss, sg, si)
        sf, ss, sg, si = self._static
        draw = gumbel is None
        key = (reuse, draw)
        if key not in self._graphs:
            g = torch.cuda.CUDAGraph()
            g.register_generator_state(self.gen)
            with torch.cuda.graph(g):
                if draw:
                    sg.copy_(self.draw_gumbel())
                out = self._run(sf, ss, sg, si, reuse)
            self._graphs[key] = (g, out)
        g, out = self._graphs[key]
        for dst, srcd in ((sf, first), (ss, second)):
            for k in dst:
                if srcd[k].data_ptr() != dst[k].data_ptr():
                    dst[k].copy_(srcd[k])
        if not draw:
            sg.copy_(gumbel)
        if invalid_actions is None:
            si.zero_()
        else:
            si.copy_(invalid_actions)
        g.replay()
        self._stale = False
        return self.targets, out


_runners: dict = {}


def reanalyze(model, config, context, experience_pair, rng_key=None) -> ReanalyzeOutput:
    """Drop-in for reanalyze.reanalyze (reanalyze.py:52-131).  `config` needs the attributes the reference reads:
    reanalyze_beta, reanalyze_simulations_per_step, discount, exploration_beta, exploration_ube_target,
    exploration_policy_target_temperature (+ optional rescale_q_values_in_search, mlp_mode); `context.env` is an
    e_alphazero_b200.pgx env, `experience_pair.first/.second` pgx States."""
    from .context import as_fc_params
    from .emctx import PreDrawnGumbel

    torch = require_cuda()
    env = context.env
    net = as_fc_params(model, env=env)
    first, second = _leaves(experience_pair.first), _leaves(experience_pair.second)
    B = first["terminated"].shape[0]
    key = (id(env), id(net), B, int(config.reanalyze_simulations_per_step), float(config.discount), float(config.reanalyze_beta),
           float(config.exploration_beta), bool(config.exploration_ube_target), float(config.exploration_policy_target_temperature))
    hit = _runners.get(key)  # (env, net, runner): the strong references keep the id()s in the key unique
    if hit is None or hit[0] is not env or hit[1] is not net:
        while len(_runners) >= 8:
            _runners.pop(next(iter(_runners)))
        hit = _runners[key] = (env, net, ReanalyzeRunner(env.spec, net, B, int(config.reanalyze_simulations_per_step), float(config.discount),
                                            reanalyze_beta=float(config.reanalyze_beta), exploration_beta=float(config.exploration_beta),
                                            exploration_ube_target=bool(config.exploration_ube_target),
                                            temperature=float(config.exploration_policy_target_temperature),
                                            # reanalyze.py:84 passes the bare qtransform: rescale_values stays at its default (True);
                                            # config.rescale_q_values_in_search only reaches the self-play search (selfplay.py:115)
                                            rescale_values=True,
                                            mlp_mode=int(getattr(config, "mlp_mode", _abi.MLP_EXACT)), device=str(first["terminated"].device)))
    r = hit[2]
    tok = net.content_token()
    if getattr(r, "_net_token", None) != tok:  # rebuild the parameter-derived tables only when the parameters moved
        r.params_updated()
        r._net_token = tok
    gumbel = rng_key.gumbel if isinstance(rng_key, PreDrawnGumbel) else None
    targets, out = r(first, second, gumbel=gumbel)  # legal_action_mask is all True for DeepSea / Subleq: no invalid actions
    obs = ops.env_observe(env.spec, first)
    nobs = ops.env_observe(env.spec, second)
    return ReanalyzeOutput(observation=obs, next_observation=nobs, value_target=targets["value_target"].clone(), ube_target=targets["ube_target"].clone(),
                           exploration_policy_target=targets["exploration_policy_target"].clone(), exploitation_policy_target=out["action_weights"].clone())
