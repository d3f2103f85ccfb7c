"""Mirror of the reference's evaluate.py:13-57 on the fused CUDA path: greedy (gumbel_scale = 0) episodes from `env.init`
until every env has terminated or `max_episode_length` steps have passed; plain `env.step` (no auto-reset: terminated envs are
absorbing and contribute zero reward), exploitation logits at the root, the recurrent_fn's exploitation head inside the tree,
beta = config.exploitation_beta.  Returns the mean sum of rewards (evaluate.py:57)."""
from __future__ import annotations

from . import _abi, ops
from ._lib import require_cuda


def evaluate(net: ops.FcParams, env_spec: ops.EnvSpec, num_eval_episodes: int, num_simulations: int, discount: float, max_episode_length: int,
             exploitation_beta: float = 0.0, rescale_values: bool = True, mlp_mode: int = _abi.MLP_EXACT, tasks=(1,), device="cuda", seed: int = 0):
    torch = require_cuda()
    B = int(num_eval_episodes)
    gen = torch.Generator(device=device).manual_seed(seed)
    task_ids = None
    if env_spec.kind == _abi.ENV_SUBLEQ:  # Subleq._init draws the task (subleq.py:624)
        tt = torch.tensor(list(tasks), dtype=torch.int32, device=device)
        task_ids = tt[torch.randint(0, tt.numel(), (B,), device=device, generator=gen)].contiguous()
    states = ops.env_init(env_spec, B, task_ids=task_ids, device=device)  # evaluate.py:52-54
    cfg = _abi.default_search_config(batch=B, num_simulations=int(num_simulations), discount=float(discount), exploration=0, gumbel_scale=0.0,
                                     rescale_values=int(rescale_values), mlp_mode=int(mlp_mode))
    plan = ops.SearchPlan(cfg, env_spec, net, want_tree=False, device=device)
    beta = torch.full((B,), float(exploitation_beta), device=device)  # :34
    gumbel = torch.zeros((B, env_spec.num_actions), device=device)    # scaled by gumbel_scale = 0 (:44)
    total = torch.zeros(B, device=device)
    counter = 0
    while not bool(states["terminated"].all().item()) and counter <= max_episode_length:  # cond_fn :18-20
        out = plan.run(dict(beta=beta, embedding=states, gumbel=gumbel), reuse_prepared=counter > 0)  # fused root: forward.apply + search (:26-45)
        states = ops.env_step(env_spec, states, out["action"])  # :47 (pgx.Env.step: absorbing once terminated)
        total += states["rewards"][:, 0]                        # :48-50 (current_player == 0)
        counter += 1
    return total.mean(), total
