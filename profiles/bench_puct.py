"""PUCT (emctx.epistemic_muzero_policy, EAZ_FLAG_PUCT) search time at the C2 / C3 shapes -- a measurement aid.
Usage (on a B200): [EAZ_NO_STAGING=1] python profiles/bench_puct.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from e_alphazero_b200 import _abi, ops
from tests import helpers as H

for kind, kw, B, n in (("deepsea", dict(size=30), 4096, 64), ("subleq", dict(word_size=16), 8192, 64)):
    env = H.make_env(kind, seed=1, **kw)
    net = H.make_net(env, seed=2, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    dst = ops.state_to_device(denv, H.random_states(env, B, seed=3, max_steps=8))
    ev = ops.mlp_forward_states(dnet, denv, dst)
    root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"], beta=torch.linspace(0, 1, B, device="cuda"),
                embedding=dst, gumbel=torch.zeros((B, env.num_actions), device="cuda"))
    cfg = _abi.default_search_config(num_simulations=n, discount=0.97, mlp_mode=_abi.MLP_TENSOR, flags=_abi.SEARCH_DEFAULT_FLAGS | _abi.FLAG_PUCT)
    cfg.batch = B
    plan = ops.SearchPlan(cfg, denv, dnet)
    for _ in range(3):
        plan.run(root)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.run(root, reuse_prepared=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"{kind} PUCT {B} x {n}: {e0.elapsed_time(e1) / 5:.3f} ms / search ({'DIRECT' if os.environ.get('EAZ_NO_STAGING') else 'staged'})")
