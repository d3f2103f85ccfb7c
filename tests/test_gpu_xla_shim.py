"""The XLA-FFI handlers of csrc/xla_ffi_shim.cc, driven through the stand-in call frame of tests/xla_ffi_stub/ with device buffers:
what a JAX caller gets must be bit-identical to the C ABI called directly (and therefore to the oracle)."""
import numpy as np
import pytest

from e_alphazero_b200 import _abi
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

LEAVES = {_abi.ENV_DEEPSEA: ["step_count", "rewards", "terminated", "truncated", "col"],
          _abi.ENV_SUBLEQ: ["step_count", "rewards", "terminated", "truncated", "memory", "task", "solved", "input_after", "output_after"]}


@pytest.fixture(scope="module")
def shim():
    import torch

    assert torch.cuda.is_available()
    from tests.xla_ffi_stub import harness as X

    return X, X.load()


def env_attrs(env):
    return dict(env_kind=env.kind, size=env.size if env.kind == _abi.ENV_DEEPSEA else 0, word_size=env.word_size,
                binary_encoding=int(env.binary_encoding), reward_fn=int(env.reward_fn))


@pytest.mark.parametrize("kind,kw,B,n,mode", [("deepsea", dict(size=10), 200, 24, _abi.MLP_EXACT), ("deepsea", dict(size=10), 130, 16, _abi.MLP_TENSOR),
                                              ("subleq", dict(word_size=16), 70, 16, _abi.MLP_EXACT)])
def test_shim_search_equals_oracle_and_abi(shim, kind, kw, B, n, mode):
    import torch

    X, lib = shim
    from e_alphazero_b200 import _lib, ops

    env = H.make_env(kind, seed=31, **kw)
    net = H.make_net(env, seed=32, fill=0.5)
    root = H.make_root(env, net, B, seed=33, invalid_frac=0.2 if kind == "subleq" else 0.0)
    denv, dnet = H.device_env(env), H.device_net(net)
    droot = H.device_root(env, denv, root)
    A = env.num_actions
    cfg = _abi.default_search_config(num_simulations=n, discount=0.97, mlp_mode=mode)
    direct = {k: v.clone() for k, v in ops.search(cfg, denv, dnet, droot).items()}

    cfg.batch = B
    e = denv.struct()
    import ctypes as C

    ws_bytes = _lib.load().eaz_search_workspace_bytes(C.byref(cfg), C.byref(e)) + 256
    dev = "cuda"
    f = lambda *s: torch.full(s, float("nan"), dtype=torch.float32, device=dev)
    out = dict(action=torch.zeros(B, dtype=torch.int32, device=dev), action_weights=f(B, A), value=f(B), value_epistemic_std=f(B),
               visit_counts=f(B, A), visit_probs=f(B, A), qvalues=f(B, A), qvalues_epistemic_variance=f(B, A), root_value=f(B), root_ube=f(B))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    dummy = torch.zeros(1, dtype=torch.uint8, device=dev)
    inv = droot.get("invalid_actions")
    amap = denv.action_map if denv.action_map is not None else dummy
    st = droot["embedding"]
    params = [t for h in range(4) for l in range(3) for t in (dnet.w[h][l], dnet.b[h][l])]
    args = [droot["beta"], droot["gumbel"], inv if inv is not None else dummy, amap, dnet.binary_set, droot["prior_logits"], droot["value"],
            droot["value_epistemic_variance"], ws] + [st[k] for k in LEAVES[env.kind]] + params
    rets = list(out.values()) + [ws]
    attrs = dict(env_attrs(env), num_simulations=n, max_depth=0, max_num_considered_actions=16, gumbel_scale=1.0, discount=0.97,
                 two_players_game=0, exploration=0, value_scale=0.1, maxvisit_init=50.0, rescale_values=1, flags=_abi.SEARCH_DEFAULT_FLAGS,
                 mlp_mode=mode, fused_root=0, draw_gumbel=0, noise_seed=0, reuse_prepared=0, hash_bits=net.hash_bits, hash_io=net.hash_io,
                 max_u=1.0, novelty_scale=1.0)
    stream = torch.cuda.current_stream().cuda_stream
    rc, msg = X.call(lib, "EazSearch", [X.tbuf(t) for t in args], [X.tbuf(t) for t in rets], attrs, stream)
    assert rc == 0, msg
    torch.cuda.synchronize()
    for name, _, _ in _abi.SUMMARY_FIELDS:
        H.assert_same_bits(out[name].cpu().numpy(), direct[name].cpu().numpy(), f"shim vs ABI {name}")
    if mode == _abi.MLP_EXACT:
        exp = O.search(_abi.default_search_config(num_simulations=n, discount=0.97), env, net, root, want_tree=False)
        for name, _, _ in _abi.SUMMARY_FIELDS:
            H.assert_same_bits(out[name].cpu().numpy(), exp[name], f"shim vs oracle {name}")
    # second call on the aliased workspace with reuse_prepared (the tables of the first call are still in it): same results
    out2 = {k: torch.zeros_like(v) for k, v in out.items()}
    rc, msg = X.call(lib, "EazSearch", [X.tbuf(t) for t in args], [X.tbuf(t) for t in list(out2.values()) + [ws]], dict(attrs, reuse_prepared=1), stream)
    assert rc == 0, msg
    torch.cuda.synchronize()
    for name, _, _ in _abi.SUMMARY_FIELDS:
        H.assert_same_bits(out2[name].cpu().numpy(), out[name].cpu().numpy(), f"reuse {name}")
    # ... and reuse_prepared without the alias is refused
    ws2 = torch.zeros_like(ws)
    rc, msg = X.call(lib, "EazSearch", [X.tbuf(t) for t in args], [X.tbuf(t) for t in list(out2.values()) + [ws2]], dict(attrs, reuse_prepared=1), stream)
    assert rc == 3 and "input_output_aliases" in msg


@pytest.mark.parametrize("kind,kw,B", [("deepsea", dict(size=8), 300), ("subleq", dict(word_size=16), 257)])
def test_shim_env_step_and_forward(shim, kind, kw, B):
    import torch

    X, lib = shim
    from e_alphazero_b200 import ops

    env = H.make_env(kind, seed=41, **kw)
    net = H.make_net(env, seed=42, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    st0 = H.random_states(env, B, seed=43)
    rng = np.random.default_rng(44)
    act = rng.integers(0, env.num_actions, B).astype(np.int32)
    tasks = rng.integers(1, 4, B).astype(np.int32)
    exp = O.env_step(env, st0, act, auto_reset=True, task_ids=tasks)
    dst = ops.state_to_device(denv, st0)
    direct = ops.env_step(denv, dst, act, auto_reset=True, task_ids=tasks)
    dev = "cuda"
    dummy = torch.zeros(1, dtype=torch.uint8, device=dev)
    amap = denv.action_map if denv.action_map is not None else dummy
    leaves_in = [dst[k] for k in LEAVES[env.kind]]
    leaves_out = [torch.zeros_like(t) for t in leaves_in]
    stream = torch.cuda.current_stream().cuda_stream
    d_act, d_tasks = torch.as_tensor(act, device=dev), torch.as_tensor(tasks, device=dev)  # (kept alive: tbuf only takes the pointer)
    rc, msg = X.call(lib, "EazEnvStep", [X.tbuf(d_act), X.tbuf(d_tasks), X.tbuf(amap)] +
                     [X.tbuf(t) for t in leaves_in], [X.tbuf(t) for t in leaves_out], dict(env_attrs(env), auto_reset=1), stream)
    assert rc == 0, msg
    torch.cuda.synchronize()
    for k, t in zip(LEAVES[env.kind], leaves_out):
        H.assert_same_bits(t.cpu().numpy(), direct[k].cpu().numpy(), f"env_step {k}")
        H.assert_same_bits(t.cpu().numpy().reshape(exp[k].shape), exp[k].astype(t.cpu().numpy().dtype), f"env_step vs oracle {k}")
        H.assert_same_bits(dst[k].cpu().numpy().reshape(st0[k].shape), st0[k], f"input leaf {k} untouched")
    # forward.apply on states
    A = env.num_actions
    f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    outs = [f(B, A), f(B, A), f(B), f(B), f(B), torch.zeros(B * denv.compact_bytes + 16, dtype=torch.uint8, device=dev)]
    params = [t for h in range(4) for l in range(3) for t in (dnet.w[h][l], dnet.b[h][l])]
    rc, msg = X.call(lib, "EazMlpForwardStates", [X.tbuf(amap), X.tbuf(dnet.binary_set)] + [X.tbuf(t) for t in leaves_in] + [X.tbuf(t) for t in params],
                     [X.tbuf(t) for t in outs], dict(env_attrs(env), hash_bits=net.hash_bits, hash_io=net.hash_io, max_u=1.0, novelty_scale=1.0), stream)
    assert rc == 0, msg
    torch.cuda.synchronize()
    ev = O.mlp_forward_states(net, env, st0)
    for t, k in zip(outs, ("exploit_logits", "explore_logits", "value", "ube", "novelty")):
        H.assert_same_bits(t.cpu().numpy(), ev[k], f"forward {k}")
