"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer / byte / index work and -- in EXACT network mode -- every float are bit-exact."""
import numpy as np
import pytest

from e_alphazero_b200 import _abi
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from e_alphazero_b200 import ops as _ops

    return _ops


def host(t):
    return t.detach().cpu().numpy()


def assert_state_equal(env, dst, st, ctx=""):
    for k in O.state_fields(env):
        H.assert_same_bits(host(dst[k]), st[k], f"{ctx}{k}")


# ----------------------------------------------------------------------------- envs
@pytest.mark.parametrize("N,B", [(4, 33), (10, 257), (30, 1000)])
def test_deepsea_env(ops, N, B):
    rng = np.random.default_rng(N)
    env = H.make_env("deepsea", seed=N, size=N)
    denv = H.device_env(env)
    st, dst = O.env_init(env, B), ops.env_init(denv, B)
    assert_state_equal(env, dst, st, "init ")
    for t in range(2 * N + 3):
        act = rng.integers(0, 2, B).astype(np.int32)
        auto = t % 3 == 2
        st = O.env_step(env, st, act, auto_reset=auto)
        dst = ops.env_step(denv, dst, H.to_device(act), auto_reset=auto)
        assert_state_equal(env, dst, st, f"t={t} ")
        H.assert_same_bits(host(ops.env_observe(denv, dst)), O.env_observe(env, st), f"obs t={t}")
        H.assert_same_bits(host(ops.env_compact(denv, dst)), O.env_compact(env, st), f"compact t={t}")


def test_deepsea_golden(ops, golden_dir):
    import os

    g = np.load(os.path.join(golden_dir, "deepsea.npz"))
    for N in (4, 10, 30):
        denv = ops.deepsea_spec(N, g[f"N{N}_action_map"])
        actions = g[f"N{N}_actions"]
        E, T = actions.shape
        dst = ops.env_init(denv, E)
        for t in range(T + 1):
            assert (host(dst["step_count"]) == g[f"N{N}_step_count"][:, t]).all()
            assert (host(dst["col"]) == g[f"N{N}_col"][:, t]).all()
            assert (host(dst["terminated"]) == g[f"N{N}_terminated"][:, t]).all()
            assert (host(dst["rewards"])[:, 0] == g[f"N{N}_rewards"][:, t]).all()
            assert (host(ops.env_observe(denv, dst)).argmax(1) == g[f"N{N}_obs_index"][:, t]).all()
            if t < T:
                dst = ops.env_step(denv, dst, H.to_device(actions[:, t]))


@pytest.mark.parametrize("ws,binary,reward_fn,B", [(16, True, 0, 515), (16, False, 1, 100), (23, True, 0, 64), (256, True, 0, 40)])
def test_subleq_env(ops, ws, binary, reward_fn, B):
    rng = np.random.default_rng(ws + reward_fn)
    env = H.make_env("subleq", word_size=ws, binary=binary, reward_fn=reward_fn)
    denv = H.device_env(env)
    tasks = rng.integers(0, 9, B).astype(np.int32)  # includes out-of-range ids (lax.switch clamps)
    st, dst = O.env_init(env, B, tasks), ops.env_init(denv, B, tasks)
    assert_state_equal(env, dst, st, "init ")
    steps = 16 if ws <= 32 else 8
    saw_solved = saw_term = False
    for t in range(steps):
        act = np.where(rng.random(B) < 0.5, rng.integers(ws - 4, ws, B), rng.integers(0, ws, B)).astype(np.int32)
        if t == 0:
            act[: B // 4] = ws - 2
        if t == 1:
            act[: B // 4] = ws - 3
        auto = t % 5 == 4
        newt = rng.integers(1, 7, B).astype(np.int32)
        st = O.env_step(env, st, act, auto_reset=auto, task_ids=newt)
        dst = ops.env_step(denv, dst, H.to_device(act), auto_reset=auto, task_ids=newt)
        assert_state_equal(env, dst, st, f"t={t} ")
        H.assert_same_bits(host(ops.env_observe(denv, dst)), O.env_observe(env, st), f"obs t={t}")
        H.assert_same_bits(host(ops.env_compact(denv, dst)), O.env_compact(env, st), f"compact t={t}")
        saw_solved |= bool(st["solved"].any())
        saw_term |= bool(st["terminated"].any())
    assert saw_solved and saw_term


def _loopy_programs(ws, B, rng):
    """Random programs biased towards what makes the interpreter spin: short instruction blocks whose operands and jump targets
    point back into the block (memory cells, IN / OUT / HALT addresses), counters that walk through all residues, plain noise."""
    mem = np.zeros((B, ws), np.int32)
    L = rng.integers(3, min(ws - 4, 30), B)
    for b in range(B):
        n = int(L[b])
        style = b % 4
        if style == 0:    # anything goes
            mem[b, :n] = rng.integers(0, ws, n)
        elif style == 1:  # operands inside the program, jumps to instruction starts
            v = rng.integers(0, max(n, 1), n)
            v[2::3] = 3 * rng.integers(0, max(n // 3, 1), len(v[2::3]))
            mem[b, :n] = v
        elif style == 2:  # IO-heavy loops
            v = rng.integers(0, max(n, 1), n)
            io = rng.random(n) < 0.4
            v[io] = rng.integers(ws - 4, ws, int(io.sum()))
            v[2::3] = 3 * rng.integers(0, max(n // 3, 1), len(v[2::3]))
            mem[b, :n] = v
        else:             # a counter: mem[x] -= mem[y] with a small constant, jump back to 0
            x, y = rng.integers(3, ws - 4, 2)
            mem[b, :3] = (x, y, 0)
            mem[b, y] = rng.integers(1, ws)
            mem[b, x] = rng.integers(0, ws)
            mem[b, 3:6] = rng.integers(0, ws, 3)
    return mem, L


@pytest.mark.parametrize("ws,B", [(16, 24000), (23, 6000), (64, 6000), (256, 3000)])
def test_subleq_cycle_detection_exact(ops, ws, B):
    """The interpreter's exact loop shortcut (common.cuh: a recurring machine state = a loop without exit = the reference's result
    at MAX_CYCLE_COUNT) against the oracle's plain 200-cycle loop (subleq.py:297-299) on programs built to spin, count and do IO."""
    rng = np.random.default_rng(1000 + ws)
    env = H.make_env("subleq", word_size=ws)
    denv = H.device_env(env)
    tasks = rng.integers(1, 7, B).astype(np.int32)
    st = O.env_init(env, B, tasks)
    mem, L = _loopy_programs(ws, B, rng)
    st["memory"][:] = mem
    st["step_count"][:] = np.minimum(L, ws - 5)  # the next action lands right behind the program
    act = np.where(rng.random(B) < 0.5, rng.integers(0, 6, B), rng.integers(0, ws, B)).astype(np.int32)
    exp = O.env_step(env, st, act)
    got = ops.env_step(denv, ops.state_to_device(denv, st), H.to_device(act))
    assert_state_equal(env, got, exp, f"ws={ws} ")
    assert exp["solved"].sum() >= 0 and (exp["input_after"] != st["input_after"]).any()
    # and through the in-tree step of a search (subleq_tree_step_kernel): covered by the search parity tests on these roots
    if ws == 16:
        net = H.make_net(env, seed=5, fill=0.5)
        sub = {k: v[:256] for k, v in st.items()}
        root = H.make_root(env, net, 256, seed=6, states=sub)
        cfg_kw = dict(num_simulations=24, discount=0.97)
        e = O.search(_abi.default_search_config(**cfg_kw), env, net, root, want_tree=True)
        g = ops.search(_abi.default_search_config(**cfg_kw), denv, H.device_net(net), H.device_root(env, denv, root), want_tree=True)
        assert_tree_equal(e, {k: host(v) for k, v in g.items()})


def test_subleq_golden(ops, golden_dir):
    import os

    g = np.load(os.path.join(golden_dir, "subleq_env.npz"))
    for i in range(int(g["num_cases"])):
        ws, binary, rf, task = (int(g[f"c{i}_{k}"]) for k in ("ws", "binary", "reward_fn", "task"))
        denv = ops.subleq_spec(ws, bool(binary), rf)
        dst = ops.env_init(denv, 1, [task])
        acts = g[f"c{i}_actions"]
        for t in range(len(acts) + 1):
            ctx = f"case {i} t={t}"
            assert host(dst["step_count"])[0] == g[f"c{i}__step_count"][t], ctx
            assert host(dst["solved"])[0] == g[f"c{i}__solved"][t], ctx
            assert host(dst["terminated"])[0] == g[f"c{i}_terminated"][t], ctx
            assert (host(dst["memory"])[0] == g[f"c{i}__memory_state"][t]).all(), ctx
            assert (host(dst["input_after"])[0] == g[f"c{i}__example_input_after"][t]).all(), ctx
            assert (host(dst["output_after"])[0] == g[f"c{i}__example_output_after"][t]).all(), ctx
            assert host(dst["rewards"])[0, 0] == g[f"c{i}_rewards"][t, 0], ctx
            assert (host(ops.env_observe(denv, dst))[0] == g[f"c{i}_observation"][t]).all(), ctx
            if t < len(acts):
                dst = ops.env_step(denv, dst, [int(acts[t])])
    tin, tout = ops.subleq_test_cases(4, 100)
    oin, oout = O.subleq_test_cases(4, 100)
    assert (tin == oin).all() and (tout == oout).all()


# ----------------------------------------------------------------------------- hash
@pytest.mark.parametrize("B,D,bits", [(1, 4, 24), (77, 100, 24), (9, 10000, 24), (33, 160, 16), (5, 2592, 32)])
def test_xxhash(ops, B, D, bits):
    rng = np.random.default_rng(D)
    x = np.where(rng.random((B, D)) < 0.5, rng.standard_normal((B, D)), (rng.random((B, D)) < 0.5)).astype(np.float32)
    idx = host(ops.xxhash_indices(H.to_device(x), bits)).view(np.uint32)
    assert (idx == O.xxhash_indices(x, bits)).all()
    if bits >= 8:
        bset = np.zeros(1 << (bits - 3) if bits <= 24 else 1 << 21, np.uint8)
        b24 = min(bits, 24)
        dset = H.to_device(bset)
        ops.hash_update_(H.to_device(x[: B // 2 + 1]), dset, b24)
        O.hash_update(x[: B // 2 + 1], bset, b24)
        assert (host(dset) == bset).all()
        assert (host(ops.hash_lookup(H.to_device(x), dset, b24)) == O.hash_lookup(x, bset, b24)).all()


def test_xxhash_golden(ops, golden_dir):
    import os

    g = np.load(os.path.join(golden_dir, "xxhash.npz"))
    for i in range(int(g["num"])):
        idx = host(ops.xxhash_indices(H.to_device(g[f"h{i}_x"]), int(g[f"h{i}_bits"]))).view(np.uint32)
        assert (idx == g[f"h{i}_idx"]).all(), i


# ----------------------------------------------------------------------------- network
@pytest.mark.parametrize("kind,kw,B", [("deepsea", dict(size=10), 70), ("deepsea", dict(size=30), 33), ("deepsea", dict(size=100), 40),
                                       ("subleq", dict(word_size=16), 100), ("subleq", dict(word_size=16, binary=False), 40),
                                       ("subleq", dict(word_size=40), 37), ("subleq", dict(word_size=256), 33)])
def test_network_exact(ops, kind, kw, B):
    env = H.make_env(kind, seed=1, **kw)
    net = H.make_net(env, seed=2, fill=0.5)
    st = H.random_states(env, B, seed=3)
    exp = O.mlp_forward_states(net, env, st)
    denv, dnet = H.device_env(env), H.device_net(net)
    dst = ops.state_to_device(denv, st)
    got = ops.mlp_forward_states(dnet, denv, dst)
    for k in exp:
        H.assert_same_bits(host(got[k]), exp[k], f"states {k}")
    got = ops.mlp_forward(dnet, ops.env_observe(denv, dst))
    for k in exp:
        H.assert_same_bits(host(got[k]), exp[k], f"dense {k}")
    assert 0 < exp["novelty"].sum() < B


def test_network_golden(ops, golden_dir):
    import os

    g = np.load(os.path.join(golden_dir, "fcnet.npz"))
    for tag, A, hash_io, ws in (("ds10", 2, 0, 0), ("sub16", 16, 1, 16)):
        w = [[g[f"{tag}_w{h * 3 + l}"] for l in range(3)] for h in range(4)]
        b = [[g[f"{tag}_b{h * 3 + l}"] for l in range(3)] for h in range(4)]
        bset = np.zeros(1 << 21, np.uint8)
        bset[g[f"{tag}_set_idx"]] = g[f"{tag}_set_val"]
        dnet = ops.FcParams.from_numpy(w, b, bset, A, 24, hash_io, ws)
        got = ops.mlp_forward(dnet, H.to_device(g[f"{tag}_obs"]))
        assert (host(got["novelty"]) == g[f"{tag}_novelty"]).all()
        for k, gk in (("exploit_logits", "exploit"), ("explore_logits", "explore"), ("value", "value"), ("ube", "ube")):
            np.testing.assert_allclose(host(got[k]), g[f"{tag}_{gk}"], rtol=1e-5, atol=2e-6, err_msg=f"{tag} {k}")


# ----------------------------------------------------------------------------- search
def run_both(ops, env, net, root, cfg_kw):
    cfg = _abi.default_search_config(**cfg_kw)
    exp = O.search(cfg, env, net, root, want_tree=True)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg2 = _abi.default_search_config(**cfg_kw)
    got = ops.search(cfg2, denv, dnet, H.device_root(env, denv, root), want_tree=True)
    return exp, {k: host(v) for k, v in got.items()}


def assert_tree_equal(exp, got):
    for name, _, _ in _abi.SUMMARY_FIELDS + _abi.TREE_FIELDS:
        H.assert_same_bits(got[name], exp[name], name)


SEARCH_CASES = [
    # BASELINE config C1: DeepSea size=10, 64 envs, 32 simulations
    ("deepsea", dict(size=10), 64, dict(num_simulations=32, discount=0.997), dict(beta_max=1.0)),
    ("deepsea", dict(size=10), 64, dict(num_simulations=32, discount=0.997, rescale_values=0, exploration=1), dict(beta_max=0.0)),
    ("deepsea", dict(size=30), 130, dict(num_simulations=64, discount=0.997), dict(beta_max=1.0)),
    ("deepsea", dict(size=6), 50, dict(num_simulations=40, discount=0.997, max_depth=3, gumbel_scale=0.0), dict(beta_max=0.5)),
    ("deepsea", dict(size=10), 40, dict(num_simulations=16, discount=0.9, flags=_abi.FLAG_BACKUP_STD | _abi.FLAG_BETA_FINAL, two_players_game=1),
     dict(beta_max=2.0, invalid_frac=0.3)),
    ("subleq", dict(word_size=16), 96, dict(num_simulations=32, discount=0.97), dict(beta_max=1.0)),
    ("subleq", dict(word_size=16), 48, dict(num_simulations=64, discount=0.97, max_num_considered_actions=4), dict(beta_max=0.0, invalid_frac=0.4)),
    ("subleq", dict(word_size=20, binary=False, reward_fn=1), 20, dict(num_simulations=24, discount=0.97), dict(beta_max=1.0)),
    ("subleq", dict(word_size=40), 24, dict(num_simulations=20, discount=0.97), dict(beta_max=1.0)),
    ("subleq", dict(word_size=256), 10, dict(num_simulations=40, discount=0.97), dict(beta_max=1.0)),
]


@pytest.mark.parametrize("kind,kw,B,cfg_kw,root_kw", SEARCH_CASES)
def test_search_bit_exact(ops, kind, kw, B, cfg_kw, root_kw):
    env = H.make_env(kind, seed=5, **kw)
    net = H.make_net(env, seed=6, fill=0.5)
    root = H.make_root(env, net, B, seed=8, **root_kw)
    exp, got = run_both(ops, env, net, root, cfg_kw)
    assert_tree_equal(exp, got)
    n = cfg_kw["num_simulations"]
    assert (got["visit_counts"].sum(1) == n).all()
    assert (got["node_visits"][:, 0] == n + 1).all()


@pytest.mark.parametrize("kind,kw,B,n", [("deepsea", dict(size=8), 96, 24), ("subleq", dict(word_size=16), 40, 20)])
def test_search_assumption_switches_all_combinations(ops, kind, kw, B, n):
    """the four emctx assumptions are switches (EAZ_FLAG_BETA_INTERIOR / BETA_RAW / BETA_FINAL / BACKUP_STD).
    All 16 combinations, x {mixed value on / off}, must agree with the oracle bit for bit, and the switches must not be dead:
    different combinations produce different trees."""
    env = H.make_env(kind, seed=31, **kw)
    net = H.make_net(env, seed=32, fill=0.5)
    root = H.make_root(env, net, B, seed=33, beta_max=2.0, invalid_frac=0.15)
    seen = set()
    for combo in range(16):
        flags = ((_abi.FLAG_BETA_INTERIOR if combo & 1 else 0) | (_abi.FLAG_BETA_RAW if combo & 2 else 0) |
                 (_abi.FLAG_BETA_FINAL if combo & 4 else 0) | (_abi.FLAG_BACKUP_STD if combo & 8 else 0))
        for mixed in (1, 0):
            cfg_kw = dict(num_simulations=n, discount=0.97, flags=flags, use_mixed_value=mixed, rescale_values=combo & 1)
            exp, got = run_both(ops, env, net, root, cfg_kw)
            assert_tree_equal(exp, got)
            if mixed:
                seen.add((got["node_visits"].tobytes(), got["node_values_epistemic_variance"].tobytes(), got["action_weights"].tobytes()))
    assert len(seen) >= 12, f"only {len(seen)} distinct results over 16 switch combinations"


def test_search_summary_only_and_plan_reuse(ops):
    env = H.make_env("deepsea", seed=5, size=10)
    net = H.make_net(env, seed=6, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(batch=64, num_simulations=32)
    plan = ops.SearchPlan(cfg, denv, dnet, want_tree=False)
    for seed in (1, 2):
        root = H.make_root(env, net, 64, seed=seed)
        exp = O.search(_abi.default_search_config(num_simulations=32), env, net, root, want_tree=False)
        got = plan.run(H.device_root(env, denv, root))
        for name, _, _ in _abi.SUMMARY_FIELDS:
            H.assert_same_bits(host(got[name]), exp[name], name)


# ----------------------------------------------------------------------------- tensor-core network mode
TENSOR_CASES = [
    ("deepsea", dict(size=10), 64, dict(num_simulations=32, discount=0.997), dict(beta_max=1.0)),
    ("deepsea", dict(size=30), 300, dict(num_simulations=64, discount=0.997, exploration=1), dict(beta_max=1.0)),
    ("subleq", dict(word_size=16), 200, dict(num_simulations=32, discount=0.97), dict(beta_max=1.0)),
    ("subleq", dict(word_size=20, binary=False), 40, dict(num_simulations=16, discount=0.97), dict(beta_max=0.0)),
    ("subleq", dict(word_size=256), 20, dict(num_simulations=24, discount=0.97, exploration=1), dict(beta_max=1.0)),
]


@pytest.mark.parametrize("kind,kw,B,cfg_kw,root_kw", TENSOR_CASES)
def test_search_tensor_mode(ops, kind, kw, B, cfg_kw, root_kw):
    """mlp_mode=TENSOR (tcgen05 3xTF32): (a) every node's network outputs agree with the fp32 oracle network to 1e-5;
    (b) the tree is bit-identical to the oracle search replayed with the GPU's own per-node network outputs."""
    env = H.make_env(kind, seed=11, **kw)
    net = H.make_net(env, seed=12, fill=0.5)
    root = H.make_root(env, net, B, seed=13, **root_kw)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(mlp_mode=_abi.MLP_TENSOR, **cfg_kw)
    got = {k: host(v) for k, v in ops.search(cfg, denv, dnet, H.device_root(env, denv, root), want_tree=True).items()}
    n, N, A = cfg_kw["num_simulations"], cfg_kw["num_simulations"] + 1, env.num_actions
    # (a) network accuracy on the expanded nodes
    emb = got["embeddings"][:, 1:].reshape(B * n, -1)
    st = H.uncompact(env, emb)
    ev = O.mlp_forward_states(net, env, st)
    lg = ev["explore_logits"] if cfg_kw.get("exploration") else ev["exploit_logits"]
    lg = lg - lg.max(1, keepdims=True)
    term = st["terminated"].astype(bool)
    np.testing.assert_allclose(got["children_prior_logits"][:, 1:].reshape(B * n, A), lg, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["raw_values"][:, 1:].reshape(-1), np.where(term, 0, ev["value"]), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["raw_values_epistemic_variance"][:, 1:].reshape(-1), np.where(term, 0, ev["ube"]), rtol=1e-5, atol=2e-6)
    # (b) search logic bit-exact under the GPU's own network outputs
    replay = dict(states=got["embeddings"], logits=got["children_prior_logits"], value=got["raw_values"], var=got["raw_values_epistemic_variance"])
    exp = O.search(_abi.default_search_config(**cfg_kw), env, None, root, want_tree=True, replay=replay)
    assert exp["replay_misses"] == 0
    assert_tree_equal(exp, got)


def test_tensor_mode_one_hot_subleq_repeated(ops):
    """Regression: with the NON-binary (one-hot) Subleq encoding the tensor-core network kernel's two producer groups used to build a row's
    observation bit-string concurrently (zero the words, then OR bits in), and on some boxes ~15 % of the searches evaluated a few rows with
    one input bit dropped (logits off by ~3e-2).  60 searches, every node's network outputs against the fp32 oracle each time."""
    env = H.make_env("subleq", seed=11, word_size=20, binary=False)
    net = H.make_net(env, seed=12, fill=0.5)
    B, n = 40, 16
    root = H.make_root(env, net, B, seed=13, beta_max=0.0)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(mlp_mode=_abi.MLP_TENSOR, num_simulations=n, discount=0.97)
    cfg.batch = B
    plan = ops.SearchPlan(cfg, denv, dnet, want_tree=True)
    droot = H.device_root(env, denv, root)
    A = env.num_actions
    for rep in range(60):
        got = {k: host(v) for k, v in plan.run(droot).items() if k in ("embeddings", "children_prior_logits", "raw_values", "raw_values_epistemic_variance")}
        st = H.uncompact(env, got["embeddings"][:, 1:].reshape(B * n, -1))
        ev = O.mlp_forward_states(net, env, st)
        lg = ev["exploit_logits"] - ev["exploit_logits"].max(1, keepdims=True)
        term = st["terminated"].astype(bool)
        np.testing.assert_allclose(got["children_prior_logits"][:, 1:].reshape(B * n, A), lg, rtol=1e-5, atol=2e-6, err_msg=f"search {rep}")
        np.testing.assert_allclose(got["raw_values"][:, 1:].reshape(-1), np.where(term, 0, ev["value"]), rtol=1e-5, atol=2e-6, err_msg=f"search {rep}")
        np.testing.assert_allclose(got["raw_values_epistemic_variance"][:, 1:].reshape(-1), np.where(term, 0, ev["ube"]), rtol=1e-5, atol=2e-6,
                                   err_msg=f"search {rep}")


def _wide_range_net(env, seed, w_big):
    """Sparse network with |w| up to `w_big` in layers 2 / 3 (two non-zeros per output column), small dense layer 1."""
    net = H.make_net(env, seed=seed, fill=0.5)
    rng = np.random.default_rng(seed + 100)
    for h in range(4):
        net.w[h][0] *= 0.25  # hidden layer 1 stays below 1, so layer 2 (two terms of up to 1e3) stays inside the split's range
        if h < _abi.HEAD_EXPLOIT:
            net.w[h][2] *= 0.01  # value / UBE heads: keep the pre-tanh output O(1) so that the comparison is not saturated
        for l in ((1, 2) if h >= _abi.HEAD_EXPLOIT else (1,)):
            w = net.w[h][l]
            K, N = w.shape
            w[:] = 0
            for n in range(N):
                rows = rng.choice(K, 2, replace=False)
                w[rows, n] = rng.uniform(-w_big, w_big, 2).astype(np.float32)
            w[rng.integers(0, K), rng.integers(0, N)] = w_big  # the extreme itself is present
    return net


def _net_terms_bound(net, env, st, head):
    """sum_k |h2_k| |W3_kn| + |b3_n|: the magnitude the 1e-5 relative bound of the scaled split refers to (cancellation-free)."""
    obs = O.env_observe(env, st).reshape(len(st["step_count"]), -1).astype(np.float32)
    h = obs
    for l in range(2):
        h = np.maximum(h @ net.w[head][l] + net.b[head][l], 0)
    return np.abs(h) @ np.abs(net.w[head][2]) + np.abs(net.b[head][2])


@pytest.mark.parametrize("kind,kw,B,n", [("deepsea", dict(size=10), 150, 24), ("deepsea", dict(size=10), 5000, 8), ("subleq", dict(word_size=16), 64, 16)])
def test_tensor_mode_range_guard(ops, kind, kw, B, n):
    """The scaled 3xFP16 split of mlp_mode TENSOR (north_star tolerance 1e-5): weights up to 1e3 keep the bound (per-matrix
    power-of-two scale from max |w|, tile_weights.cu) and leave the sticky status clean; a network whose hidden activations
    exceed the split's range (|h| > 4094) or whose weights are non-finite is REPORTED by eaz_search_numeric_status
    (persistent kernel B=150, per-simulation kernels B=5000 / Subleq)."""
    env = H.make_env(kind, seed=21, **kw)
    net = _wide_range_net(env, 22, 1000.0)
    root = H.make_root(env, net, B, seed=23)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(mlp_mode=_abi.MLP_TENSOR, num_simulations=n, discount=0.97)
    cfg.batch = B
    plan = ops.SearchPlan(cfg, denv, dnet, want_tree=True)
    got = {k: host(v) for k, v in plan.run(H.device_root(env, denv, root)).items()}
    assert plan.numeric_status() == 0
    A = env.num_actions
    st = H.uncompact(env, got["embeddings"][:, 1:].reshape(B * n, -1))
    ev = O.mlp_forward_states(net, env, st)
    live = ~st["terminated"].astype(bool)
    for name, ref, head in (("raw_values", ev["value"], _abi.HEAD_VALUE), ("raw_values_epistemic_variance", ev["ube"], _abi.HEAD_UBE)):
        bound = 1e-5 * _net_terms_bound(net, env, st, head)[:, 0] + 2e-6
        err = np.abs(got[name][:, 1:].reshape(-1) - ref)
        assert (err[live] <= bound[live]).all(), (name, float((err[live] / bound[live]).max()))
    lg = ev["exploit_logits"]
    bound = 1e-5 * _net_terms_bound(net, env, st, _abi.HEAD_EXPLOIT) + 2e-6
    d = got["children_prior_logits"][:, 1:].reshape(B * n, A) - (lg - lg.max(1, keepdims=True))
    assert (np.abs(d) <= 2 * bound.max(1, keepdims=True)).all()
    # the search logic stays bit-exact under the GPU's own network outputs
    replay = dict(states=got["embeddings"], logits=got["children_prior_logits"], value=got["raw_values"], var=got["raw_values_epistemic_variance"])
    exp = O.search(_abi.default_search_config(num_simulations=n, discount=0.97), env, None, root, want_tree=True, replay=replay)
    assert exp["replay_misses"] == 0
    assert_tree_equal(exp, got)

    # (2) hidden activations beyond 4094: clamped AND reported
    big = H.make_net(env, seed=22, fill=0.5)
    big.b[_abi.HEAD_VALUE][0][:8] = 6000.0  # layer-1 outputs are the (only) activations that are split for the tensor pipe in every kernel
    plan2 = ops.SearchPlan(cfg, denv, H.device_net(big), want_tree=False)
    plan2.run(H.device_root(env, denv, root))
    with pytest.raises(ops.EazError, match="activation"):
        plan2.numeric_status()
    # (3) a non-finite weight
    bad = H.make_net(env, seed=22, fill=0.5)
    bad.w[_abi.HEAD_UBE][1][3, 5] = np.inf
    plan3 = ops.SearchPlan(cfg, denv, H.device_net(bad), want_tree=False)
    plan3.run(H.device_root(env, denv, root))
    with pytest.raises(ops.EazError, match="weight"):
        plan3.numeric_status()
    # rebuilding the tables from a sane model clears the sticky flags
    plan3.net = dnet
    plan3.run(H.device_root(env, denv, root))
    assert plan3.numeric_status() == 0


@pytest.mark.parametrize("kind,kw,B,expl", [("deepsea", dict(size=10), 64, 0), ("subleq", dict(word_size=16), 40, 1)])
def test_search_fused_root(ops, kind, kw, B, expl):
    """prior_logits/value/variance == NULL: the library evaluates the root network itself (selfplay.py:89); with the
    EXACT network this must equal the oracle's forward + search bit for bit."""
    env = H.make_env(kind, seed=21, **kw)
    net = H.make_net(env, seed=22, fill=0.5)
    st = H.random_states(env, B, seed=23)
    ev = O.mlp_forward_states(net, env, st)
    rng = np.random.default_rng(5)
    root = dict(prior_logits=ev["explore_logits"] if expl else ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"],
                beta=np.linspace(0, 1, B).astype(np.float32), embedding=st, gumbel=rng.gumbel(size=(B, env.num_actions)).astype(np.float32))
    kw_cfg = dict(num_simulations=24, discount=0.97, exploration=expl)
    exp = O.search(_abi.default_search_config(**kw_cfg), env, net, root, want_tree=True)
    denv, dnet = H.device_env(env), H.device_net(net)
    droot = H.device_root(env, denv, root)
    for k in ("prior_logits", "value", "value_epistemic_variance"):
        del droot[k]
    got = {k: host(v) for k, v in ops.search(_abi.default_search_config(batch=B, **kw_cfg), denv, dnet, droot, want_tree=True).items()}
    assert_tree_equal(exp, got)
    H.assert_same_bits(got["root_value"], ev["value"], "root_value")
    H.assert_same_bits(got["root_ube"], ev["ube"], "root_ube")


@pytest.mark.parametrize("mlp_mode", [_abi.MLP_EXACT, _abi.MLP_TENSOR])
def test_search_reuse_prepared_tables(ops, mlp_mode):
    """EAZ_FLAG_REUSE_PREPARED: a second search on the same workspace that skips rebuilding the parameter-derived tables
    (weight images, novelty table, seq-halving table) returns exactly what a full search returns."""
    env = H.make_env("deepsea", seed=5, size=10)
    net = H.make_net(env, seed=6, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(batch=64, num_simulations=32, mlp_mode=mlp_mode)
    plan = ops.SearchPlan(cfg, denv, dnet, want_tree=True)
    fresh = ops.SearchPlan(cfg, denv, dnet, want_tree=True)
    for seed in (1, 2, 3):
        root = H.device_root(env, denv, H.make_root(env, net, 64, seed=seed))
        got = {k: host(v).copy() for k, v in plan.run(root, reuse_prepared=seed > 1).items()}  # seeds 2, 3 reuse the tables
        ref = {k: host(v).copy() for k, v in fresh.run(root, reuse_prepared=False).items()}
        for name, _, _ in _abi.SUMMARY_FIELDS + _abi.TREE_FIELDS:
            H.assert_same_bits(got[name], ref[name], f"seed {seed} {name}")


@pytest.mark.parametrize("kind,kw,B", [("deepsea", dict(size=10), 96), ("subleq", dict(word_size=16), 64)])
def test_selfplay_runner_graph_equals_eager(ops, kind, kw, B):
    """SelfplayRunner: CUDA-graph replay (both graph variants: tables rebuilt / reused) == eager launches, given the same noise."""
    import torch

    from e_alphazero_b200.selfplay import SelfplayRunner

    env = H.make_env(kind, seed=3, **kw)
    net = H.make_net(env, seed=4, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    runs = {}
    for mode in ("eager", "graph"):
        r = SelfplayRunner(denv, dnet, B, 16, 0.97, exploration_beta=1.0, directed_exploration=True, mlp_mode=_abi.MLP_TENSOR, use_graph=(mode == "graph"),
                           fused_root=True, seed=1)
        st = ops.env_init(denv, B, task_ids=torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None)
        gen = torch.Generator(device="cuda").manual_seed(9)
        acts = []
        for step in range(5):
            if step == 3:
                r.params_updated()
            u = torch.rand((B, env.num_actions), device="cuda", generator=gen).clamp_(1e-20, 1.0 - 1e-7)
            tasks = torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None
            st, out = r.step(st, gumbel=(-(-u.log()).log()).contiguous(), task_ids=tasks)
            acts.append(host(out.action).copy())
        runs[mode] = (acts, {k: host(v).copy() for k, v in st.items()})
    for a, b in zip(runs["eager"][0], runs["graph"][0]):
        assert (a == b).all()
    for k in runs["eager"][1]:
        H.assert_same_bits(runs["eager"][1][k], runs["graph"][1][k], k)


@pytest.mark.parametrize("kind,kw,B", [("deepsea", dict(size=10), 200), ("subleq", dict(word_size=16), 48)])
def test_selfplay_runner_host_arenas(ops, kind, kw, B):
    """A host caller of a graph runner uploads all state fields with ONE copy of the state arena and reads states / search outputs back with
    one copy per arena (SelfplayRunner.host_arenas, ops.alloc_arena): identical to a runner whose states stay on the device."""
    import torch

    from e_alphazero_b200.selfplay import SelfplayRunner

    env = H.make_env(kind, seed=3, **kw)
    net = H.make_net(env, seed=4, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    tasks = torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None
    runs = {}
    for mode in ("device", "host"):
        r = SelfplayRunner(denv, dnet, B, 12, 0.97, exploration_beta=1.0, directed_exploration=True, mlp_mode=_abi.MLP_TENSOR, use_graph=True,
                           fused_root=True, seed=1)
        st = ops.env_init(denv, B, task_ids=tasks)
        gen = torch.Generator(device="cuda").manual_seed(9)
        rec, arenas = [], None
        for step in range(4):
            u = torch.rand((B, env.num_actions), device="cuda", generator=gen).clamp_(1e-20, 1.0 - 1e-7)
            g = (-(-u.log()).log()).contiguous()
            if mode == "host" and step > 0:  # the previous step's states come back from the host with one copy
                r.static_flat.copy_(arenas[0], non_blocking=True)
                st = r.static_states()
            st, out = r.step(st, gumbel=g, task_ids=tasks)
            if mode == "host":
                if arenas is None:
                    arenas = r.host_arenas()
                    assert arenas[0].is_pinned() and arenas[0].numel() == r.static_flat.numel() and arenas[2].numel() == r.plan.out_flat.numel()
                sf, sv, of, ov = arenas
                sf.copy_(r.static_flat, non_blocking=True)
                of.copy_(r.plan.out_flat, non_blocking=True)
                torch.cuda.synchronize()
                rec.append(({k: sv[k].numpy().copy() for k in sv}, ov["action"].numpy().copy(), ov["value"].numpy().copy(), ov["root_value"].numpy().copy()))
                r.static_flat.fill_(0xAB)  # the device copy is gone: only the upload above can restore it
            else:
                rec.append(({k: host(v).copy() for k, v in st.items()}, host(out.action).copy(), host(out.root_value).copy(), host(out.value_prediction).copy()))
        runs[mode] = rec
    for (sa, aa, va, pa), (sb, ab, vb, pb) in zip(runs["device"], runs["host"]):
        for k in sa:
            H.assert_same_bits(sa[k], sb[k].reshape(sa[k].shape), k)
        H.assert_same_bits(aa, ab, "action")
        H.assert_same_bits(va, vb, "value")
        H.assert_same_bits(pa, pb, "root_value")


# ----------------------------------------------------------------------------- reanalyze (reanalyze.py:52-131)
@pytest.mark.parametrize("B,A", [(1, 2), (257, 2), (100, 16), (33, 40), (9, 256)])
def test_reanalyze_targets_bit_exact(ops, B, A):
    rng = np.random.default_rng(A + B)
    vis = rng.integers(0, 3, (B, A)).astype(np.float32)
    args = dict(discount=0.97, exploration_beta=0.8, exploration_ube_target=bool(B % 2), temperature=1.5, action=rng.integers(0, A, B).astype(np.int32),
                qvalues=rng.standard_normal((B, A)).astype(np.float32), qvar=(rng.random((B, A)) ** 2).astype(np.float32), visit_counts=vis,
                value=rng.standard_normal(B).astype(np.float32), value_std=rng.random(B).astype(np.float32),
                next_state_value=rng.standard_normal(B).astype(np.float32), next_rewards=rng.random(B).astype(np.float32),
                next_terminated=(rng.random(B) < 0.3).astype(np.uint8), terminated=(rng.random(B) < 0.2).astype(np.uint8),
                invalid_actions=(rng.random((B, A)) < 0.25).astype(np.uint8))
    args["invalid_actions"][args["invalid_actions"].all(1), 0] = 0  # (all-invalid rows overflow to NaN under temperature > 1, as in the reference)
    exp = O.reanalyze_targets(**args)
    d = {k: H.to_device(v) if isinstance(v, np.ndarray) else v for k, v in args.items()}
    got = ops.reanalyze_targets(d["discount"], d["exploration_beta"], d["exploration_ube_target"], d["temperature"], d["action"], d["qvalues"], d["qvar"],
                                d["visit_counts"], d["value"], d["value_std"], d["next_state_value"], d["next_rewards"], d["next_terminated"],
                                d["terminated"], d["invalid_actions"])
    for k in exp:
        H.assert_same_bits(host(got[k]), exp[k], k)


def test_reanalyze_targets_golden(ops, golden_dir):
    import os

    from tests.test_golden import _reanalyze_case

    g = np.load(os.path.join(golden_dir, "reanalyze.npz"))
    for i in range(int(g["num_cases"])):
        a, exp = _reanalyze_case(g, i)
        d = {k: H.to_device(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v for k, v in a.items()}
        got = ops.reanalyze_targets(d["discount"], d["exploration_beta"], d["exploration_ube_target"], d["temperature"], d["action"], d["qvalues"], d["qvar"],
                                    d["visit_counts"], d["value"], d["value_std"], d["next_state_value"], d["next_rewards"], d["next_terminated"],
                                    d["terminated"], d["invalid_actions"])
        H.assert_same_bits(host(got["value_target"]), exp["value_target"], "value_target")
        H.assert_same_bits(host(got["ube_target"]), exp["ube_target"], "ube_target")
        np.testing.assert_allclose(host(got["exploration_policy_target"]), exp["exploration_policy_target"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind,kw,B,use_graph", [("deepsea", dict(size=10), 70, False), ("subleq", dict(word_size=16), 48, True)])
def test_reanalyze_runner(ops, kind, kw, B, use_graph):
    """reanalyze(): root forward + search + next-state forward + targets on device == the same pipeline on the oracle (EXACT network)."""
    import torch

    from e_alphazero_b200.reanalyze import ReanalyzeRunner

    env = H.make_env(kind, seed=31, **kw)
    net = H.make_net(env, seed=32, fill=0.5)
    first = H.random_states(env, B, seed=33)
    rng = np.random.default_rng(34)
    second = O.env_step(env, O.copy_state(first), rng.integers(0, env.num_actions, B).astype(np.int32))
    gum = rng.gumbel(size=(B, env.num_actions)).astype(np.float32)
    n, gamma, rbeta, ebeta = 16, 0.97, -0.5, 0.6
    # oracle pipeline
    ev = O.mlp_forward_states(net, env, first)
    root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"], beta=np.full(B, rbeta, np.float32),
                embedding=first, gumbel=gum)
    so = O.search(_abi.default_search_config(num_simulations=n, discount=gamma, exploration=0), env, net, root, want_tree=False)
    nv = O.mlp_forward_states(net, env, second)["value"]
    exp = O.reanalyze_targets(gamma, ebeta, True, 1.0, so["action"], so["qvalues"], so["qvalues_epistemic_variance"], so["visit_counts"], so["value"],
                              so["value_epistemic_std"], nv, second["rewards"][:, 0], second["terminated"], first["terminated"], None)
    # device pipeline
    denv, dnet = H.device_env(env), H.device_net(net)
    r = ReanalyzeRunner(denv, dnet, B, n, gamma, reanalyze_beta=rbeta, exploration_beta=ebeta, exploration_ube_target=True, temperature=1.0,
                        mlp_mode=_abi.MLP_EXACT, use_graph=use_graph)
    df, dsec = ops.state_to_device(denv, first), ops.state_to_device(denv, second)
    for _ in range(2):  # second call exercises the table-reuse path / graph replay
        targets, out = r(df, dsec, gumbel=H.to_device(gum))
        torch.cuda.synchronize()
        for k in exp:
            H.assert_same_bits(host(targets[k]), exp[k], k)
        H.assert_same_bits(host(out["action_weights"]), so["action_weights"], "action_weights")


# ----------------------------------------------------------------------------- compact replay ring
@pytest.mark.parametrize("kind,kw", [("deepsea", dict(size=30)), ("subleq", dict(word_size=16)), ("subleq", dict(word_size=256, binary=False))])
def test_uncompact_round_trip_and_replay_ring(ops, kind, kw):
    import torch

    from e_alphazero_b200.replay import CompactReplayRing

    env = H.make_env(kind, seed=41, **kw)
    denv = H.device_env(env)
    B = 40
    st = H.random_states(env, B, seed=42)
    dst = ops.state_to_device(denv, st)
    # decode(encode(state)) == state, every leaf and the observation, bit for bit
    back = ops.env_uncompact(denv, ops.env_compact(denv, dst), dst["rewards"], with_obs=True)
    for k in O.state_fields(env):
        H.assert_same_bits(host(back[k]), st[k], k)
    H.assert_same_bits(host(back["observation"]), O.env_observe(env, st), "observation")
    # ring: sampled pairs are consecutive states of one env, i.e. second == env.step(first, the action that was played)
    ring = CompactReplayRing(denv, 6, B, seed=3)
    rng = np.random.default_rng(43)
    cur, traj = dst, []
    for t in range(9):  # wraps around the 6-slot ring
        ring.add(cur)
        traj.append({k: host(v).copy() for k, v in cur.items()})
        cur = ops.env_step(denv, cur, H.to_device(rng.integers(0, env.num_actions, B).astype(np.int32)), auto_reset=True,
                           task_ids=np.ones(B, np.int32) if kind == "subleq" else None)
    pair = ring.sample(256)
    stored = traj[-6:]
    f, s2 = {k: host(v) for k, v in pair.first.items()}, {k: host(v) for k, v in pair.second.items()}
    keys = O.state_fields(env)
    for i in range(256):
        hits = [(t, b) for t in range(5) for b in range(B) if all((stored[t][k][b] == f[k][i]).all() for k in keys)
                and all((stored[t + 1][k][b] == s2[k][i]).all() for k in keys)]
        assert hits, f"sample {i} is not a stored consecutive pair"


def test_tile_flag_protocol_subprocess():
    """EAZ_TILE_FLAGS=2 (per-tile release/acquire counters instead of grid-wide PDL waits between the tree kernel and the
    DeepSea network kernel; read once per process, hence the subprocess): the tensor-mode search still replays bit-exactly."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EAZ_TILE_FLAGS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu", "-k",
                        "test_search_tensor_mode and deepsea or test_selfplay_runner_graph_equals_eager and deepsea"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


@pytest.mark.parametrize("kind,kw,B,k,mode", [("deepsea", dict(size=10), 300, 2, _abi.MLP_EXACT), ("deepsea", dict(size=10), 700, 4, _abi.MLP_TENSOR),
                                              ("subleq", dict(word_size=16), 260, 3, _abi.MLP_EXACT)])
def test_search_streams_identical(ops, kind, kw, B, k, mode):
    """EAZ_FLAG_STREAMS(k): the batch searched as k concurrent sub-batches on auxiliary streams returns, array for array, what the
    single-stream search returns (ragged last sub-batch included)."""
    env = H.make_env(kind, seed=51, **kw)
    net = H.make_net(env, seed=52, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    root = H.device_root(env, denv, H.make_root(env, net, B, seed=53, invalid_frac=0.2))
    kwc = dict(batch=B, num_simulations=16, discount=0.97, mlp_mode=mode)
    one = {n: host(v).copy() for n, v in ops.search(_abi.default_search_config(**kwc), denv, dnet, root, want_tree=True).items()}
    cfg = _abi.default_search_config(**kwc)
    cfg.flags |= _abi.flag_streams(k)
    many = {n: host(v).copy() for n, v in ops.search(cfg, denv, dnet, root, want_tree=True).items()}
    for name, _, _ in _abi.SUMMARY_FIELDS + _abi.TREE_FIELDS:
        H.assert_same_bits(many[name], one[name], name)


# ----------------------------------------------------------------------------- PUCT (emctx.epistemic_muzero_policy)
PUCT_CASES = [
    ("deepsea", dict(size=10), 64, dict(num_simulations=32, discount=0.997), dict(beta_max=1.0)),
    ("deepsea", dict(size=6), 50, dict(num_simulations=24, discount=0.9, max_depth=4, temperature=0.5, noise_seed=7), dict(beta_max=0.5, invalid_frac=0.3)),
    ("subleq", dict(word_size=16), 40, dict(num_simulations=32, discount=0.97, pb_c_init=2.0, gumbel_scale=0.0), dict(beta_max=1.0)),
    ("subleq", dict(word_size=40), 12, dict(num_simulations=20, discount=0.97, flags=_abi.FLAG_PUCT), dict(beta_max=0.0)),
]


@pytest.mark.parametrize("kind,kw,B,cfg_kw,root_kw", PUCT_CASES)
def test_search_puct_bit_exact(ops, kind, kw, B, cfg_kw, root_kw):
    """EAZ_FLAG_PUCT: PUCT selection (mctx muzero_action_selection + qtransform_by_parent_and_siblings, beta bonus as in the Gumbel
    path) and the muzero_policy output, CUDA vs oracle, every tree array bit for bit."""
    env = H.make_env(kind, seed=61, **kw)
    net = H.make_net(env, seed=62, fill=0.5)
    root = H.make_root(env, net, B, seed=63, **root_kw)
    cfg_kw = dict(cfg_kw)
    cfg_kw["flags"] = cfg_kw.get("flags", _abi.SEARCH_DEFAULT_FLAGS) | _abi.FLAG_PUCT
    exp, got = run_both(ops, env, net, root, cfg_kw)
    assert_tree_equal(exp, got)
    n = cfg_kw["num_simulations"]
    assert (got["visit_counts"].sum(1) == n).all()
    np.testing.assert_array_equal(got["action_weights"], got["visit_probs"])  # muzero_policy: action_weights = visit_probs
    gum = cfg_kw.get("gumbel_scale", 1.0)
    if gum == 0.0:  # greedy draw: the most visited action
        assert (got["visit_counts"][np.arange(B), got["action"]] == got["visit_counts"].max(1)).all()


def test_puct_policy_facade(ops):
    import torch

    from e_alphazero_b200 import context, emctx, pgx

    amap = (np.random.default_rng(0).random((8, 8)) < 0.5).astype(np.uint8)
    env = pgx.DeepSea(size_of_grid=8, action_map=amap)
    onet = H.make_net(H.make_env("deepsea", seed=0, size=8), seed=2, fill=0.5)
    net = H.device_net(onet)
    B = 32
    states = env.init(batch_size=B)
    fwd = context.get_forward_fn(env)
    (ex, _, v, u, _), _ = fwd.apply(net, None, states, is_training=False)
    rf = context.get_epistemic_recurrent_fn(env, fwd, B, exploration=False, discount=0.997, two_players_game=False)
    root = emctx.EpistemicRootFnOutput(prior_logits=ex, value=v, value_epistemic_variance=u, embedding=states, beta=torch.zeros(B, device="cuda"))
    out = emctx.epistemic_muzero_policy(net, 3, root, rf, 16)
    s = out.search_tree.epistemic_summary()
    assert (s.visit_counts.sum(1) == 16).all() and torch.equal(out.action_weights, s.visit_probs)


# ----------------------------------------------------------------------------- edge shapes
@pytest.mark.parametrize("kind,kw,B,n,mode", [
    ("deepsea", dict(size=4), 1, 1, _abi.MLP_EXACT),       # one tree, one simulation
    ("deepsea", dict(size=4), 1, 5, _abi.MLP_TENSOR),      # one row in a 128-row network tile
    ("deepsea", dict(size=10), 129, 8, _abi.MLP_TENSOR),   # one row past a tile boundary
    ("deepsea", dict(size=4), 37, 70, _abi.MLP_EXACT),     # far more simulations than reachable states: deep absorbing chains (DIRECT path, L > 32)
    ("subleq", dict(word_size=16), 1, 3, _abi.MLP_TENSOR),
    ("subleq", dict(word_size=16), 131, 4, _abi.MLP_TENSOR),
])
def test_search_edge_shapes(ops, kind, kw, B, n, mode):
    env = H.make_env(kind, seed=71, **kw)
    net = H.make_net(env, seed=72, fill=0.5)
    root = H.make_root(env, net, B, seed=73, beta_max=1.0)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(batch=B, num_simulations=n, discount=0.97, mlp_mode=mode)
    cfg.flags |= _abi.flag_streams(2)  # fewer tiles than streams must degrade gracefully
    got = {k: host(v) for k, v in ops.search(cfg, denv, dnet, H.device_root(env, denv, root), want_tree=True).items()}
    ocfg = _abi.default_search_config(num_simulations=n, discount=0.97)
    if mode == _abi.MLP_EXACT:
        exp = O.search(ocfg, env, net, root, want_tree=True)
    else:
        replay = dict(states=got["embeddings"], logits=got["children_prior_logits"], value=got["raw_values"], var=got["raw_values_epistemic_variance"])
        exp = O.search(ocfg, env, None, root, want_tree=True, replay=replay)
        assert exp["replay_misses"] == 0
    assert_tree_equal(exp, got)
    assert (got["node_visits"][:, 0] == n + 1).all()


def test_search_rejects_bad_arguments(ops):
    env = H.make_env("deepsea", seed=1, size=4)
    net = H.make_net(env, seed=2, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    root = H.device_root(env, denv, H.make_root(env, net, 8, seed=3))
    from e_alphazero_b200._lib import EazError

    with pytest.raises(EazError):
        ops.search(_abi.default_search_config(batch=8, num_simulations=0), denv, dnet, root)
    with pytest.raises(EazError):
        ops.search(_abi.default_search_config(batch=8, num_simulations=8, max_num_considered_actions=0), denv, dnet, root)


@pytest.mark.parametrize("B,mode,streams,puct", [(6201, _abi.MLP_EXACT, 1, False), (6400, _abi.MLP_TENSOR, 3, False), (6150, _abi.MLP_EXACT, 1, True)])
def test_search_many_trees_two_per_warp(ops, B, mode, streams, puct):
    """Batches of >= 6144 trees switch the DeepSea tree kernel to two trees per warp (tree_step2_kernel): same trees, bit for bit
    (odd batch: the last warp holds one tree)."""
    env = H.make_env("deepsea", seed=81, size=10)
    net = H.make_net(env, seed=82, fill=0.5)
    n = 24
    root = H.make_root(env, net, B, seed=83, beta_max=1.0, invalid_frac=0.1)
    denv, dnet = H.device_env(env), H.device_net(net)
    cfg = _abi.default_search_config(batch=B, num_simulations=n, discount=0.997, mlp_mode=mode)
    if streams > 1:
        cfg.flags |= _abi.flag_streams(streams)
    ocfg = _abi.default_search_config(num_simulations=n, discount=0.997)
    if puct:  # PUCT selection is not staged: both trees of a warp take the DIRECT path
        cfg.flags |= _abi.FLAG_PUCT
        ocfg.flags |= _abi.FLAG_PUCT
    got = {k: host(v) for k, v in ops.search(cfg, denv, dnet, H.device_root(env, denv, root), want_tree=True).items()}
    if mode == _abi.MLP_EXACT:
        exp = O.search(ocfg, env, net, root, want_tree=True)
    else:
        replay = dict(states=got["embeddings"], logits=got["children_prior_logits"], value=got["raw_values"], var=got["raw_values_epistemic_variance"])
        exp = O.search(ocfg, env, None, root, want_tree=True, replay=replay)
        assert exp["replay_misses"] == 0
    assert_tree_equal(exp, got)


def test_facade_param_cache_never_stale(ops):
    """context.as_fc_params: the device copy of a haiku (params, state) pair follows in-place leaf updates, and freshly built
    pytrees never hit an old entry (a cache keyed on bare id()s returned stale weights once CPython reused the addresses)."""
    import gc

    import torch

    from e_alphazero_b200 import context, pgx

    env = pgx.DeepSea(size_of_grid=6)
    oenv = H.make_env("deepsea", seed=1, size=6)

    def pytrees(seed):
        net = H.make_net(oenv, seed=seed, fill=0.3)
        names = ["fc_az_net/linear" + ("" if i == 0 else f"_{i}") for i in range(12)]
        params = {names[h * 3 + l]: {"w": net.w[h][l].copy(), "b": net.b[h][l].copy()} for h in range(4) for l in range(3)}
        state = {"fc_az_net/xxhash32": {"binary_set": net.binary_set.copy()}}
        return net, params, state

    for seed in range(12):  # fresh pytrees every "learner update"; the old ones are freed, so their addresses get reused
        net, params, state = pytrees(seed)
        fc = context.as_fc_params((params, state), env=env)
        assert (fc.w[2][1].cpu().numpy() == net.w[2][1]).all() and (fc.binary_set.cpu().numpy() == net.binary_set).all(), seed
        # in-place update of a numpy leaf: picked up by the next call, version bumped
        v0 = fc.version
        params["fc_az_net/linear_4"]["w"][:] = 0.5
        fc2 = context.as_fc_params((params, state), env=env)
        assert fc2 is fc and fc.version > v0 and bool((fc.w[1][1] == 0.5).all())
        del net, params, state, fc, fc2
        gc.collect()
    # torch leaves: unchanged leaves are not re-copied (version stays), an in-place update is
    net, params, state = pytrees(99)
    tparams = {k: {kk: torch.as_tensor(vv).cuda() for kk, vv in v.items()} for k, v in params.items()}
    tstate = {"fc_az_net/xxhash32": {"binary_set": torch.as_tensor(state["fc_az_net/xxhash32"]["binary_set"]).cuda()}}
    fc = context.as_fc_params((tparams, tstate), env=env)
    v0 = fc.version
    assert context.as_fc_params((tparams, tstate), env=env) is fc and fc.version == v0
    tok = fc.content_token()
    tparams["fc_az_net/linear"]["b"].add_(1.0)
    assert context.as_fc_params((tparams, tstate), env=env).version == v0 + 1 and fc.content_token() != tok
    assert torch.allclose(fc.b[0][0], tparams["fc_az_net/linear"]["b"])


# ----------------------------------------------------------------------------- BASELINE full sizes: size-independent properties
# name: (env kind, env kwargs, envs, simulations, discount, sub-batch streams, sampled trees for the oracle replay)
FULL_SIZE = {
    "c2": ("deepsea", dict(size=30), 4096, 64, 0.997, 3, 48),
    "c3": ("subleq", dict(word_size=16), 8192, 64, 0.97, 1, 48),
    "c3_streams3": ("subleq", dict(word_size=16), 8192, 64, 0.97, 3, 32),  # as bench.py runs it: fused transition + 3 sub-batch streams
    # C4: the per-GPU shard of BASELINE config 4 (DeepSea-100, 65 536 envs over 8 GPUs = 8192 per GPU, 128 simulations):
    # 30 MB layer-1 row table per head, two-trees-per-warp tree kernel, 129-node trees
    "c4": ("deepsea", dict(size=100), 8192, 128, 0.997, 3, 48),
    # C5 corner points of the Subleq sweep (16k-256k envs x 32-256 simulations): the deepest trees and the widest batch
    "c5_64k_n256": ("subleq", dict(word_size=16), 65536, 256, 0.97, 1, 32),
    "c5_256k_n32": ("subleq", dict(word_size=16), 262144, 32, 0.97, 1, 48),
    # the reference's code default word size (main.py:159) at scale: A = 256 actions (32 lanes x 8 slots), 296-byte states
    "ws256": ("subleq", dict(word_size=256), 2048, 64, 0.97, 1, 24),
}


@pytest.mark.parametrize("wl", list(FULL_SIZE))
def test_full_size_properties(ops, wl):
    """BASELINE configs at full size (C2, C3, the C4 per-GPU shard, the corner points of the C5 sweep, ws=256), tensor-core network:
    mctx tree invariants over EVERY tree, network accuracy (1e-5) on the sampled trees' nodes, and bit-exact oracle replay of a
    sample of the trees.  Root network outputs are inputs of the search; for the large batches they come from the (bit-exact,
    test_network_exact) CUDA fp32 network instead of the slower CPU oracle."""
    import torch

    import bench

    kind, kw, B, n, gamma, streams, nsample = FULL_SIZE[wl]
    envp, netp = bench.synth_params(kind, kw, 0)
    if kind == "deepsea":
        env = O.Env.deepsea(envp["size"], envp["action_map"])
    else:
        env = O.Env.subleq(envp["word_size"], True)
    net = O.FcNet(netp["in_dim"], 256, netp["num_actions"], netp["w"], netp["b"], netp["binary_set"], 24, netp["hash_io"], netp["word_size"])
    st = H.random_states(env, B, seed=5)
    denv, dnet = H.device_env(env), H.device_net(net)
    A = env.num_actions
    if B <= 8192 and wl != "c4":
        root = H.make_root(env, net, B, seed=6, beta_max=1.0, states=st)
    else:
        ev = {k: host(v) for k, v in ops.mlp_forward_states(dnet, denv, ops.state_to_device(denv, st)).items()}
        root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"],
                    beta=np.linspace(0, 1, B).astype(np.float32), embedding=st,
                    gumbel=np.random.default_rng(13).gumbel(size=(B, A)).astype(np.float32))
    cfg = _abi.default_search_config(batch=B, num_simulations=n, discount=gamma, exploration=1, mlp_mode=_abi.MLP_TENSOR)
    if streams > 1:
        cfg.flags |= _abi.flag_streams(streams)
    dgot = ops.search(cfg, denv, dnet, H.device_root(env, denv, root), want_tree=True)
    torch.cuda.synchronize()
    pick = np.sort(np.random.default_rng(7).choice(B, nsample, replace=False))
    # ---- invariants over every tree, evaluated on the device (the big shapes hold > 10^8 edges), then cross-checked on the host sample
    nv, cv, ci = dgot["node_visits"], dgot["children_visits"], dgot["children_index"]
    par, afp = dgot["parents"].long(), dgot["action_from_parent"].long()
    assert bool((dgot["visit_counts"].sum(1) == n).all()) and bool((nv[:, 0] == n + 1).all()) and bool((nv[:, 1:] >= 1).all())
    assert bool((nv == 1 + cv.sum(2)).all())                                             # node_visits[parent] = 1 + sum(children_visits)
    idx = torch.arange(1, n + 1, device=nv.device)
    assert bool((par[:, 1:] < idx).all()) and bool((par[:, 1:] >= 0).all())             # node i hangs under an older node
    flat = ci.reshape(B, -1).gather(1, par[:, 1:] * A + afp[:, 1:])
    assert bool((flat == idx).all())                                                     # children_index[parent, action] = child
    assert bool(((ci >= 0).sum((1, 2)) == n).all())                                      # exactly n edges are expanded
    assert bool((dgot["action"] >= 0).all()) and bool((dgot["action"] < A).all())
    assert torch.allclose(dgot["action_weights"].sum(1), torch.ones(B, device=nv.device), rtol=1e-5, atol=1e-6)
    assert bool(torch.isfinite(dgot["qvalues"]).all()) and bool((dgot["qvalues_epistemic_variance"] >= 0).all())
    tpick = torch.as_tensor(pick, device=nv.device)
    got = {k: host(v[tpick]) for k, v in dgot.items()}
    del dgot
    # ---- network accuracy on the sampled trees' expanded nodes (fp32 oracle network, 1e-5)
    emb = got["embeddings"][:, 1:].reshape(nsample * n, -1)
    nst = H.uncompact(env, emb)
    ev = O.mlp_forward_states(net, env, nst)
    lg = ev["explore_logits"] - ev["explore_logits"].max(1, keepdims=True)
    term = nst["terminated"].astype(bool)
    np.testing.assert_allclose(got["children_prior_logits"][:, 1:].reshape(nsample * n, A), lg, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["raw_values"][:, 1:].reshape(-1), np.where(term, 0, ev["value"]), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["raw_values_epistemic_variance"][:, 1:].reshape(-1), np.where(term, 0, ev["ube"]), rtol=1e-5, atol=2e-6)
    # ---- oracle replay (the GPU's own per-node network outputs) on the sample: bit for bit
    sub_root = {k: (v[pick] if k != "embedding" else {kk: vv[pick] for kk, vv in v.items()}) for k, v in root.items()}
    replay = dict(states=got["embeddings"], logits=got["children_prior_logits"], value=got["raw_values"],
                  var=got["raw_values_epistemic_variance"])
    exp = O.search(_abi.default_search_config(num_simulations=n, discount=gamma, exploration=1), env, None, sub_root, want_tree=True, replay=replay)
    assert exp["replay_misses"] == 0
    for name, _, _ in _abi.SUMMARY_FIELDS + _abi.TREE_FIELDS:
        H.assert_same_bits(got[name], exp[name], name)


@pytest.mark.parametrize("kind,kw,steps", [("deepsea", dict(size=6), 10), ("subleq", dict(word_size=16), 6)])
def test_evaluate_mirror(ops, kind, kw, steps):
    """evaluate.py:13-57: greedy episodes with gumbel_scale = 0 -- device loop == the same loop on the oracle (EXACT network)."""
    from e_alphazero_b200.evaluate import evaluate

    env = H.make_env(kind, seed=91, **kw)
    net = H.make_net(env, seed=92, fill=0.5)
    denv, dnet = H.device_env(env), H.device_net(net)
    B, n, gamma = 24, 12, 0.97
    mean, total = evaluate(dnet, denv, B, n, gamma, max_episode_length=steps, exploitation_beta=0.3)
    st = O.env_init(env, B, np.ones(B, np.int32))
    tot = np.zeros(B, np.float32)
    counter = 0
    while not st["terminated"].all() and counter <= steps:
        ev = O.mlp_forward_states(net, env, st)
        root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"], beta=np.full(B, 0.3, np.float32),
                    embedding=st, gumbel=np.zeros((B, env.num_actions), np.float32))
        out = O.search(_abi.default_search_config(num_simulations=n, discount=gamma, gumbel_scale=0.0), env, net, root, want_tree=False)
        st = O.env_step(env, st, out["action"])
        tot += st["rewards"][:, 0]
        counter += 1
    H.assert_same_bits(host(total), tot, "sum_of_rewards")
    assert abs(float(mean) - tot.mean()) < 1e-6


@pytest.mark.parametrize("kind,kw,B,n", [("deepsea", dict(size=6), 9, 600), ("subleq", dict(word_size=16), 5, 330)])
def test_search_many_simulations(ops, kind, kw, B, n):
    """Trees of more than 512 nodes (no staging area) and descents past node 288 (pointer chase through memory instead of registers)."""
    env = H.make_env(kind, seed=101, **kw)
    net = H.make_net(env, seed=102, fill=0.5)
    root = H.make_root(env, net, B, seed=103, beta_max=1.0)
    exp, got = run_both(ops, env, net, root, dict(num_simulations=n, discount=0.97))
    assert_tree_equal(exp, got)
    assert (got["node_visits"][:, 0] == n + 1).all()


# ----------------------------------------------------------------------------- convolutional evaluators
CONVNET_SHAPES = [
    ("resnet", dict(H=8, W=8, Cc=2, A=65), 37),                                   # othello-sized board, the reference's 64 channels x 5 blocks
    ("resnet", dict(H=19, W=19, Cc=16, A=362, num_blocks=2), 5),                   # go-sized board: 113 KB tile per CTA
    ("resnet", dict(H=6, W=7, Cc=2, A=7, resnet_v2=False, num_blocks=3), 130),     # connect-four-sized, BlockV1
    ("resnet", dict(H=3, W=3, Cc=4, A=9, num_channels=32, num_blocks=1), 300),
    ("minatar", dict(H=10, W=10, Cc=4, A=6), 129),
    ("minatar", dict(H=10, W=10, Cc=10, A=3, hidden=32), 17),
]


@pytest.mark.parametrize("kind,kw,B", CONVNET_SHAPES)
def test_convnet_bit_exact(ops, kind, kw, B):
    """eaz_convnet_forward (csrc/convnet.cu) vs the oracle's fixed-order fp32 restatement of resnet.py / minatar.py: every output bit."""
    import torch

    kw = dict(kw)
    k = _abi.CONVNET_RESNET if kind == "resnet" else _abi.CONVNET_MINATAR
    Hh, W, Cc, A = kw.pop("H"), kw.pop("W"), kw.pop("Cc"), kw.pop("A")
    desc = H.random_convnet(k, Hh, W, Cc, A, seed=B, **kw)
    rng = np.random.default_rng(B + 1)
    obs = (rng.random((B, Hh, W, Cc)) < 0.35).astype(np.uint8)
    obs[B // 2:] = obs[: B - B // 2]  # repeated observations hash alike
    exp = O.convnet_forward(desc, obs)
    net = ops.ConvNetParams(desc)
    got = net.forward(torch.as_tensor(obs).cuda())
    for name in ("exploit_logits", "explore_logits", "value", "ube", "novelty"):
        H.assert_same_bits(host(got[name]), exp[name], f"{kind} {name}")
    assert 0 < exp["novelty"].sum() < B


@pytest.mark.parametrize("tag", ["resnet_v2", "resnet_v1", "minatar"])
def test_convnet_golden(ops, golden_dir, tag):
    """... and against the reference's own modules (golden file generated by executing resnet.py / minatar.py), 1e-5."""
    import os

    import torch

    g = np.load(os.path.join(golden_dir, "convnet.npz"))
    desc, obs, exp = H.load_convnet_golden(g, tag)
    got = {k: host(v) for k, v in ops.ConvNetParams(desc).forward(torch.as_tensor(obs).cuda()).items()}
    np.testing.assert_allclose(got["exploit_logits"], exp["exploit"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(got["explore_logits"], exp["explore"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(got["value"], exp["value"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["ube"], exp["ube"], rtol=1e-5, atol=2e-6)
    np.testing.assert_array_equal(got["novelty"], exp["novelty"])


def test_convnet_rejects_bad_arguments(ops):
    desc = H.random_convnet(_abi.CONVNET_RESNET, 19, 19, 17, 362, num_blocks=1)  # 19 * 19 * 17 is not a multiple of 4 (hashes.py:210)
    import torch

    with pytest.raises(ops.EazError, match="multiple of 4"):
        ops.ConvNetParams(desc).forward(torch.zeros((2, 19, 19, 17), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("Hh,W,Cc,A,B,blocks", [(8, 8, 2, 65, 70, 5), (19, 19, 16, 362, 3, 2), (6, 7, 2, 7, 257, 1), (3, 3, 4, 9, 500, 5)])
def test_convnet_tensor_mode(ops, Hh, W, Cc, A, B, blocks):
    """mlp_mode TENSOR: the residual blocks' 64 -> 64 convolutions as tcgen05 implicit GEMMs on the scaled 3xFP16 split
    (conv_tensor_kernel) against the fp32 oracle.  Bound: 1e-5 of the magnitude of the terms -- here the largest |logit| / 1 for the
    tanh-squashed heads -- after up to 10 chained convolutions; the hash novelty stays exact; the range status stays clean."""
    import torch

    desc = H.random_convnet(_abi.CONVNET_RESNET, Hh, W, Cc, A, seed=B, num_blocks=blocks)
    rng = np.random.default_rng(B + 1)
    obs = (rng.random((B, Hh, W, Cc)) < 0.35).astype(np.uint8)
    exp = O.convnet_forward(desc, obs)
    net = ops.ConvNetParams(dict(desc, mlp_mode=_abi.MLP_TENSOR))
    got = {k: host(v) for k, v in net.forward(torch.as_tensor(obs).cuda()).items()}
    assert net.numeric_status() == 0
    for k in ("exploit_logits", "explore_logits"):
        scale = max(1.0, float(np.abs(exp[k]).max()))
        err = float(np.abs(got[k] - exp[k]).max())
        assert err <= 1e-5 * scale, (k, err, scale)
    np.testing.assert_allclose(got["value"], exp["value"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(got["ube"], exp["ube"], rtol=0, atol=1e-5)
    H.assert_same_bits(got["novelty"], exp["novelty"], "novelty")
    # a BatchNorm that blows the activations past 4094 is clamped AND reported
    big = dict(desc, mlp_mode=_abi.MLP_TENSOR)
    big["blocks"] = [dict(b) for b in desc["blocks"]]
    big["blocks"][0] = dict(bn=[dict(desc["blocks"][0]["bn"][0], offset=np.full(64, 9000.0, np.float32)), desc["blocks"][0]["bn"][1]], conv=desc["blocks"][0]["conv"])
    bad = ops.ConvNetParams(big)
    bad.forward(torch.as_tensor(obs).cuda())
    with pytest.raises(ops.EazError, match="activation"):
        bad.numeric_status()
    # networks outside the tensor path are refused, not silently run in fp32
    with pytest.raises(ops.EazError, match="TENSOR"):
        ops.ConvNetParams(dict(H.random_convnet(_abi.CONVNET_MINATAR, 10, 10, 4, 6), mlp_mode=_abi.MLP_TENSOR)).forward(torch.zeros((2, 10, 10, 4), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("tag,env_id", [("resnet_v2", "othello"), ("minatar", "minatar-breakout")])
def test_convnet_forward_facade(ops, golden_dir, tag, env_id):
    """context.get_forward_fn dispatches on the env id like the reference's get_network (context.py:40-82); forward.apply on the haiku
    pytrees of the golden file returns the reference modules' outputs."""
    import os

    import torch

    from e_alphazero_b200 import context

    g = np.load(os.path.join(golden_dir, "convnet.npz"))
    params, state = {}, {}
    for key in g.files:
        if key.startswith(f"{tag}_P|") or key.startswith(f"{tag}_S|"):
            _, mod, name = key.split("|")
            (params if key.startswith(f"{tag}_P|") else state).setdefault(mod, {})[name] = g[key]
    bset = np.zeros(1 << 21, np.uint8)
    bset[g[f"{tag}_set_idx"]] = g[f"{tag}_set_val"]
    state[str(g[f"{tag}_set_mod"])] = {"binary_set": bset}
    obs = g[f"{tag}_obs"]

    class Cfg:
        discount = 0.99
        num_channels, linear_layer_size, max_ube, max_epistemic_variance_reward = 16, 64, 1.0, 1.0

    fwd = context.get_forward_fn(context.BoardEnvSpec(env_id, obs.shape[1:], H.CONVNET_CASES[tag]["num_actions"]), Cfg())
    (ex, xp, v, u, nov), _ = fwd.apply(params, state, torch.as_tensor(obs).cuda(), is_training=False)
    np.testing.assert_allclose(host(ex), g[f"{tag}_exploit"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(host(xp), g[f"{tag}_explore"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(host(v), g[f"{tag}_value"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(host(u), g[f"{tag}_ube"], rtol=1e-5, atol=2e-6)
    np.testing.assert_array_equal(host(nov), g[f"{tag}_novelty"])
    assert fwd._net(params, state) is fwd._net(params, state)  # cached per pytree pair

