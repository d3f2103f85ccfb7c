"""CPU oracle vs golden vectors produced by executing the reference's own sources
(oracle/make_golden.py over oracle/jaxshim).  Integer / byte results: bit-exact.
Network outputs: 1e-5 (the goldens come from numpy BLAS matmul, the oracle from
the fp32 FMA-chain contract; north_star tolerance for fp32 is 1e-5 relative)."""
import os

import numpy as np
import pytest

from e_alphazero_b200 import _abi
from oracle import oracle as O
from tests import helpers as H

RTOL = 1e-5
ATOL = 2e-6


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("N", [4, 10, 30])
def test_deepsea_trajectories(golden_dir, N):
    g = load(golden_dir, "deepsea.npz")
    env = O.Env.deepsea(N, g[f"N{N}_action_map"])
    actions = g[f"N{N}_actions"]
    E, T = actions.shape
    st = O.env_init(env, E)
    for t in range(T + 1):
        assert (st["step_count"] == g[f"N{N}_step_count"][:, t]).all()
        assert (st["col"] == g[f"N{N}_col"][:, t]).all()
        assert (st["terminated"] == g[f"N{N}_terminated"][:, t]).all()
        assert (st["rewards"][:, 0] == g[f"N{N}_rewards"][:, t]).all()
        obs = O.env_observe(env, st)
        assert (obs.sum(1) == 1).all()
        assert (obs.argmax(1) == g[f"N{N}_obs_index"][:, t]).all()
        if t < T:
            st = O.env_step(env, st, actions[:, t])


def test_subleq_trajectories(golden_dir):
    g = load(golden_dir, "subleq_env.npz")
    n = int(g["num_cases"])
    assert n >= 60
    for i in range(n):
        ws, binary, rf, task = (int(g[f"c{i}_{k}"]) for k in ("ws", "binary", "reward_fn", "task"))
        env = O.Env.subleq(ws, bool(binary), rf)
        st = O.env_init(env, 1, [task])
        acts = g[f"c{i}_actions"]
        for t in range(len(acts) + 1):
            ctx = f"case {i} ws={ws} task={task} t={t}"
            assert st["step_count"][0] == g[f"c{i}__step_count"][t], ctx
            assert st["solved"][0] == g[f"c{i}__solved"][t], ctx
            assert st["terminated"][0] == g[f"c{i}_terminated"][t], ctx
            assert (st["memory"][0] == g[f"c{i}__memory_state"][t]).all(), ctx
            assert (st["input_after"][0] == g[f"c{i}__example_input_after"][t]).all(), ctx
            assert (st["output_after"][0] == g[f"c{i}__example_output_after"][t]).all(), ctx
            assert st["rewards"][0, 0] == g[f"c{i}_rewards"][t, 0], ctx
            assert (O.env_observe(env, st)[0] == g[f"c{i}_observation"][t]).all(), ctx
            tin, tout = O.subleq_test_cases(task, ws)
            assert (tin == g[f"c{i}_test_in"][t]).all() and (tout == g[f"c{i}_test_out"][t]).all(), ctx
            assert (tin[0] == g[f"c{i}__example_input"][t]).all() and (tout[0] == g[f"c{i}__example_output"][t]).all()
            if t < len(acts):
                st = O.env_step(env, st, [acts[t]])


def test_subleq_simulate(golden_dir):
    g = load(golden_dir, "subleq_simulate.npz")
    outcomes = set()
    for i in range(len(g["ws"])):
        ws = int(g["ws"][i])
        r = O.subleq_simulate(ws, g["memory"][i][:ws], g["tin"][i], g["tout"][i])
        assert (r["input_after"] == g["in_after"][i]).all(), i
        assert (r["output_after"] == g["out_after"][i]).all(), i
        assert [r["bytes_used"], r["cycles_used"], int(r["correct"])] == g["bcc"][i].tolist(), i
        outcomes.add((r["cycles_used"] >= 200, r["correct"]))
    assert (True, False) in outcomes and (False, False) in outcomes  # both timeouts and early errors are covered


def test_subleq_tables_and_encoders(golden_dir):
    g = load(golden_dir, "subleq_tables.npz")
    for ws in (16, 100, 256):
        for task in range(1, 8):
            tin, tout = O.subleq_test_cases(task, ws)
            assert (tin == g[f"ws{ws}_t{task}_in"]).all() and (tout == g[f"ws{ws}_t{task}_out"]).all()
    # docstring examples subleq.py:30-41 / 67-78 via the observation of a crafted state
    env = O.Env.subleq(16, True)
    st = O.env_init(env, 1, [1])
    st["memory"][0, :5] = [1, 3, 5, -1 % 16, 0]
    st["output_after"][0, 0] = 16
    obs = O.env_observe(env, st)[0].reshape(48, 5)
    assert (obs[:4] == g["binary_ws16"][:4]).all() and (obs[16 + 24] == g["binary_ws16"][4]).all()


def test_xxhash(golden_dir):
    g = load(golden_dir, "xxhash.npz")
    for i in range(int(g["num"])):
        idx = O.xxhash_indices(g[f"h{i}_x"], int(g[f"h{i}_bits"]))
        assert (idx == g[f"h{i}_idx"]).all(), i
    x = g["lk_x"]
    bset = np.zeros(1 << 21, np.uint8)
    assert (O.hash_lookup(x, bset) == g["lk_seen0"]).all()
    O.hash_update(x[:8], bset)
    assert (O.hash_lookup(x, bset) == g["lk_seen1"]).all()
    nz = np.flatnonzero(bset)
    assert (nz == g["lk_set_nonzero"]).all() and (bset[nz] == g["lk_set_values"]).all()
    # SURVEY Appendix C known answers
    z = np.zeros((1, 100), np.float32)
    assert O.xxhash_indices(z, 32)[0] == 0xE0BE2238 and O.xxhash_indices(z)[0] == 14728738
    with pytest.raises(ValueError):
        O.xxhash_indices(np.zeros((1, 25), np.float32))  # hashes.py:210


def _net_from_golden(g, tag, A, D, hash_io):
    w = [[g[f"{tag}_w{h * 3 + l}"] for l in range(3)] for h in range(4)]
    b = [[g[f"{tag}_b{h * 3 + l}"] for l in range(3)] for h in range(4)]
    bset = np.zeros(1 << 21, np.uint8)
    bset[g[f"{tag}_set_idx"]] = g[f"{tag}_set_val"]
    return O.FcNet(D, 256, A, w, b, bset, 24, hash_io)


@pytest.mark.parametrize("tag", ["ds10", "sub16"])
def test_fc_network_and_recurrent_fn(golden_dir, tag):
    g = load(golden_dir, "fcnet.npz")
    if tag == "ds10":
        env, A, hash_io, gamma = O.Env.deepsea(10, g["ds10_action_map"]), 2, 0, 0.997
    else:
        env, A, hash_io, gamma = O.Env.subleq(16, True), 16, 1, 0.97
    obs = g[f"{tag}_obs"]
    net = _net_from_golden(g, tag, A, obs.shape[1], hash_io)
    out = O.mlp_forward(net, obs, env.hash_dim(hash_io))
    assert out["novelty"].tolist() == g[f"{tag}_novelty"].tolist()  # exact: 0/1 from the bitset
    assert 0 < out["novelty"].sum() < len(obs)
    for k, gk in (("exploit_logits", "exploit"), ("explore_logits", "explore"), ("value", "value"), ("ube", "ube")):
        np.testing.assert_allclose(out[k], g[f"{tag}_{gk}"], rtol=RTOL, atol=ATOL, err_msg=k)
    # rebuild the states and run one search-free recurrent step through the oracle's search glue:
    B = len(obs)
    st = O.alloc_state(env, B)
    st["terminated"][:] = g[f"{tag}_terminated"]
    st["rewards"][:] = g[f"{tag}_rewards"]
    if tag == "ds10":
        st["step_count"][:], st["col"][:] = g["ds10_step_count"], g["ds10_col"]
    else:
        st["step_count"][:], st["task"][:], st["solved"][:] = g["sub16__step_count"], g["sub16__task"], g["sub16__solved"]
        st["memory"][:], st["input_after"][:], st["output_after"][:] = (g["sub16__memory_state"], g["sub16__example_input_after"],
                                                                        g["sub16__example_output_after"])
    assert (O.env_observe(env, st) == obs).all()
    for expl in (0, 1):
        p = f"{tag}_rf{expl}_"
        act = g[p + "action"]
        # A 1-simulation search whose root prior forces `act`: node 1 then holds the recurrent_fn output.
        prior = np.full((B, A), -50.0, np.float32)
        prior[np.arange(B), act] = 0.0
        cfg = _abi.default_search_config(num_simulations=1, discount=gamma, exploration=expl, gumbel_scale=0.0)
        root = dict(prior_logits=prior, value=np.zeros(B, np.float32), value_epistemic_variance=np.zeros(B, np.float32),
                    beta=np.zeros(B, np.float32), embedding=st, gumbel=np.zeros((B, A), np.float32))
        t = O.search(cfg, env, net, root)
        assert (t["children_index"][np.arange(B), 0, act] == 1).all()
        assert (t["children_rewards"][np.arange(B), 0, act] == g[p + "reward"]).all()
        assert (t["children_discounts"][np.arange(B), 0, act] == g[p + "discount"]).all()
        assert (t["children_rewards_epistemic_variance"] == 0).all() and (g[p + "reward_epistemic_variance"] == 0).all()
        np.testing.assert_allclose(t["children_prior_logits"][:, 1], g[p + "prior_logits"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(t["raw_values"][:, 1], g[p + "value"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(t["raw_values_epistemic_variance"][:, 1], g[p + "value_epistemic_variance"], rtol=RTOL, atol=ATOL)
        nxt = O.env_step(env, st, act)
        assert (nxt["terminated"] == g[p + "terminated"]).all() and (nxt["step_count"] == g[p + "step_count"]).all()
        assert (O.env_observe(env, nxt) == g[p + "obs"]).all()
        assert g[p + "terminated"].any() and not g[p + "terminated"].all()


def test_mask_invalid_actions(golden_dir):
    g = load(golden_dir, "fcnet.npz")
    lg, inv = g["mask_logits"], g["mask_invalid"]
    exp = g["mask_out"]
    got = np.where(inv > 0, np.float32(np.finfo(np.float32).min), lg - lg.max(1, keepdims=True))
    assert (got == exp).all()


def _reanalyze_case(g, i):
    p = f"c{i}_"
    gamma, ebeta, ube_expl, temp = (float(x) for x in g[p + "cfg"])
    args = dict(discount=gamma, exploration_beta=ebeta, exploration_ube_target=bool(ube_expl), temperature=temp, action=g[p + "action"],
                qvalues=g[p + "qvalues"], qvar=g[p + "qvar"], visit_counts=g[p + "visit_counts"], value=g[p + "value"], value_std=g[p + "value_std"],
                next_state_value=g[p + "next_value"], next_rewards=g[p + "next_rewards"], next_terminated=g[p + "next_terminated"],
                terminated=g[p + "terminated"], invalid_actions=g[p + "invalid"])
    exp = {k: g[p + k] for k in ("value_target", "ube_target", "exploration_policy_target")}
    return args, exp


def test_reanalyze_targets(golden_dir):
    """Oracle vs the outputs of the reference's own reanalyze() (reanalyze.py:52-131) run on stubbed search / network
    outputs: value and UBE targets bit for bit, the softmax policy target to 1e-6 (numpy's exp vs the eaz_math.h exp)."""
    g = np.load(os.path.join(golden_dir, "reanalyze.npz"))
    for i in range(int(g["num_cases"])):
        args, exp = _reanalyze_case(g, i)
        got = O.reanalyze_targets(**args)
        assert (got["value_target"].view(np.uint32) == exp["value_target"].view(np.uint32)).all(), i
        assert (got["ube_target"].view(np.uint32) == exp["ube_target"].view(np.uint32)).all(), i
        np.testing.assert_allclose(got["exploration_policy_target"], exp["exploration_policy_target"], rtol=1e-6, atol=1e-7)
        assert (got["exploration_policy_target"][args["invalid_actions"].astype(bool)] == 0).all()


def test_eaz_log_accuracy():
    """include/eaz_math.h eaz_log (used by the PUCT exploration constant and the muzero_policy visit-count logits)."""
    rng = np.random.default_rng(3)
    xs = np.concatenate([np.exp(rng.uniform(-87, 88, 4000)), 1.0 + rng.uniform(-1e-3, 1e-3, 500), [1.0, 2.0, 0.5, 1e-40, 3.4e38, 1.17549435e-38]]).astype(np.float32)
    got = np.array([O.logf(float(x)) for x in xs], np.float32)
    ref = np.log(xs.astype(np.float64))
    ulp = np.abs(got - ref) / np.maximum(np.spacing(np.abs(ref).astype(np.float32)), 1e-45)
    assert ulp.max() <= 1.0, ulp.max()
    assert O.logf(0.0) == -np.inf and np.isnan(O.logf(-1.0)) and O.logf(float("inf")) == np.inf


@pytest.mark.parametrize("kind,kw", [("deepsea", dict(size=8)), ("subleq", dict(word_size=16))])
def test_puct_oracle_tree_invariants(kind, kw):
    """Oracle PUCT search (EAZ_FLAG_PUCT): mctx tree invariants and the muzero_policy outputs."""
    from e_alphazero_b200 import _abi
    from tests import helpers as H

    env = H.make_env(kind, seed=1, **kw)
    net = H.make_net(env, seed=2, fill=0.5)
    B, n = 24, 20
    root = H.make_root(env, net, B, seed=3, beta_max=1.0, invalid_frac=0.2)
    cfg = _abi.default_search_config(num_simulations=n, discount=0.97, gumbel_scale=0.0)
    cfg.flags |= _abi.FLAG_PUCT
    out = O.search(cfg, env, net, root, want_tree=True)
    assert (out["visit_counts"].sum(1) == n).all() and (out["node_visits"][:, 0] == n + 1).all()
    cv, ci = out["children_visits"], out["children_index"]
    assert (out["node_visits"] == np.where(out["node_visits"] > 0, 1 + cv.sum(2), 0)).all()  # node_visits[parent] = 1 + sum(children_visits)
    for b in range(B):
        for node in range(n + 1):
            for a in np.flatnonzero(ci[b, node] >= 0):
                c = ci[b, node, a]
                assert out["parents"][b, c] == node and out["action_from_parent"][b, c] == a
    np.testing.assert_array_equal(out["action_weights"], out["visit_probs"])
    assert (out["visit_counts"][np.arange(B), out["action"]] == out["visit_counts"].max(1)).all()  # gumbel_scale 0: greedy in the visit counts
    inv = root["invalid_actions"].astype(bool)
    some_valid = ~inv.all(1)  # (all actions invalid: masked_argmax falls back to action 0, like mctx)
    assert (out["visit_counts"][some_valid][inv[some_valid]] == 0).all()  # invalid root actions are never selected


def test_seq_halving_schedules_known_answers():
    """considered-visit schedules of mctx seq_halving.get_sequence_of_considered_visits."""
    t = O.seq_halving_table(16, 32)
    assert t[16].tolist() == [0] * 16 + [1] * 8 + [2] * 4 + [3] * 4
    assert t[2].tolist() == [v for v in range(16) for _ in range(2)] and t[2].sum() == 240
    assert t[1].tolist() == list(range(32)) and t[0].tolist() == list(range(32))
    t = O.seq_halving_table(16, 64)
    assert t[16].tolist()[:48] == [0] * 16 + [1] * 8 + [2] * 8 + [3] * 4 + [4] * 4 + [5] * 4 + [6] * 4 and t[16].sum() == 264 and t[2].sum() == 992
    t = O.seq_halving_table(16, 128)
    assert t[16].sum() == 1120 and t[16].max() == 29 and t[2].sum() == 4032
    t = O.seq_halving_table(16, 256)
    assert t[16].sum() == 4608 and t[16].max() == 59


def test_subleq_known_answers():
    """Subleq ws=16 NEGATION_POSITIVE: empty program spins 200 cycles, [14] errors after one cycle, [14, 13]
    (`subleq OUT IN 0`) solves the task; reward on the solving step, termination (with zero reward) one step later."""
    env = O.Env.subleq(16, True)
    st = O.env_init(env, 1, np.array([1], np.int32))
    assert not st["solved"][0] and st["step_count"][0] == 0
    st = O.env_step(env, st, np.array([14], np.int32))
    assert st["rewards"][0, 0] == 0 and not st["solved"][0] and st["output_after"][0].tolist()[:2] == [2, 16]
    st = O.env_step(env, st, np.array([13], np.int32))
    assert st["rewards"][0, 0] == 1 and st["solved"][0] and not st["terminated"][0]
    assert st["output_after"][0].tolist() == [15, 14, 13, 12, 11, 10, 9, 16] and st["input_after"][0].tolist() == [16] * 8
    st = O.env_step(env, st, np.array([0], np.int32))
    assert st["terminated"][0] and st["rewards"][0, 0] == 0
    st2 = O.env_step(env, O.copy_state(st), np.array([5], np.int32))  # absorbing: same state, zero reward
    assert st2["terminated"][0] and st2["rewards"][0, 0] == 0 and (st2["memory"] == st["memory"]).all()


try:
    from hypothesis import given, settings, strategies as hst

    @settings(max_examples=12, deadline=None)
    @given(kind=hst.sampled_from(["deepsea", "subleq"]), n=hst.integers(1, 40), seed=hst.integers(0, 1000), max_depth=hst.sampled_from([0, 0, 3, 7]),
           rescale=hst.booleans(), gscale=hst.sampled_from([0.0, 1.0]), considered=hst.sampled_from([1, 2, 4, 16]))
    def test_gumbel_oracle_tree_invariants(kind, n, seed, max_depth, rescale, gscale, considered):
        """Property test: mctx tree invariants hold for the oracle search under random configurations."""
        from e_alphazero_b200 import _abi
        from tests import helpers as H

        env = H.make_env(kind, seed=seed, **(dict(size=6) if kind == "deepsea" else dict(word_size=16)))
        net = H.make_net(env, seed=seed + 1, fill=0.5)
        B = 6
        root = H.make_root(env, net, B, seed=seed + 2, beta_max=1.0, invalid_frac=0.25)
        cfg = _abi.default_search_config(num_simulations=n, discount=0.97, max_depth=max_depth, rescale_values=int(rescale), gumbel_scale=gscale,
                                         max_num_considered_actions=considered)
        out = O.search(cfg, env, net, root, want_tree=True)
        nv, cv, ci, par, afp = out["node_visits"], out["children_visits"], out["children_index"], out["parents"], out["action_from_parent"]
        assert (out["visit_counts"].sum(1) == n).all() and (nv[:, 0] == n + 1).all()
        assert (par[:, 0] == -1).all() and (afp[:, 0] == -1).all()
        for b in range(B):
            live = np.flatnonzero(nv[b] > 0)
            if max_depth == 0:
                assert live.tolist() == list(range(len(live)))                   # node i is first expanded by simulation i - 1
            # (under a max_depth cut-off a simulation re-expands an existing node and its slot sim + 1 stays empty, as in mctx)
            for i in live[1:]:
                assert 0 <= par[b, i] < i and ci[b, par[b, i], afp[b, i]] == i   # node i was expanded from an older node
            for i in live:
                kids = ci[b, i][ci[b, i] >= 0]
                assert len(set(kids.tolist())) == len(kids)
                if max_depth == 0:
                    assert nv[b, i] == 1 + cv[b, i].sum()                        # node_visits[parent] = 1 + sum(children_visits)
        inv = root["invalid_actions"].astype(bool)
        some_valid = ~inv.all(1)
        assert (out["visit_counts"][some_valid][inv[some_valid]] == 0).all()
        assert not inv[some_valid][np.arange(some_valid.sum()), out["action"][some_valid]].any()
        np.testing.assert_allclose(out["action_weights"].sum(1), 1.0, rtol=1e-5)
        again = O.search(cfg, env, net, root, want_tree=True)                     # deterministic
        assert all((again[k] == out[k]).all() for k in ("action", "children_index", "node_visits"))
except ImportError:  # pragma: no cover
    pass


@pytest.mark.parametrize("tag", ["resnet_v2", "resnet_v1", "minatar"])
def test_convnet_oracle_matches_reference_modules(golden_dir, tag):
    """EpistemicResidualAZNet (v2 as configured by the reference, and v1) / EpistemicMinatarAZNet: the reference's own modules executed
    on the numpy stand-in (oracle/make_golden.py gen_convnet) vs the oracle's fixed-order fp32 restatement."""
    g = np.load(os.path.join(golden_dir, "convnet.npz"))
    desc, obs, exp = H.load_convnet_golden(g, tag)
    got = O.convnet_forward(desc, obs)
    np.testing.assert_allclose(got["exploit_logits"], exp["exploit"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(got["explore_logits"], exp["explore"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(got["value"], exp["value"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["ube"], exp["ube"], rtol=1e-5, atol=2e-6)
    np.testing.assert_array_equal(got["novelty"], exp["novelty"])
    assert 0 < exp["novelty"].sum() < len(exp["novelty"])  # both seen and unseen observations

