"""Generate tests/golden/*.npz by EXECUTING the reference's own source files.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference).
The reference modules are imported unmodified from /root/reference/src on top of
oracle/jaxshim (a numpy stand-in for jax/chex/haiku/pgx, see its README); their
outputs on seeded inputs are stored as golden vectors, and tests/test_golden.py
holds the CPU oracle to them.  Usage:  python oracle/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "jaxshim"), "/root/reference/src"]

import numpy as np  # noqa: E402
import jax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
import haiku as hk  # noqa: E402

from envs.deep_sea import DeepSea  # noqa: E402
from envs import subleq as rsub  # noqa: E402
from network.hashes import XXHash  # noqa: E402
from network.fully_connected import EpistemicFullyConnectedAZNet  # noqa: E402
import context as rctx  # noqa: E402
import reanalyze as rrean  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
KEY = jax.random.PRNGKey(0)


def np_(x, dt=None):
    a = np.asarray(x)
    return a.astype(dt) if dt is not None else a


# ----------------------------------------------------------------------------- DeepSea
def gen_deepsea():
    rng = np.random.default_rng(1)
    out = {}
    for N in (4, 10, 30):
        env = DeepSea(N)
        amap = rng.random((N, N)) < 0.5
        env.action_map = jnp.array(amap)
        E, T = 6, N + 2
        actions = rng.integers(0, 2, size=(E, T)).astype(np.int32)
        rec = {k: np.zeros((E, T + 1), np.int32) for k in ("step_count", "col", "terminated", "obs_index")}
        rec["rewards"] = np.zeros((E, T + 1), np.float32)
        for e in range(E):
            s = env.init(KEY)
            for t in range(T + 1):
                rec["step_count"][e, t] = int(s._step_count)
                rec["col"][e, t] = int(s._horizontal_position)
                rec["terminated"][e, t] = int(s.terminated)
                rec["rewards"][e, t] = float(np_(s.rewards)[0])
                idx = np.flatnonzero(np_(s.observation).reshape(-1))
                assert idx.size == 1 and np_(s.legal_action_mask).all()
                rec["obs_index"][e, t] = idx[0]
                if t < T:
                    s = env.step(s, jnp.int32(actions[e, t]), None)
        out[f"N{N}_action_map"] = amap.astype(np.uint8)
        out[f"N{N}_actions"] = actions
        for k, v in rec.items():
            out[f"N{N}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "deepsea.npz"), **out)
    print("deepsea.npz", len(out))


# ----------------------------------------------------------------------------- Subleq
SUB_STATE = ("_step_count", "_task", "_solved", "terminated", "_memory_state", "_example_input", "_example_output",
             "_example_input_after", "_example_output_after")


def _sub_record(s):
    r = {k: np_(getattr(s, k)).astype(np.int32) for k in SUB_STATE}
    r["rewards"] = np_(s.rewards, np.float32)
    r["observation"] = np_(s.observation).astype(np.uint8).reshape(-1)
    r["test_in"] = np_(s._test_cases[0], np.int32)
    r["test_out"] = np_(s._test_cases[1], np.int32)
    assert np_(s.legal_action_mask).all()
    return r


KNOWN = {  # (ws, task): programs that exercise long/looping/solving executions
    (16, 1): [[14, 13], [14, 13, 0], [13, 14, 0], [14, 14, 14], [13, 13, 13], [12, 12, 12, 12], [3, 13, 0, 14, 3, 0]],
    (16, 2): [[3, 13, 6, 14, 3, 0]], (16, 3): [[14, 13, 0], [4, 13, 3, 14, 4, 0]],
}


def gen_subleq():
    rng = np.random.default_rng(2)
    cases = []
    cfgs = [(16, True, "solved"), (16, False, "solved"), (16, True, "bytes"), (32, True, "solved"), (20, True, "solved"), (256, True, "solved")]
    for ws, binary, rf in cfgs:
        reward_fn = rsub.solved_or_not if rf == "solved" else rsub.lowest_bytes
        tasks = [1, 2, 3, 4, 5, 6] if ws == 16 and binary and rf == "solved" else [1, 3]
        for task in tasks:
            env = rsub.Subleq([rsub.SubleqTask(task) if task <= 14 else task], word_size=ws, reward_fn=reward_fn, use_binary_encoding=binary)
            progs = [list(p) for p in KNOWN.get((ws, task), [])] if (binary and rf == "solved") else [[ws - 2, ws - 3]]
            nrand = 4 if ws <= 32 else 1
            for _ in range(nrand):
                L = int(rng.integers(1, ws - 1))
                L = min(L, 18)
                # bias towards special addresses so IN/OUT paths are hit
                p = np.where(rng.random(L) < 0.4, rng.integers(ws - 4, ws, L), rng.integers(0, ws, L))
                progs.append([int(v) for v in p])
            if ws == 16 and task == 1 and binary and rf == "solved":
                progs.append([0] * 15)  # runs into the step_count >= ws-3 termination
            for prog in progs:
                s = env.init(KEY)
                traj = [_sub_record(s)]
                for a in prog:
                    s = env.step(s, jnp.int32(a), None)
                    traj.append(_sub_record(s))
                cases.append(dict(ws=ws, binary=int(binary), reward_fn=0 if rf == "solved" else 1, task=task,
                                  actions=np.array(prog, np.int32), traj=traj))
            print("subleq", ws, binary, rf, task, len(progs))
    flat = {"num_cases": np.int32(len(cases))}
    for i, c in enumerate(cases):
        for k in ("ws", "binary", "reward_fn", "task"):
            flat[f"c{i}_{k}"] = np.int32(c[k])
        flat[f"c{i}_actions"] = c["actions"]
        for k in c["traj"][0]:
            flat[f"c{i}_{k}"] = np.stack([t[k] for t in c["traj"]])
    np.savez_compressed(os.path.join(OUT, "subleq_env.npz"), **flat)

    # raw simulate() on arbitrary memory images
    sims = {"ws": [], "memory": [], "tin": [], "tout": [], "in_after": [], "out_after": [], "bcc": []}
    for ws in (16, 24, 256):
        for _ in range(60 if ws < 256 else 12):
            mem = np.where(rng.random(ws) < 0.35, rng.integers(ws - 4, ws, ws), rng.integers(0, ws, ws)).astype(np.int32)
            if rng.random() < 0.5:
                mem[int(rng.integers(3, ws)):] = 0
            task = int(rng.integers(1, 7))
            tin, tout = rsub.get_test_cases(jnp.int32(task), ws)
            k = int(rng.integers(0, 3))
            r = rsub.simulate(ws, jnp.array(mem), tin[k], tout[k])
            sims["ws"].append(ws)
            sims["memory"].append(np.pad(mem, (0, 256 - ws)))
            sims["tin"].append(np_(tin[k], np.int32))
            sims["tout"].append(np_(tout[k], np.int32))
            sims["in_after"].append(np_(r.input_after, np.int32))
            sims["out_after"].append(np_(r.output_after, np.int32))
            sims["bcc"].append(np.array([int(r.bytes_used), int(r.cycles_used), int(r.correct)], np.int32))
    np.savez_compressed(os.path.join(OUT, "subleq_simulate.npz"), **{k: np.array(v) for k, v in sims.items()})
    # test-case tables + encoder docstring examples (subleq.py:30-41, 67-78)
    tc = {}
    for ws in (16, 100, 256):
        for task in range(1, 8):
            tin, tout = rsub.get_test_cases(jnp.int32(task), ws)
            tc[f"ws{ws}_t{task}_in"], tc[f"ws{ws}_t{task}_out"] = np_(tin, np.int32), np_(tout, np.int32)
    tc["onehot_ws8"] = np_(rsub.subleq_words_to_observation_one_hot(jnp.array([1, 3, 5, -1, 8]), 8)).astype(np.uint8)
    tc["binary_ws16"] = np_(rsub.subleq_words_to_observation_binary(jnp.array([1, 3, 5, -1, 16]), 16)).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "subleq_tables.npz"), **tc)
    print("subleq_env.npz", len(cases), "subleq_simulate.npz", len(sims["ws"]))


# ----------------------------------------------------------------------------- XXHash
def gen_hash():
    rng = np.random.default_rng(3)
    out = {}
    for i, (B, D, bits) in enumerate([(5, 4, 24), (7, 100, 24), (3, 900, 24), (2, 10000, 24), (6, 160, 24), (6, 240, 16), (4, 64, 32), (4, 2592, 24)]):
        kind = i % 3
        if kind == 0:
            x = rng.standard_normal((B, D)).astype(np.float32)
        elif kind == 1:
            x = (rng.random((B, D)) < 0.3).astype(np.float32)
        else:
            x = np.zeros((B, D), np.float32)
            x[np.arange(B), rng.integers(0, D, B)] = 1.0
        f = hk.without_apply_rng(hk.transform_with_state(lambda v, bits=bits: XXHash(bits_per_hash=bits).get_indices(v)))
        _, st = f.init(None, jnp.array(x))
        idx, _ = f.apply({}, st, jnp.array(x))
        out[f"h{i}_x"], out[f"h{i}_bits"], out[f"h{i}_idx"] = x, np.int32(bits), np_(idx).astype(np.uint32)
    out["num"] = np.int32(8)
    # lookup / update round trip through BaseHash.__call__ / update
    x = (rng.random((16, 100)) < 0.5).astype(np.float32)

    def probe(v, upd):
        h = XXHash()
        if upd:
            h.update(v)
        return h(v)

    f = hk.without_apply_rng(hk.transform_with_state(probe))
    _, st0 = f.init(None, jnp.array(x), False)
    seen0, _ = f.apply({}, st0, jnp.array(x), False)
    _, st1 = f.apply({}, st0, jnp.array(x[:8]), True)
    seen1, _ = f.apply({}, st1, jnp.array(x), False)
    out["lk_x"], out["lk_seen0"], out["lk_seen1"] = x, np_(seen0).astype(np.uint8), np_(seen1).astype(np.uint8)
    out["lk_set_nonzero"] = np.flatnonzero(np_(st1["xxhash32"]["binary_set"])).astype(np.int64)
    out["lk_set_values"] = np_(st1["xxhash32"]["binary_set"])[out["lk_set_nonzero"]]
    np.savez_compressed(os.path.join(OUT, "xxhash.npz"), **out)
    print("xxhash.npz")


# ----------------------------------------------------------------------------- FC net + recurrent_fn
def _rand_params(rng, D, A, H=256):
    params = {}
    for i in range(12):
        head, layer = divmod(i, 3)
        fin = D if layer == 0 else H
        fout = H if layer < 2 else (1 if head < 2 else A)
        name = "fc_az_net/linear" + ("" if i == 0 else f"_{i}")
        params[name] = {"w": jnp.array((rng.standard_normal((fin, fout)).clip(-2, 2) / np.sqrt(fin)).astype(np.float32)),
                        "b": jnp.array((rng.standard_normal(fout) * 0.05).astype(np.float32))}
    return params


def _flat_params(params, out, prefix):
    for i in range(12):
        name = "fc_az_net/linear" + ("" if i == 0 else f"_{i}")
        out[f"{prefix}w{i}"], out[f"{prefix}b{i}"] = np_(params[name]["w"]), np_(params[name]["b"])


class _Cfg:  # the attributes context.get_network reads (context.py:41-82)
    hash_class = "XXHash"
    subleq_hash_only_io = True
    linear_layer_size = 256
    discount = 0.997

    def __init__(self, env_id):
        self.env_id = env_id


def gen_net():
    rng = np.random.default_rng(4)
    out = {}
    setups = [("ds10", DeepSea(10), _Cfg("deep_sea-10"), 0.997),
              ("sub16", rsub.Subleq([rsub.SubleqTask.NEGATION_POSITIVE], word_size=16, use_binary_encoding=True), _Cfg("subleq-16"), 0.97)]
    for tag, env, cfg, gamma in setups:
        if tag == "ds10":
            env.action_map = jnp.array(rng.random((10, 10)) < 0.5)
            out["ds10_action_map"] = np_(env.action_map).astype(np.uint8)
        forward = rctx.get_forward_fn(env, cfg)
        B = 12
        # roll a batch of states forward a random number of steps
        keys = jax.random.split(KEY, B)
        states = jax.vmap(env.init)(keys)
        A = env.num_actions
        for t in range(11 if tag == "ds10" else 9):
            act = rng.integers(0, A, B).astype(np.int32)
            if tag == "sub16":
                act = np.where(rng.random(B) < 0.5, rng.integers(12, 16, B), act).astype(np.int32)
                if t == 0:
                    act[:2] = 14
                if t == 1:
                    act[:2] = 13
            adv = rng.random(B) < (0.85 if tag == "ds10" else 0.7)
            new = jax.vmap(env.step)(states, jnp.array(act), keys)
            states = jax.tree.map(lambda n, o: jnp.where(adv.reshape((-1,) + (1,) * (np.ndim(n) - 1)), n, o), new, states)
        obs = np_(states.observation)
        D = int(np.prod(obs.shape[1:]))
        params = _rand_params(rng, D, A)
        _, st = forward.init(None, jnp.array(obs), is_training=False)
        bset = (rng.random(1 << 21) < 0.5).astype(np.uint8) * 0  # start empty, then insert half of the batch
        st = {"fc_az_net/xxhash32": {"binary_set": jnp.array(bset)}}
        (_, _, _, _, _), st = forward.apply(params, st, jnp.array(obs[: B // 2]), is_training=False, update_hash=True)
        outs, _ = forward.apply(params, st, jnp.array(obs), is_training=False)
        _flat_params(params, out, f"{tag}_")
        bs = np_(st["fc_az_net/xxhash32"]["binary_set"])
        out[f"{tag}_set_idx"] = np.flatnonzero(bs).astype(np.int64)
        out[f"{tag}_set_val"] = bs[out[f"{tag}_set_idx"]]
        out[f"{tag}_obs"] = obs.reshape(B, -1).astype(np.uint8)
        for name, v in zip(("exploit", "explore", "value", "ube", "novelty"), outs):
            out[f"{tag}_{name}"] = np_(v, np.float32)
        # state leaves for the recurrent_fn test
        if tag == "ds10":
            out["ds10_step_count"], out["ds10_col"] = np_(states._step_count, np.int32), np_(states._horizontal_position, np.int32)
        else:
            for k in ("_step_count", "_task", "_solved", "_memory_state", "_example_input_after", "_example_output_after"):
                out[f"sub16_{k}"] = np_(getattr(states, k)).astype(np.int32)
        out[f"{tag}_terminated"] = np_(states.terminated).astype(np.uint8)
        out[f"{tag}_rewards"] = np_(states.rewards, np.float32)
        # context.get_epistemic_recurrent_fn (context.py:109-157), both policy heads
        for expl in (False, True):
            fn = rctx.get_epistemic_recurrent_fn(env, forward, B, expl, gamma, False)
            act = rng.integers(0, A, B).astype(np.int32)
            ro, ns = fn((params, st), KEY, jnp.array(act), states)
            p = f"{tag}_rf{int(expl)}_"
            out[p + "action"] = act
            for k in ("reward", "reward_epistemic_variance", "discount", "prior_logits", "value", "value_epistemic_variance"):
                out[p + k] = np_(getattr(ro, k), np.float32)
            out[p + "terminated"] = np_(ns.terminated).astype(np.uint8)
            out[p + "step_count"] = np_(ns._step_count, np.int32)
            out[p + "obs"] = np_(ns.observation).reshape(B, -1).astype(np.uint8)
    # the two in-tree copies of mctx helpers (reanalyze.py:16-40)
    lg = rng.standard_normal((5, 6)).astype(np.float32)
    inv = rng.random((5, 6)) < 0.3
    out["mask_logits"], out["mask_invalid"] = lg, inv.astype(np.uint8)
    out["mask_out"] = np_(rrean.mask_invalid_actions(jnp.array(lg), jnp.array(inv)), np.float32)
    np.savez_compressed(os.path.join(OUT, "fcnet.npz"), **out)
    print("fcnet.npz", len(out))


# ----------------------------------------------------------------------------- reanalyze targets (reanalyze.py:52-131)
def gen_reanalyze():
    """Runs the reference's reanalyze() itself.  Its two externals are stubbed with seeded arrays: the search
    (emctx.epistemic_gumbel_muzero_policy -- not in /root/reference) returns a fixed PolicyOutput / summary, and
    context.forward.apply returns fixed network outputs, so what is pinned is the target arithmetic :86-129."""
    import types

    import emctx as shim_emctx

    rng = np.random.default_rng(11)
    out = {}
    cases = [(2, 33, 0.997, 0.0, True, 1.0), (16, 21, 0.97, 0.7, False, 0.5), (6, 40, 0.9, 1.5, True, 2.0)]
    for ci, (A, B, gamma, ebeta, ube_expl, temp) in enumerate(cases):
        visits = rng.integers(0, 4, (B, A)).astype(np.float32)
        visits[rng.random(B) < 0.15] = 0  # rows without any visited action
        summ = types.SimpleNamespace(
            qvalues=jnp.array(rng.standard_normal((B, A)).astype(np.float32)),
            qvalues_epistemic_variance=jnp.array((rng.random((B, A)) ** 2).astype(np.float32)),
            visit_counts=jnp.array(visits),
            value=jnp.array(rng.standard_normal(B).astype(np.float32)),
            value_epistemic_std=jnp.array(rng.random(B).astype(np.float32)))
        action = rng.integers(0, A, B).astype(np.int32)
        w = rng.random((B, A)).astype(np.float32)
        pol = types.SimpleNamespace(action=jnp.array(action), action_weights=jnp.array(w / w.sum(1, keepdims=True)),
                                    search_tree=types.SimpleNamespace(epistemic_summary=lambda summ=summ: summ))
        shim_emctx.epistemic_gumbel_muzero_policy = lambda pol=pol, **kw: pol
        shim_emctx.epistemic_qtransform_completed_by_mix_value = object()
        next_value = rng.standard_normal(B).astype(np.float32)
        zeros = jnp.zeros((B, A), jnp.float32)

        class _Forward:  # forward.apply(params, state, observation, is_training=False) -> ((exploit, explore, value, ube, rvar), state)
            calls = 0

            def apply(self, params, state, observation, is_training=False):
                type(self).calls += 1
                v = jnp.zeros(B, jnp.float32) if type(self).calls == 1 else jnp.array(next_value)
                return (zeros, zeros, v, jnp.zeros(B, jnp.float32), jnp.zeros(B, jnp.float32)), state

        legal = rng.random((B, A)) < 0.8
        legal[:, 0] |= ~legal.any(1)
        term = rng.random(B) < 0.2
        nterm = rng.random(B) < 0.3
        nrew = (rng.random(B) < 0.3).astype(np.float32).reshape(B, 1)
        first = types.SimpleNamespace(observation=jnp.zeros((B, 4), bool), legal_action_mask=jnp.array(legal), terminated=jnp.array(term))
        second = types.SimpleNamespace(observation=jnp.zeros((B, 4), bool), rewards=jnp.array(nrew), terminated=jnp.array(nterm))
        config = types.SimpleNamespace(reanalyze_beta=0.0, reanalyze_simulations_per_step=8, discount=gamma, exploration_ube_target=ube_expl,
                                       exploration_beta=ebeta, exploration_policy_target_temperature=temp)
        context = types.SimpleNamespace(forward=_Forward(), reanalyze_recurrent_fn=None)
        ro = rrean.reanalyze((None, None), config, context, types.SimpleNamespace(first=first, second=second), KEY)
        p = f"c{ci}_"
        out[p + "cfg"] = np.array([gamma, ebeta, float(ube_expl), temp], np.float64)
        for k, v in (("action", action), ("qvalues", summ.qvalues), ("qvar", summ.qvalues_epistemic_variance), ("visit_counts", visits),
                     ("value", summ.value), ("value_std", summ.value_epistemic_std), ("next_value", next_value), ("next_rewards", nrew[:, 0]),
                     ("next_terminated", nterm.astype(np.uint8)), ("terminated", term.astype(np.uint8)), ("invalid", (~legal).astype(np.uint8)),
                     ("value_target", np_(ro.value_target, np.float32)), ("ube_target", np_(ro.ube_target, np.float32)),
                     ("exploration_policy_target", np_(ro.exploration_policy_target, np.float32)),
                     ("exploitation_policy_target", np_(ro.exploitation_policy_target, np.float32))):
            out[p + k] = np.asarray(v)
    out["num_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(OUT, "reanalyze.npz"), **out)
    print("reanalyze.npz", len(out))


# ----------------------------------------------------------------------------- convolutional evaluators (resnet.py / minatar.py)
def _randomise(tree, rng, positive=()):
    """Fill every leaf of a haiku pytree with seeded values (weights ~ their init scale, BatchNorm statistics non-trivial)."""
    out = {}
    for mod, leaves in tree.items():
        out[mod] = {}
        for name, v0 in leaves.items():
            v = np.asarray(v0)
            if v.dtype == np.uint8 or name in ("hidden", "counter"):
                out[mod][name] = v0
            elif any(mod.endswith(p) for p in positive) and name == "average":  # variances
                out[mod][name] = jnp.array(rng.uniform(0.3, 2.0, v.shape).astype(np.float32))
            elif name == "scale":
                out[mod][name] = jnp.array(rng.uniform(0.5, 1.5, v.shape).astype(np.float32))
            elif name in ("offset", "b", "average"):
                out[mod][name] = jnp.array((rng.standard_normal(v.shape) * 0.2).astype(np.float32))
            else:  # conv / linear weights: haiku's TruncatedNormal(1 / sqrt(fan_in))
                fan_in = int(np.prod(v.shape[:-1]))
                out[mod][name] = jnp.array((rng.standard_normal(v.shape).clip(-2, 2) / np.sqrt(fan_in)).astype(np.float32))
    return out


def gen_convnet():
    from network.resnet import EpistemicResidualAZNet
    from network.minatar import EpistemicMinatarAZNet

    rng = np.random.default_rng(7)
    out = {}
    cases = [("resnet_v2", EpistemicResidualAZNet, dict(num_actions=65, resnet_v2=True), (8, 8, 2)),          # an othello-sized board
             ("resnet_v1", EpistemicResidualAZNet, dict(num_actions=10, resnet_v2=False, num_blocks=2), (6, 6, 4)),
             ("minatar", EpistemicMinatarAZNet, dict(num_actions=6, discount=0.99), (10, 10, 4))]              # MinAtar frame, 4 channels
    for tag, cls, kw, (H, W, Cc) in cases:
        def fwd(x, update_hash=False):
            net = cls(hash_class=XXHash, hash_args=dict(bits_per_hash=24), **kw)
            return net(x, is_training=False, test_local_stats=False, update_hash=update_hash)

        f = hk.without_apply_rng(hk.transform_with_state(fwd))
        B = 6
        obs = rng.random((B, H, W, Cc)) < 0.3
        params, state = f.init(None, jnp.array(obs))
        params = _randomise(params, rng)
        state = _randomise(state, rng, positive=("var_ema",))
        _, state = f.apply(params, state, jnp.array(obs[: B // 2]), update_hash=True)  # half of the batch becomes "seen"
        outs, _ = f.apply(params, state, jnp.array(obs))
        out[f"{tag}_obs"] = obs.astype(np.uint8)
        for name, v in zip(("exploit", "explore", "value", "ube", "novelty"), outs):
            out[f"{tag}_{name}"] = np_(v, np.float32)
        for mod, leaves in params.items():
            for name, v in leaves.items():
                out[f"{tag}_P|{mod}|{name}"] = np_(v, np.float32)
        for mod, leaves in state.items():
            for name, v in leaves.items():
                if name == "binary_set":
                    bs = np_(v)
                    out[f"{tag}_set_idx"] = np.flatnonzero(bs).astype(np.int64)
                    out[f"{tag}_set_val"] = bs[out[f"{tag}_set_idx"]]
                    out[f"{tag}_set_mod"] = np.array(mod)
                elif name == "average":
                    out[f"{tag}_S|{mod}|{name}"] = np_(v, np.float32)
    np.savez_compressed(os.path.join(OUT, "convnet.npz"), **out)
    print("convnet.npz", len(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "convnet":
        gen_convnet()
        sys.exit(0)
    gen_convnet()
    gen_reanalyze()
    gen_deepsea()
    gen_hash()
    gen_net()
    gen_subleq()
