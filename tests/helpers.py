"""Shared builders for the parity tests: seeded synthetic inputs on the host (numpy),
mirrored to the device for the CUDA path."""
import numpy as np

from e_alphazero_b200 import _abi
from oracle import oracle as O


def make_env(kind, seed=0, **kw):
    rng = np.random.default_rng(seed)
    if kind == "deepsea":
        N = kw.get("size", 10)
        return O.Env.deepsea(N, (rng.random((N, N)) < 0.5).astype(np.uint8))
    return O.Env.subleq(kw.get("word_size", 16), kw.get("binary", True), kw.get("reward_fn", 0))


def make_net(env, seed=0, hash_io=None, fill=0.0):
    if hash_io is None:
        hash_io = int(env.kind == _abi.ENV_SUBLEQ)
    net = O.FcNet.random(env.obs_dim, env.num_actions, seed=seed, hash_io=hash_io, word_size=env.word_size)
    if fill > 0:
        rng = np.random.default_rng(seed + 1)
        net.binary_set[:] = (rng.random(net.binary_set.size) < fill) * rng.integers(1, 256, net.binary_set.size)
    return net


def random_states(env, B, seed=0, max_steps=None):
    """States reached by random play from init (exercises every depth, terminal and solved states)."""
    rng = np.random.default_rng(seed)
    A = env.num_actions
    if env.kind == _abi.ENV_DEEPSEA:
        st = O.env_init(env, B)
        T = max_steps if max_steps is not None else env.size
    else:
        st = O.env_init(env, B, rng.integers(1, 4, B))
        T = max_steps if max_steps is not None else 10
    stop = rng.integers(0, T + 1, B)
    for t in range(T):
        act = rng.integers(0, A, B)
        if env.kind == _abi.ENV_SUBLEQ:  # bias towards IN/OUT addresses so programs do something
            ws = env.word_size
            act = np.where(rng.random(B) < 0.5, rng.integers(ws - 4, ws, B), act)
            if t == 0:
                act[: B // 8] = ws - 2
            if t == 1:
                act[: B // 8] = ws - 3
        nxt = O.env_step(env, st, act)
        adv = t < stop
        for k in st:
            st[k] = np.where(adv.reshape((-1,) + (1,) * (st[k].ndim - 1)), nxt[k], st[k])
    return st


def make_root(env, net, B, seed=0, beta_max=1.0, invalid_frac=0.0, states=None):
    rng = np.random.default_rng(seed + 7)
    A = env.num_actions
    st = states if states is not None else random_states(env, B, seed)
    ev = O.mlp_forward_states(net, env, st)
    root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"],
                beta=(beta_max * np.linspace(0, 1, B)).astype(np.float32), embedding=st,
                gumbel=rng.gumbel(size=(B, A)).astype(np.float32))
    if invalid_frac > 0:
        inv = rng.random((B, A)) < invalid_frac
        inv[:, 0] &= rng.random(B) < 0.5
        inv[0, :] = True  # one all-invalid row: masked_argmax must return 0
        root["invalid_actions"] = inv.astype(np.uint8)
    return root


def to_device(x):
    import torch

    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def device_env(env):
    from e_alphazero_b200 import ops

    if env.kind == _abi.ENV_DEEPSEA:
        return ops.deepsea_spec(env.size, env.action_map)
    return ops.subleq_spec(env.word_size, env.binary_encoding, env.reward_fn)


def device_net(net):
    from e_alphazero_b200 import ops

    return ops.FcParams.from_numpy(net.w, net.b, net.binary_set, net.num_actions, net.hash_bits, net.hash_io, net.word_size,
                                   net.max_u, net.novelty_scale)


def device_root(env, denv, root):
    from e_alphazero_b200 import ops

    d = {k: to_device(v) for k, v in root.items() if k != "embedding"}
    d["embedding"] = ops.state_to_device(denv, root["embedding"])
    return d


def assert_same_bits(a, b, name=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype.kind == "f":
        same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0))
    else:
        same = a == b
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ; first at {bad[0].tolist()}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}")


def uncompact(env, emb):
    """Compact in-tree states [M,S] (uint8) -> oracle state dict (numpy); inverse of the packing in DESIGN.md."""
    emb = np.ascontiguousarray(emb)
    M = emb.shape[0]
    st = O.alloc_state(env, M)
    if env.kind == _abi.ENV_DEEPSEA:
        v = emb.view(np.uint32).reshape(M)
        st["step_count"][:] = v & 0xFFF
        st["col"][:] = (v >> 12) & 0xFFF
        st["terminated"][:] = (v >> 24) & 1
        st["truncated"][:] = (v >> 25) & 1
        return st
    h = emb[:, :40].copy().view(np.uint16).reshape(M, 20)
    st["input_after"][:] = h[:, 0:8]
    st["output_after"][:] = h[:, 8:16]
    st["step_count"][:] = h[:, 16]
    st["task"][:] = emb[:, 34]
    st["terminated"][:] = emb[:, 35] & 1
    st["truncated"][:] = (emb[:, 35] >> 1) & 1
    st["solved"][:] = (emb[:, 35] >> 2) & 1
    st["memory"][:] = emb[:, 40:40 + env.word_size]
    return st


# --------------------------------------------------------------------------- convolutional evaluators
CONVNET_CASES = {"resnet_v2": dict(kind=_abi.CONVNET_RESNET, num_actions=65, resnet_v2=True, num_blocks=5),
                 "resnet_v1": dict(kind=_abi.CONVNET_RESNET, num_actions=10, resnet_v2=False, num_blocks=2),
                 "minatar": dict(kind=_abi.CONVNET_MINATAR, num_actions=6, discount=0.99)}


def load_convnet_golden(g, tag):
    """tests/golden/convnet.npz -> (description for _abi.fill_convnet_params, observation, expected outputs)."""
    params, state = {}, {}
    for key in g.files:
        if key.startswith(f"{tag}_P|") or key.startswith(f"{tag}_S|"):
            _, mod, name = key.split("|")
            (params if key.startswith(f"{tag}_P|") else state).setdefault(mod, {})[name] = g[key]
    bset = np.zeros(1 << 21, np.uint8)
    bset[g[f"{tag}_set_idx"]] = g[f"{tag}_set_val"]
    state[str(g[f"{tag}_set_mod"])] = {"binary_set": bset}
    obs = g[f"{tag}_obs"]
    _, H, W, Cc = obs.shape
    kw = dict(CONVNET_CASES[tag])
    desc = _abi.convnet_description(params, state, kw.pop("kind"), H, W, Cc, kw.pop("num_actions"), **kw)
    exp = {k: g[f"{tag}_{k}"] for k in ("exploit", "explore", "value", "ube", "novelty")}
    return desc, obs, exp


def random_convnet(kind, H, W, Cc, A, seed=0, num_channels=None, hidden=64, num_blocks=5, resnet_v2=True, fill=0.5):
    """A seeded random network description (numpy leaves) without going through haiku pytrees."""
    rng = np.random.default_rng(seed)
    C_ = num_channels or (64 if kind == _abi.CONVNET_RESNET else 16)

    def cv(*shape):
        fan_in = int(np.prod(shape[:-1]))
        return dict(w=(rng.standard_normal(shape).clip(-2, 2) / np.sqrt(fan_in)).astype(np.float32), b=(rng.standard_normal(shape[-1]) * 0.2).astype(np.float32))

    def bn(c):
        return dict(scale=rng.uniform(0.5, 1.5, c).astype(np.float32), offset=(rng.standard_normal(c) * 0.2).astype(np.float32),
                    mean=(rng.standard_normal(c) * 0.2).astype(np.float32), var=rng.uniform(0.3, 2.0, c).astype(np.float32))

    bset = ((rng.random(1 << 21) < fill) * rng.integers(1, 256, 1 << 21)).astype(np.uint8)
    d = dict(kind=kind, height=H, width=W, in_channels=Cc, num_actions=A, num_channels=C_, hidden=C_ if kind == _abi.CONVNET_RESNET else hidden,
             num_blocks=num_blocks if kind == _abi.CONVNET_RESNET else 0, resnet_v2=int(resnet_v2), binary_set=bset, hash_bits=24, max_u=1.0,
             novelty_scale=1.0, local_unc_scale=1.0 / (1 - 0.99 ** 2))
    if kind == _abi.CONVNET_RESNET:
        d["stem"] = cv(3, 3, Cc, C_)
        d["stem_bn"], d["final_bn"] = bn(C_), bn(C_)
        d["blocks"] = [dict(bn=[bn(C_), bn(C_)], conv=[cv(3, 3, C_, C_), cv(3, 3, C_, C_)]) for _ in range(num_blocks)]
        d["heads"] = []
        for h in range(4):
            k = 2 if h < 2 else 1
            hd = dict(conv=cv(C_, k), bn=bn(k), fc=cv(H * W * k, A if h < 2 else C_))
            if h >= 2:
                hd["out"] = cv(C_, 1)
            d["heads"].append(hd)
    else:
        d["towers"] = [dict(conv=cv(3, 3, Cc, C_), fc=[cv(H * W * C_, hidden), cv(hidden, hidden)]) for _ in range(2)]
        d["mheads"] = [[cv(hidden, hidden), cv(hidden, A if h in (0, 2) else 1)] for h in range(4)]
    return d
