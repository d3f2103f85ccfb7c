"""numpy/ctypes binding of the CPU oracle (oracle/eaz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing in e_alphazero_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

from e_alphazero_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libeaz_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "eaz_oracle.c")
    deps = [src, os.path.join(_HERE, "..", "include", "eaz_b200.h"), os.path.join(_HERE, "..", "include", "eaz_math.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()  # (a no-op when libeaz_oracle.so is newer than its three sources)
        _lib = C.CDLL(_SO)
        _lib.orc_expf.restype = C.c_float
        _lib.orc_expf.argtypes = [C.c_float]
        _lib.orc_tanhf.restype = C.c_float
        _lib.orc_tanhf.argtypes = [C.c_float]
        _lib.orc_logf.restype = C.c_float
        _lib.orc_logf.argtypes = [C.c_float]
        _lib.orc_tree_sum_probe.restype = C.c_float
    return _lib


def set_threads(n: int) -> int:
    return lib().orc_set_threads(int(n))


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- envs
@dataclass
class Env:
    kind: int
    size: int = 0
    action_map: np.ndarray | None = None
    word_size: int = 0
    binary_encoding: int = 1
    reward_fn: int = _abi.SUBLEQ_REWARD_SOLVED

    @staticmethod
    def deepsea(size, action_map=None):
        am = None if action_map is None else np.ascontiguousarray(action_map, dtype=np.uint8).reshape(size, size)
        return Env(_abi.ENV_DEEPSEA, size=size, action_map=am)

    @staticmethod
    def subleq(word_size, binary_encoding=True, reward_fn=_abi.SUBLEQ_REWARD_SOLVED):
        return Env(_abi.ENV_SUBLEQ, word_size=word_size, binary_encoding=int(binary_encoding), reward_fn=reward_fn)

    def struct(self) -> _abi.EazEnv:
        return _abi.EazEnv(self.kind, self.size, _ptr(self.action_map), self.word_size, self.binary_encoding, self.reward_fn)

    @property
    def num_actions(self):
        s = self.struct()
        return lib().orc_env_num_actions(C.byref(s))

    @property
    def obs_dim(self):
        s = self.struct()
        return lib().orc_env_obs_dim(C.byref(s))

    @property
    def obs_cols(self):
        s = self.struct()
        return lib().orc_env_obs_cols(C.byref(s))

    def hash_dim(self, hash_io):
        s = self.struct()
        return lib().orc_env_hash_dim(C.byref(s), int(hash_io))

    @property
    def compact_bytes(self):
        s = self.struct()
        return lib().orc_env_compact_bytes(C.byref(s))


STATE_FIELDS = {
    # name: (dtype, trailing-shape fn(env))
    "step_count": (np.int32, lambda e: ()),
    "rewards": (np.float32, lambda e: (1,)),
    "terminated": (np.uint8, lambda e: ()),
    "truncated": (np.uint8, lambda e: ()),
    "col": (np.int32, lambda e: ()),
    "memory": (np.int32, lambda e: (e.word_size,)),
    "task": (np.int32, lambda e: ()),
    "solved": (np.uint8, lambda e: ()),
    "input_after": (np.int32, lambda e: (8,)),
    "output_after": (np.int32, lambda e: (8,)),
}
DEEPSEA_FIELDS = ["step_count", "rewards", "terminated", "truncated", "col"]
SUBLEQ_FIELDS = ["step_count", "rewards", "terminated", "truncated", "memory", "task", "solved", "input_after", "output_after"]


def state_fields(env) -> list[str]:
    return DEEPSEA_FIELDS if env.kind == _abi.ENV_DEEPSEA else SUBLEQ_FIELDS


def alloc_state(env: Env, B: int, with_obs: bool = False) -> dict:
    st = {}
    for name in state_fields(env):
        dt, shp = STATE_FIELDS[name]
        st[name] = np.zeros((B,) + shp(env), dtype=dt)
    if with_obs:
        st["observation"] = np.zeros((B, env.obs_dim), dtype=np.uint8)
    return st


def state_struct(st: dict) -> _abi.EazState:
    s = _abi.EazState()
    for name, _ in _abi.EazState._fields_:
        setattr(s, name, _ptr(st.get(name)))
    return s


def copy_state(st: dict) -> dict:
    return {k: v.copy() for k, v in st.items()}


def _chk(rc, what):
    if rc < 0:
        raise ValueError(f"oracle {what} failed with {rc}")
    return rc


def env_init(env: Env, B: int, task_ids=None, with_obs=False) -> dict:
    st = alloc_state(env, B, with_obs)
    e, s = env.struct(), state_struct(st)
    t = None if task_ids is None else np.ascontiguousarray(task_ids, dtype=np.int32)
    _chk(lib().orc_env_init(C.byref(e), _ptr(t), C.byref(s), B), "env_init")
    return st


def env_step(env: Env, st: dict, action, auto_reset=False, task_ids=None) -> dict:
    st = copy_state(st)
    B = st["step_count"].shape[0]
    a = np.ascontiguousarray(action, dtype=np.int32)
    t = None if task_ids is None else np.ascontiguousarray(task_ids, dtype=np.int32)
    e, s = env.struct(), state_struct(st)
    _chk(lib().orc_env_step(C.byref(e), C.byref(s), _ptr(a), int(auto_reset), _ptr(t), B), "env_step")
    return st


def env_observe(env: Env, st: dict) -> np.ndarray:
    B = st["step_count"].shape[0]
    obs = np.zeros((B, env.obs_dim), dtype=np.uint8)
    e, s = env.struct(), state_struct(st)
    _chk(lib().orc_env_observe(C.byref(e), C.byref(s), _ptr(obs), B), "env_observe")
    return obs


def env_compact(env: Env, st: dict) -> np.ndarray:
    B = st["step_count"].shape[0]
    out = np.zeros((B, env.compact_bytes), dtype=np.uint8)
    e, s = env.struct(), state_struct(st)
    _chk(lib().orc_env_compact(C.byref(e), C.byref(s), _ptr(out), B), "env_compact")
    return out


def subleq_test_cases(task: int, ws: int):
    i = np.zeros((3, 8), np.int32)
    o = np.zeros((3, 8), np.int32)
    _chk(lib().orc_subleq_test_cases(int(task), int(ws), _ptr(i), _ptr(o)), "test_cases")
    return i, o


def subleq_simulate(ws: int, memory, test_in, test_out):
    m = np.ascontiguousarray(memory, np.int32)
    ti = np.ascontiguousarray(test_in, np.int32)
    to = np.ascontiguousarray(test_out, np.int32)
    ia, oa, bcc = np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(3, np.int32)
    lib().orc_subleq_simulate(int(ws), _ptr(m), _ptr(ti), _ptr(to), _ptr(ia), _ptr(oa), _ptr(bcc))
    return dict(input_after=ia, output_after=oa, bytes_used=int(bcc[0]), cycles_used=int(bcc[1]), correct=bool(bcc[2]))


def seq_halving_table(max_considered: int, n: int) -> np.ndarray:
    t = np.zeros((max_considered + 1, n), np.int32)
    _chk(lib().orc_seq_halving_table(int(max_considered), int(n), _ptr(t)), "seq_halving_table")
    return t


# --------------------------------------------------------------------------- hash
def xxhash_indices(x, bits=24) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    x = x.reshape(x.shape[0], -1)
    out = np.zeros(x.shape[0], np.uint32)
    _chk(lib().orc_xxhash_indices(_ptr(x), x.shape[0], x.shape[1], int(bits), _ptr(out)), "xxhash_indices")
    return out


def hash_lookup(x, binary_set, bits=24) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    x = x.reshape(x.shape[0], -1)
    out = np.zeros(x.shape[0], np.uint8)
    _chk(lib().orc_hash_lookup(_ptr(x), x.shape[0], x.shape[1], int(bits), _ptr(binary_set), _ptr(out)), "hash_lookup")
    return out


def hash_update(x, binary_set, bits=24) -> None:
    x = np.ascontiguousarray(x, np.float32)
    x = x.reshape(x.shape[0], -1)
    _chk(lib().orc_hash_update(_ptr(x), x.shape[0], x.shape[1], int(bits), _ptr(binary_set)), "hash_update")


# --------------------------------------------------------------------------- network
@dataclass
class FcNet:
    """haiku params of EpistemicFullyConnectedAZNet as numpy arrays: w[h][l] is [in,out]."""

    in_dim: int
    hidden: int
    num_actions: int
    w: list
    b: list
    binary_set: np.ndarray
    hash_bits: int = 24
    hash_io: int = 0
    word_size: int = 0
    max_u: float = 1.0
    novelty_scale: float = 1.0
    _keep: list = field(default_factory=list)

    @staticmethod
    def random(in_dim, num_actions, hidden=256, seed=0, hash_bits=24, hash_io=0, bias_scale=0.05, word_size=0):
        """haiku default init (trunc-normal, std 1/sqrt(fan_in)); biases get a small
        non-zero value so that the parity tests exercise them."""
        rng = np.random.default_rng(seed)
        w, b = [], []
        for h in range(4):
            outs = [hidden, hidden, 1 if h < 2 else num_actions]
            ins = [in_dim, hidden, hidden]
            ws, bs = [], []
            for i, o in zip(ins, outs):
                x = rng.standard_normal((i, o)).clip(-2, 2) / np.sqrt(i)
                ws.append(np.ascontiguousarray(x, np.float32))
                bs.append(np.ascontiguousarray(rng.standard_normal(o) * bias_scale, np.float32))
            w.append(ws)
            b.append(bs)
        return FcNet(in_dim, hidden, num_actions, w, b, np.zeros(1 << (hash_bits - 3), np.uint8), hash_bits, hash_io, word_size)

    def struct(self) -> _abi.EazFcParams:
        s = _abi.EazFcParams()
        s.in_dim, s.hidden, s.num_actions = self.in_dim, self.hidden, self.num_actions
        for h in range(4):
            for l in range(3):
                s.w[h][l] = _ptr(self.w[h][l])
                s.b[h][l] = _ptr(self.b[h][l])
        s.binary_set = _ptr(self.binary_set)
        s.hash_bits, s.hash_io, s.word_size = self.hash_bits, self.hash_io, self.word_size
        s.max_u, s.novelty_scale = self.max_u, self.novelty_scale
        return s


def mlp_forward(net: FcNet, obs, hash_dim=None) -> dict:
    obs = np.ascontiguousarray(obs, np.uint8).reshape(len(obs), -1)
    B, A = obs.shape[0], net.num_actions
    out = dict(exploit_logits=np.zeros((B, A), np.float32), explore_logits=np.zeros((B, A), np.float32),
               value=np.zeros(B, np.float32), ube=np.zeros(B, np.float32), novelty=np.zeros(B, np.float32))
    s = net.struct()
    hd = obs.shape[1] if hash_dim is None else hash_dim
    _chk(lib().orc_mlp_forward(C.byref(s), _ptr(obs), B, int(hd), _ptr(out["exploit_logits"]), _ptr(out["explore_logits"]),
                               _ptr(out["value"]), _ptr(out["ube"]), _ptr(out["novelty"])), "mlp_forward")
    return out


def mlp_forward_states(net: FcNet, env: Env, st: dict) -> dict:
    return mlp_forward(net, env_observe(env, st), env.hash_dim(net.hash_io))


# --------------------------------------------------------------------------- search
class Replay(C.Structure):
    _fields_ = [("states", C.c_void_p), ("logits", C.c_void_p), ("value", C.c_void_p), ("var", C.c_void_p), ("S", C.c_int32)]


def alloc_search_outputs(B, N, A, S, want_tree=True) -> dict:
    out = {}
    dts = {"i32": np.int32, "f32": np.float32, "u8": np.uint8}
    shp = {"B": (B,), "BA": (B, A), "BN": (B, N), "BNA": (B, N, A), "BNS": (B, N, S)}
    fields = _abi.SUMMARY_FIELDS + (_abi.TREE_FIELDS if want_tree else [])
    for name, dt, kind in fields:
        out[name] = np.zeros(shp[kind], dts[dt])
    return out


def search(cfg: _abi.EazSearchConfig, env: Env, net: FcNet | None, root: dict, want_tree=True, replay: dict | None = None) -> dict:
    """root: prior_logits [B,A], value [B], value_epistemic_variance [B], beta [B], embedding (state dict),
    gumbel [B,A], optional invalid_actions [B,A].  replay: dict(states, logits, value, var) from another tree."""
    B, A = root["prior_logits"].shape
    cfg.batch = B
    N = cfg.num_simulations + 1
    out = alloc_search_outputs(B, N, A, env.compact_bytes, want_tree)
    f32 = lambda k: np.ascontiguousarray(root[k], np.float32)
    keep = dict(prior_logits=f32("prior_logits"), value=f32("value"), var=f32("value_epistemic_variance"),
                beta=f32("beta"), gumbel=f32("gumbel"))
    inv = root.get("invalid_actions")
    if inv is not None:
        inv = np.ascontiguousarray(inv, np.uint8)
    e, s = env.struct(), state_struct(root["embedding"])
    n = net.struct() if net is not None else None
    inp = _abi.EazSearchInputs(_ptr(keep["prior_logits"]), _ptr(keep["value"]), _ptr(keep["var"]), _ptr(keep["beta"]),
                               C.pointer(s), _ptr(inv), _ptr(keep["gumbel"]), C.pointer(e),
                               C.pointer(n) if n is not None else None)
    o = _abi.EazSearchOutputs()
    for name, _, _ in _abi.SEARCH_OUTPUT_FIELDS:
        setattr(o, name, _ptr(out.get(name)))
    rp = None
    if replay is not None:
        rk = dict(states=np.ascontiguousarray(replay["states"], np.uint8), logits=np.ascontiguousarray(replay["logits"], np.float32),
                  value=np.ascontiguousarray(replay["value"], np.float32), var=np.ascontiguousarray(replay["var"], np.float32))
        rp = Replay(_ptr(rk["states"]), _ptr(rk["logits"]), _ptr(rk["value"]), _ptr(rk["var"]), env.compact_bytes)
    rc = lib().orc_search_gumbel(C.byref(cfg), C.byref(inp), C.byref(o), C.byref(rp) if rp is not None else None)
    _chk(rc, "search")
    out["replay_misses"] = rc
    return out


def reanalyze_targets(discount, exploration_beta, exploration_ube_target, temperature, action, qvalues, qvar, visit_counts, value, value_std,
                      next_state_value, next_rewards, next_terminated, terminated, invalid_actions=None) -> dict:
    """reanalyze.py:86-129 on host arrays; returns value_target [B], ube_target [B], exploration_policy_target [B,A]."""
    f = lambda a: np.ascontiguousarray(a, np.float32)
    u = lambda a: np.ascontiguousarray(a, np.uint8)
    q = f(qvalues)
    B, A = q.shape
    cfg = _abi.EazReanalyzeConfig(float(discount), float(exploration_beta), int(bool(exploration_ube_target)), float(temperature))
    keep = [np.ascontiguousarray(action, np.int32), q, f(qvar), f(visit_counts), f(value), f(value_std), f(next_state_value), f(next_rewards),
            u(next_terminated), u(terminated), u(invalid_actions) if invalid_actions is not None else None]
    out = dict(value_target=np.zeros(B, np.float32), ube_target=np.zeros(B, np.float32), exploration_policy_target=np.zeros((B, A), np.float32))
    rc = lib().orc_reanalyze_targets(C.byref(cfg), B, A, *[_ptr(k) for k in keep], _ptr(out["value_target"]), _ptr(out["ube_target"]),
                                     _ptr(out["exploration_policy_target"]))
    _chk(rc, "reanalyze_targets")
    return out


def expf(x: float) -> float:
    return lib().orc_expf(float(x))


def logf(x: float) -> float:
    return float(lib().orc_logf(C.c_float(x)))


def tanhf(x: float) -> float:
    return lib().orc_tanhf(float(x))


def tree_sum(x) -> float:
    x = np.ascontiguousarray(x, np.float32)
    return lib().orc_tree_sum_probe(_ptr(x), x.size)


def softmax(x) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    p = np.zeros_like(x)
    lib().orc_softmax_probe(_ptr(x), x.size, _ptr(p))
    return p


# --------------------------------------------------------------------------- convolutional evaluators (resnet.py / minatar.py)
def convnet_forward(desc: dict, observation) -> dict:
    """desc: _abi.convnet_description(...) with numpy leaves; observation bool [B,H,W,C]."""
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a, dtype=np.uint8 if np.asarray(a).dtype == np.uint8 else np.float32)
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    s = _abi.fill_convnet_params(desc, ptr)
    obs = np.ascontiguousarray(observation, np.uint8).reshape(len(observation), -1)
    B, A = obs.shape[0], desc["num_actions"]
    out = dict(exploit_logits=np.zeros((B, A), np.float32), explore_logits=np.zeros((B, A), np.float32), value=np.zeros(B, np.float32),
               ube=np.zeros(B, np.float32), novelty=np.zeros(B, np.float32))
    _chk(lib().orc_convnet_forward(C.byref(s), _ptr(obs), B, _ptr(out["exploit_logits"]), _ptr(out["explore_logits"]), _ptr(out["value"]),
                                   _ptr(out["ube"]), _ptr(out["novelty"])), "convnet_forward")
    return out

