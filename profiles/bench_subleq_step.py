"""Micro-benchmark of the Subleq transition kernel (eaz_env_step, csrc/env.cu: the same interpreter the in-tree step uses) on program
populations of different character -- a measurement aid.  Usage (on a B200): [EAZ_LIB_PATH=other.so] python profiles/bench_subleq_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from e_alphazero_b200 import ops
from oracle import oracle as O
from tests import helpers as H

ws, B = 16, 8192
env = H.make_env("subleq", word_size=ws)
denv = H.device_env(env)
rng = np.random.default_rng(0)


def population(kind):
    st = O.env_init(env, B, rng.integers(1, 4, B).astype(np.int32))
    if kind == "search-like":
        st = H.random_states(env, B, seed=1, max_steps=10)
        st["terminated"][:] = 0
        st["solved"][:] = 0
    else:
        prog = {"empty": [], "self-loop (period 1)": [3, 3, 0], "long period (48)": [7, 4, 1], "honest 200": [12, 15, 2, 7, 15, 0, 15, 1],
                "one honest 200 among empty": []}[kind]
        st["memory"][:, : len(prog)] = prog
        st["step_count"][:] = max(len(prog), 1)
        if kind == "one honest 200 among empty":
            st["memory"][17, :8] = [12, 15, 2, 7, 15, 0, 15, 1]
            st["step_count"][17] = 8
    st["step_count"][:] = np.minimum(st["step_count"], ws - 5)
    return st


for kind in ("empty", "self-loop (period 1)", "long period (48)", "honest 200", "one honest 200 among empty", "search-like"):
    st = population(kind)
    act = np.zeros(B, np.int32)
    dst0 = ops.state_to_device(denv, st)
    dact = H.to_device(act)
    ts = []
    for it in range(30):
        d = {k: v.clone() for k, v in dst0.items()}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.env_step_(denv, d, dact)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    exp = O.env_step(env, st, act)
    same = all((d[k].cpu().numpy().reshape(exp[k].shape) == exp[k]).all() for k in ("memory", "solved", "input_after", "output_after", "terminated"))
    print(f"{kind:30s} median {np.median(ts[5:]):7.1f} us  min {np.min(ts[5:]):7.1f} us   parity {'ok' if same else 'MISMATCH'}")
