"""Timeline of one mlp_gather_kernel CTA (clock64 milestones) -- a measurement aid.  Usage (on a B200): python profiles/trace_mlp_gather.py [c2|c4]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from e_alphazero_b200 import _lib, ops
from e_alphazero_b200.selfplay import SelfplayRunner

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, kw, B, n, gamma, desc = bench.WORKLOADS[wl]
envp, netp = bench.synth_params(kind, kw, 0)
env = ops.deepsea_spec(envp["size"], envp["action_map"])
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"])
runner = SelfplayRunner(env, net, B, 8, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=1, fused_root=True)
states = ops.env_init(env, B)
for _ in range(2):
    states, _ = runner.step(states)
torch.cuda.synchronize()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.eaz_debug_set_gather_trace(C.c_void_p(buf.data_ptr()))
states, _ = runner.step(states)
torch.cuda.synchronize()
lib.eaz_debug_set_gather_trace(None)
t = buf.cpu().numpy()
t0 = t[0]
names = ["entry", "prologue done", "after PDL wait", "cells loaded", "-", "all A published", "acc done", "layer 3 done", "-", "-", "end"]
for i, nm in enumerate(names):
    print(f"{nm:22s} {int(t[i] - t0):8d}")
print("chunk: B ready, A ready (cycles since entry)")
for c in range(8):
    print(f"  {c}: {int(t[16 + c] - t0):8d} {int(t[32 + c] - t0):8d}")
