"""Section breakdown of tree_step_kernel (clock64 stamps per tree per simulation) -- a measurement aid.
Usage (on a B200): python profiles/trace_tree_step.py [c2|c3|c4]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from e_alphazero_b200 import _lib, ops
from e_alphazero_b200.selfplay import SelfplayRunner

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, kw, B, n, gamma, desc = bench.WORKLOADS[wl]
envp, netp = bench.synth_params(kind, kw, 0)
env = ops.deepsea_spec(envp["size"], envp["action_map"]) if kind == "deepsea" else ops.subleq_spec(envp["word_size"], True)
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"])
runner = SelfplayRunner(env, net, B, n, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=1, use_graph=False, fused_root=True)
states = ops.env_init(env, B, task_ids=torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None)
for _ in range(4):
    states, _ = runner.step(states)
torch.cuda.synchronize()
buf = torch.zeros((n + 1) * B * 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.eaz_debug_set_tree_trace(C.c_void_p(buf.data_ptr()))
states, _ = runner.step(states)
torch.cuda.synchronize()
lib.eaz_debug_set_tree_trace(None)
t = buf.cpu().numpy().reshape(n + 1, B, 8)
names = ["expand", "backward", "refresh", "chase"]
print(f"workload {wl}: per-tree cycles by section (sims 1..{n - 1}); L = backward path length, depth = next descent length")
mid = t[1:n]
sec = np.stack([mid[..., k + 1] - mid[..., k] for k in range(4)], -1)  # [sim, B, 4]
tot = mid[..., 4] - mid[..., 0]
print("mean      :", {nm: int(sec[..., k].mean()) for k, nm in enumerate(names)}, "total", int(tot.mean()), "L", float(mid[..., 6].mean()), "depth", float(mid[..., 5].mean()))
sl = tot.argmax(1)  # slowest tree per simulation
idx = np.arange(mid.shape[0])
print("slowest   :", {nm: int(sec[idx, sl, k].mean()) for k, nm in enumerate(names)}, "total", int(tot[idx, sl].mean()), "L", float(mid[idx, sl, 6].mean()),
      "depth", float(mid[idx, sl, 5].mean()))
span = mid[..., 4].max(1) - mid[..., 0].min(1)
print(f"kernel span first-start -> last-chase-end: mean {span.mean():.0f} cycles; start skew {np.mean(mid[..., 0].max(1) - mid[..., 0].min(1)):.0f}")
for s in (1, 8, 32, n - 1):
    m = t[s]
    print(f"sim {s:3d}: mean sections {[int((m[:, k + 1] - m[:, k]).mean()) for k in range(4)]} max total {int((m[:, 4] - m[:, 0]).max())} mean L {m[:, 6].mean():.1f}")
mx = (t[1:n, :, 4] - t[1:n, :, 0]).max(1)
print(f"slowest tree per launch: mean {mx.mean():.0f} cycles, median {np.median(mx):.0f}, p90 {np.percentile(mx, 90):.0f}, max {mx.max():.0f}; "
      f"share of trees on the DIRECT path {1.0 - (t[1:n, :, 7] > 0).mean():.4f}; max L {int(t[1:n, :, 6].max())}")
