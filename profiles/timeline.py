"""Launch timeline of one self-play step (globaltimer stamps written by block 0 of the per-simulation kernels):
shows kernel busy time vs the gaps between dependent launches.  Usage (on a B200): python profiles/timeline.py [c2|c3] [graph|eager]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from e_alphazero_b200 import _lib, ops
from e_alphazero_b200.selfplay import SelfplayRunner

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "graph"
kind, kw, B, n, gamma, desc = bench.WORKLOADS[wl]
envp, netp = bench.synth_params(kind, kw, 0)
env = ops.deepsea_spec(envp["size"], envp["action_map"]) if kind == "deepsea" else ops.subleq_spec(envp["word_size"], True)
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"])
lib = _lib.load()
buf = torch.zeros(8 + 4 * 2000, dtype=torch.int64, device="cuda")
lib.eaz_debug_set_timeline(C.c_void_p(buf.data_ptr()))  # set before graph capture so the pointer is baked into the graph
runner = SelfplayRunner(env, net, B, n, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=1, use_graph=(mode == "graph"), fused_root=True)
states = ops.env_init(env, B, task_ids=torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None)
for _ in range(4):
    states, _ = runner.step(states)
torch.cuda.synchronize()
buf.zero_()
torch.cuda.synchronize()
states, _ = runner.step(states)
torch.cuda.synchronize()
lib.eaz_debug_set_timeline(None)
t = buf.cpu().numpy()
cnt = int(t[0])
rec = sorted((int(t[8 + 4 * i]), int(t[9 + 4 * i]), int(t[10 + 4 * i]), int(t[11 + 4 * i])) for i in range(min(cnt, 2000)))
t0 = rec[0][0]
busy = {0: 0, 1: 0}
print(f"{wl} {mode}: {cnt} launches, span {(rec[-1][2] - t0) / 1e3:.1f} us")
prev_exit = None
gaps = []
for i, (a, w, e, k) in enumerate(rec):
    busy[k] += e - w
    if prev_exit is not None:
        gaps.append(w - prev_exit)
    prev_exit = e
    if i < 12:
        print(f"  {['tree', 'mlp '][k]} entry {(a - t0) / 1e3:8.2f}  start {(w - t0) / 1e3:8.2f}  exit {(e - t0) / 1e3:8.2f}  run {(e - w) / 1e3:6.2f} us")
n_tree = sum(1 for r in rec if r[3] == 0)
n_mlp = cnt - n_tree
print(f"block-0 run time: tree {busy[0] / 1e3 / max(n_tree, 1):.2f} us x {n_tree}, mlp {busy[1] / 1e3 / max(n_mlp, 1):.2f} us x {n_mlp}; "
      f"mean exit->next-start gap {sum(gaps) / len(gaps) / 1e3:.2f} us")
