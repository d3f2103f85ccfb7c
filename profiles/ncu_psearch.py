"""One self-play step at the given workload (for ncu captures of the persistent search kernel).  Usage: python profiles/ncu_psearch.py [c2|c4] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from e_alphazero_b200 import ops
from e_alphazero_b200.selfplay import SelfplayRunner

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind, kw, B, n, gamma, desc = bench.WORKLOADS[wl]
envp, netp = bench.synth_params(kind, kw, 0)
env = ops.deepsea_spec(envp["size"], envp["action_map"]) if kind == "deepsea" else ops.subleq_spec(envp["word_size"], True)
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"])
runner = SelfplayRunner(env, net, B, n, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=1, fused_root=True)
states = ops.env_init(env, B, task_ids=torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None)
for _ in range(steps):
    states, out = runner.step(states)
torch.cuda.synchronize()
print("ok", float(out.root_value.sum()))
