"""Timeline of one mlp_tensor_kernel CTA (clock64 per pipeline chunk) -- a measurement aid for the tcgen05 pipeline.
Usage (on a B200): python profiles/trace_mlp_tensor.py [c2|c3]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from e_alphazero_b200 import _abi, _lib, ops
from e_alphazero_b200.selfplay import SelfplayRunner

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, kw, B, n, gamma, desc = bench.WORKLOADS[wl]
envp, netp = bench.synth_params(kind, kw, 0)
env = ops.deepsea_spec(envp["size"], envp["action_map"]) if kind == "deepsea" else ops.subleq_spec(envp["word_size"], True)
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"])
runner = SelfplayRunner(env, net, B, 8, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=1)
states = ops.env_init(env, B, task_ids=torch.ones(B, dtype=torch.int32, device="cuda") if kind == "subleq" else None)
for _ in range(2):
    states, _ = runner.step(states)
torch.cuda.synchronize()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.eaz_debug_set_mlp_trace(C.c_void_p(buf.data_ptr()))
states, _ = runner.step(states)
torch.cuda.synchronize()
lib.eaz_debug_set_mlp_trace(None)
t = buf.cpu().numpy()
t0 = t[0]
print(f"state staged at {t[6] - t0}, observation bits built at {t[7] - t0}, after barrier {t[776] - t0}, first load_raw {t[777] - t0}, before first store {t[778] - t0}")
print(f"workload {wl}: kernel span {t[1] - t0} cycles; after PDL wait {t[2] - t0}; layer-3 start {t[3] - t0}, layer-3 fma done {t[4] - t0}, after barrier {t[5] - t0}")
ready = [int(x - t0) for x in t[8:8 + 256] if x]
arrive = [int(x - t0) for x in t[264:264 + 256] if x]
issued = [int(x - t0) for x in t[520:520 + 256] if x]
print("chunk: producer-arrived  mma-ready  mma-issued   (cycles since kernel start)")
for i, (a, r, s) in enumerate(zip(arrive, ready, issued)):
    print(f"{i:3d}: {a:8d} {r:8d} {s:8d}   d_ready={r - (ready[i - 1] if i else 0):6d}")
