"""BASELINE config C5: Subleq + reanalyze sweep -- 16k..256k envs x 32/64/128/256 simulations at 1/2/4/8 GPUs, next to the CPU
restatement on the host cores.  One "call" = one reanalyze() (reanalyze.py:52-131): root forward on stored states -> E-MCTS
search -> next-state forward -> targets.  Envs are sharded over the ranks (no collective on the data path); the time of a
point is the max over ranks of the CUDA-event time.  Writes one JSON line per point to stdout (rank 0).

    python profiles/sweep_c5.py [--quick] [--reps 3]                                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/sweep_c5.py   # N GPUs
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from e_alphazero_b200 import _abi, ops
from e_alphazero_b200.reanalyze import ReanalyzeRunner

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--streams", type=int, default=3, help="EAZ_FLAG_STREAMS sub-batches (measured best for Subleq: 3)")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(dev))
ws, gamma = 16, 0.97
envp, netp = bench.synth_params("subleq", dict(word_size=ws), 0)
env = ops.subleq_spec(ws, True)
net = ops.FcParams.from_numpy(netp["w"], netp["b"], netp["binary_set"], netp["num_actions"], 24, netp["hash_io"], netp["word_size"], device=dev)
Bs = [16384, 65536] if args.quick else [16384, 32768, 65536, 131072, 262144]
ns = [32, 128] if args.quick else [32, 64, 128, 256]
gen = torch.Generator(device=dev).manual_seed(5 + rank)


def stored_transitions(B):
    """Replay-buffer stand-in: states reached by random programs of length U{0..12} and their successors."""
    st = ops.env_init(env, B, task_ids=torch.ones(B, dtype=torch.int32, device=dev), device=dev)
    length = torch.randint(0, 13, (B,), device=dev, generator=gen)
    for t in range(12):
        act = torch.randint(0, ws, (B,), device=dev, generator=gen, dtype=torch.int32)
        nxt = ops.env_step(env, st, act)
        keep = (length > t)
        for k in st:
            m = keep.reshape((-1,) + (1,) * (st[k].dim() - 1))
            st[k] = torch.where(m, nxt[k], st[k]).contiguous()
    act = torch.randint(0, ws, (B,), device=dev, generator=gen, dtype=torch.int32)
    return st, ops.env_step(env, st, act)


cpu = {}
if rank == 0 and not args.no_cpu:  # CPU restatement (oracle, OpenMP, all host cores) on a bounded sample per simulation count
    from oracle import oracle as O

    O.build()
    cores = O.set_threads(os.cpu_count() or 1)
    oenv = O.Env.subleq(ws, True)
    onet = O.FcNet(netp["in_dim"], 256, netp["num_actions"], netp["w"], netp["b"], netp["binary_set"], 24, netp["hash_io"], netp["word_size"])
    rng = np.random.default_rng(0)
    for n in ns:
        Bc = max(cores * 2, 32)
        st = O.env_init(oenv, Bc, np.ones(Bc, np.int32))
        for _ in range(4):
            st = O.env_step(oenv, st, rng.integers(0, ws, Bc).astype(np.int32))
        t0 = time.perf_counter()
        ev = O.mlp_forward_states(onet, oenv, st)
        root = dict(prior_logits=ev["exploit_logits"], value=ev["value"], value_epistemic_variance=ev["ube"], beta=np.zeros(Bc, np.float32), embedding=st,
                    gumbel=rng.gumbel(size=(Bc, ws)).astype(np.float32))
        O.search(_abi.default_search_config(num_simulations=n, discount=gamma), oenv, onet, root, want_tree=False)
        cpu[n] = dict(envs_per_s=Bc / (time.perf_counter() - t0), cores=cores, sample=f"{Bc} envs x 1 call")

for B in Bs:
    Bl = B // world
    first, second = stored_transitions(Bl)
    for n in ns:
        if (n + 1) * Bl * ws >= 2 ** 31:
            continue
        try:
            r = ReanalyzeRunner(env, net, Bl, n, gamma, reanalyze_beta=0.0, exploration_beta=0.0, mlp_mode=_abi.MLP_TENSOR, device=dev, seed=rank, use_graph=True, streams=args.streams)
            for _ in range(2):
                r(first, second)
            torch.cuda.synchronize()
            ms = []
            for i in range(args.reps):
                if world > 1:
                    dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                r(first, second)
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            t = torch.tensor([min(ms)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item())
            ws_bytes = r.plan.workspace.numel()
            del r
            torch.cuda.empty_cache()
        except (RuntimeError, ops.EazError) as e:  # e.g. out of memory for the largest points
            if rank == 0:
                print(json.dumps({"workload": "C5", "envs": B, "num_simulations": n, "n_gpus": world, "error": str(e)[:120]}), flush=True)
            torch.cuda.empty_cache()
            continue
        if rank == 0:
            line = {"workload": "C5 Subleq ws=16 reanalyze", "envs": B, "envs_per_gpu": Bl, "num_simulations": n, "n_gpus": world, "ms_per_call": best,
                    "env_searches_per_s": B / (best * 1e-3), "simulations_per_s": B * n / (best * 1e-3), "workspace_gib_per_gpu": ws_bytes / 2 ** 30,
                    "mlp_mode": "tensor", "cuda_graph": True}
            if n in cpu:
                line["cpu_port_env_searches_per_s"] = cpu[n]["envs_per_s"]
                line["cpu_cores"] = cpu[n]["cores"]
            print(json.dumps(line), flush=True)
    del first, second
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
