"""Worker of tests/test_gpu_multi_rank.py: launched once per GPU by torch.distributed.run (NCCL).

Checks, on device tensors over NCCL (/root/reference/src/main.py:383-385,419, train.py:90-96):
  1. dist.broadcast_params: every rank ends with rank 0's parameters and hash bitset;
  2. shard invariance: `world` ranks x B/world envs through SelfplayRunner (graph replay, tensor-core network), gathered with
     dist.all_gather_trajectory, equal ONE rank stepping all B envs -- actions, rewards, flags and states bit for bit, for
     several consecutive steps;
  3. dist.merge_hash_sets: the OR-merged bitset equals the set built from the whole batch.
Rank 0 writes {"ok": true, ...} (or the failure) to the JSON file given as argv[1].
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main(out_path):
    import torch
    import torch.distributed as dist

    from e_alphazero_b200 import _abi, dist as D, ops
    from e_alphazero_b200.selfplay import SelfplayRunner
    from tests import helpers as H

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    report = {"world": world}
    for kind, kw, total, n, gamma in (("deepsea", dict(size=12), 1536, 24, 0.997), ("subleq", dict(word_size=16), 515, 16, 0.97)):
        env = H.make_env(kind, seed=0, **kw)
        ref_net = H.make_net(env, seed=1, fill=0.5)
        net = H.make_net(env, seed=1 if rank == 0 else 50 + rank, fill=0.5)  # other ranks start from different parameters
        denv, dnet = H.device_env(env), H.device_net(net)
        D.broadcast_params(dnet, src=0)
        for h in range(4):
            for l in range(3):
                assert (dnet.w[h][l].cpu().numpy() == ref_net.w[h][l]).all() and (dnet.b[h][l].cpu().numpy() == ref_net.b[h][l]).all()
        assert (dnet.binary_set.cpu().numpy() == ref_net.binary_set).all()

        lo, hi = D.shard_range(total, rank, world)
        states = H.random_states(env, total, seed=3)
        A = env.num_actions
        rng = np.random.default_rng(5)
        steps = 3
        gumbel = rng.gumbel(size=(steps, total, A)).astype(np.float32)
        tasks = rng.integers(1, 4, size=(steps, total)).astype(np.int32)
        full_beta = torch.linspace(0, 1, total, device=dev)

        def run(lo_, hi_, use_graph):
            B = hi_ - lo_
            r = SelfplayRunner(denv, dnet, B, n, gamma, exploration_beta=1.0, directed_exploration=True, mlp_mode=_abi.MLP_TENSOR, device=dev,
                               seed=0, use_graph=use_graph, fused_root=True, streams=2 if kind == "deepsea" else 1)
            r.beta = full_beta[lo_:hi_].contiguous()
            st = ops.state_to_device(denv, {k: np.ascontiguousarray(v[lo_:hi_]) for k, v in states.items()})
            trajs = []
            for s in range(steps):
                g = torch.as_tensor(gumbel[s, lo_:hi_]).to(dev)
                t = torch.as_tensor(tasks[s, lo_:hi_]).to(dev) if kind == "subleq" else None
                st, out = r.step(st, gumbel=g, task_ids=t)
                trajs.append(D.pack_trajectory(out.action, st["rewards"], st["terminated"], ops.env_compact(denv, st)).clone())
            return trajs, st

        mine, my_states = run(lo, hi, use_graph=True)
        gathered = [D.all_gather_trajectory(t) for t in mine]  # NCCL all-gather of uneven shards when total % world != 0
        # learner side of 8f-2: every rank marks ITS shard's observations, then the bitsets are OR-merged over NCCL
        bset = torch.zeros(1 << 21, dtype=torch.uint8, device=dev)
        ops.hash_update_(ops.env_observe(denv, my_states).to(torch.float32), bset, 24)
        D.merge_hash_sets(bset)
        if rank == 0:
            whole, whole_states = run(0, total, use_graph=False)
            for s in range(steps):
                a, b = gathered[s].cpu().numpy(), whole[s].cpu().numpy()
                assert a.shape == b.shape == (total, 4), (a.shape, b.shape)
                assert (a == b).all(), f"{kind}: step {s}: {int((a != b).any(1).sum())} of {total} trajectory rows differ between {world} shards and one rank"
            full = torch.zeros(1 << 21, dtype=torch.uint8, device=dev)
            ops.hash_update_(ops.env_observe(denv, whole_states).to(torch.float32), full, 24)
            assert bool((full == bset).all()), f"{kind}: merged hash set differs from the whole-batch set"
            report[kind] = {"envs": total, "steps": steps, "bits_set": int(torch.count_nonzero(bset).item())}
        dist.barrier()
    if rank == 0:
        report["ok"] = True
        json.dump(report, open(out_path, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main(sys.argv[1])
    except BaseException as e:  # noqa: BLE001
        if int(os.environ.get("RANK", "0")) == 0:
            json.dump({"ok": False, "error": repr(e)}, open(sys.argv[1], "w"))
        raise
