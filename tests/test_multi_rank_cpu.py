"""world_size-2 gloo test of the N>1 host path: contiguous env sharding, parameter broadcast, and the
trajectory all-gather.  Compute on each rank is the CPU oracle (no GPU here); the check is shard invariance:
the gathered result of two half-batches equals one process searching the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from e_alphazero_b200 import _abi, dist as D
    from oracle import oracle as O
    from tests import helpers as H

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = H.make_env("deepsea", seed=0, size=8)
    # rank 0 owns the "learner" params; other ranks start from garbage and must receive them
    net = H.make_net(env, seed=1 if rank == 0 else 99, fill=0.5)

    class T:  # torch views over the numpy leaves so that broadcast writes in place
        w = [[torch.from_numpy(a) for a in hw] for hw in net.w]
        b = [[torch.from_numpy(a) for a in hb] for hb in net.b]
        binary_set = torch.from_numpy(net.binary_set)

    D.broadcast_params(T, src=0)
    lo, hi = D.shard_range(total, rank, world)
    states = H.random_states(env, total, seed=3)
    shard = {k: np.ascontiguousarray(v[lo:hi]) for k, v in states.items()}
    root = H.make_root(env, net, hi - lo, seed=0, states=shard)
    full_gumbel = np.random.default_rng(5).gumbel(size=(total, 2)).astype(np.float32)
    root["gumbel"] = full_gumbel[lo:hi]
    root["beta"] = np.zeros(hi - lo, np.float32)
    out = O.search(_abi.default_search_config(num_simulations=16), env, net, root, want_tree=False)
    nxt = O.env_step(env, shard, out["action"], auto_reset=True)
    traj = D.pack_trajectory(torch.from_numpy(out["action"]), torch.from_numpy(nxt["rewards"]), torch.from_numpy(nxt["terminated"]),
                             torch.from_numpy(O.env_compact(env, nxt)))
    gathered = D.all_gather_trajectory(traj)
    # learner side: each rank marks the observations of ITS shard as seen (hashes.py:45-50), then the sets are OR-merged
    bset = np.zeros(1 << 21, np.uint8)
    O.hash_update(O.env_observe(env, shard).astype(np.float32), bset, 24)
    tb = torch.from_numpy(bset)
    D.merge_hash_sets(tb)
    np.save(os.path.join(out_dir, f"bset{rank}.npy"), np.flatnonzero(bset))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_sharded_selfplay_step_matches_single_process(tmp_path, total):
    import torch.multiprocessing as mp

    from e_alphazero_b200 import _abi, dist as D
    from oracle import oracle as O
    from tests import helpers as H

    assert D.shard_range(10, 0, 4) == (0, 3) and D.shard_range(10, 3, 4) == (8, 10)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    # single-process reference over the whole batch
    env = H.make_env("deepsea", seed=0, size=8)
    net = H.make_net(env, seed=1, fill=0.5)
    states = H.random_states(env, total, seed=3)
    root = H.make_root(env, net, total, seed=0, states=states)
    root["gumbel"] = np.random.default_rng(5).gumbel(size=(total, 2)).astype(np.float32)
    root["beta"] = np.zeros(total, np.float32)
    out = O.search(_abi.default_search_config(num_simulations=16), env, net, root, want_tree=False)
    nxt = O.env_step(env, states, out["action"], auto_reset=True)
    assert got.shape == (total, 4)
    assert (got[:, 0] == out["action"]).all()
    assert (got[:, 1].view(np.float32) == nxt["rewards"][:, 0]).all()
    assert (got[:, 2] == nxt["terminated"]).all()
    assert (got[:, 3].view(np.uint32) == O.env_compact(env, nxt).view(np.uint32)[:, 0]).all()
    # merged hash sets: both ranks hold the set a single process builds from the whole batch
    full = np.zeros(1 << 21, np.uint8)
    O.hash_update(O.env_observe(env, H.random_states(env, total, seed=3)).astype(np.float32), full, 24)
    for r in range(2):
        assert (np.load(tmp_path / f"bset{r}.npy") == np.flatnonzero(full)).all()
