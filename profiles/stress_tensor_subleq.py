"""Stress aid: repeats one TENSOR-mode Subleq search and checks every node's network outputs against the fp32 oracle network on the
node's own stored state (the (a) half of tests/test_gpu_parity.py::test_search_tensor_mode), reporting WHICH nodes differ.
Usage (on a B200): [EAZ_STRESS_POISON=171] python profiles/stress_tensor_subleq.py [reps] [word_size] [binary] [B] [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from e_alphazero_b200 import _abi, ops
from oracle import oracle as O
from tests import helpers as H

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ws = int(sys.argv[2]) if len(sys.argv) > 2 else 20
binary = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
B = int(sys.argv[4]) if len(sys.argv) > 4 else 40
n = int(sys.argv[5]) if len(sys.argv) > 5 else 16
env = H.make_env("subleq", seed=11, word_size=ws, binary=binary)
net = H.make_net(env, seed=12, fill=0.5)
root = H.make_root(env, net, B, seed=13, beta_max=0.0)
denv, dnet = H.device_env(env), H.device_net(net)
cfg = _abi.default_search_config(mlp_mode=_abi.MLP_TENSOR, num_simulations=n, discount=0.97)
A = env.num_actions
bad_runs = 0
poison = int(os.environ.get("EAZ_STRESS_POISON", "-1"))  # byte pattern written over the allocator's free memory before every search
for rep in range(reps):
    if poison >= 0:  # whatever the search reads without having written it is then garbage on every box
        x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda").fill_(poison)
        del x
    got = {k: v.cpu().numpy() for k, v in ops.search(cfg, denv, dnet, H.device_root(env, denv, root), want_tree=True).items()}
    emb = got["embeddings"][:, 1:].reshape(B * n, -1)
    st = H.uncompact(env, emb)
    ev = O.mlp_forward_states(net, env, st)
    lg = ev["exploit_logits"]
    lg = lg - lg.max(1, keepdims=True)
    term = st["terminated"].astype(bool)
    dl = np.abs(got["children_prior_logits"][:, 1:].reshape(B * n, A) - lg).max(1)
    dv = np.abs(got["raw_values"][:, 1:].reshape(-1) - np.where(term, 0, ev["value"]))
    du = np.abs(got["raw_values_epistemic_variance"][:, 1:].reshape(-1) - np.where(term, 0, ev["ube"]))
    bad = np.nonzero((dl > 1e-4) | (dv > 1e-4) | (du > 1e-4))[0]
    if len(bad):
        bad_runs += 1
        print(f"rep {rep}: {len(bad)} bad nodes:", [(int(i // n), int(i % n) + 1, float(dl[i]), float(dv[i]), float(du[i])) for i in bad[:12]], flush=True)
        if bad_runs <= 3:
            i = int(bad[0])
            g = got["children_prior_logits"][:, 1:].reshape(B * n, A)[i]
            print("   gpu row ", np.round(g, 4).tolist())
            print("   want row", np.round(lg[i], 4).tolist())
            print("   diff    ", np.round(g - lg[i], 4).tolist())
            print("   state: term", bool(term[i]), "step", int(st["step_count"][i]), "mem", st["memory"][i].tolist(), "visits", int(got["node_visits"][i // n, i % n + 1]),
                  "parent", int(got["parents"][i // n, i % n + 1]))
            # hypotheses about the policy head's layer 3 (K = 256 in 8 chunks of 32): a chunk missing, or a chunk's A operand stale
            head = _abi.HEAD_EXPLOIT
            obs = O.env_observe(env, {k: v[i : i + 1] for k, v in st.items()}).reshape(1, -1).astype(np.float32)
            h1 = np.maximum(obs @ net.w[head][0] + net.b[head][0], 0)
            h2 = np.maximum(h1 @ net.w[head][1] + net.b[head][1], 0)
            W3, b3 = net.w[head][2], net.b[head][2]
            full = (h2 @ W3 + b3)[0]
            gs = g - g.max()
            best = []
            for c in range(8):
                sl = slice(32 * c, 32 * c + 32)
                miss = full - (h2[:, sl] @ W3[sl])[0]
                best.append((float(np.abs((miss - miss.max()) - gs).max()), f"chunk {c} missing"))
                for c2 in range(8):
                    sl2 = slice(32 * c2, 32 * c2 + 32)
                    for nm, src in (("h1", h1), ("h2", h2)):
                        if nm == "h2" and c2 == c:
                            continue
                        alt = miss + (src[:, sl2] @ W3[sl])[0]
                        best.append((float(np.abs((alt - alt.max()) - gs).max()), f"chunk {c} <- {nm} chunk {c2}"))
            best.sort()
            print("   best hypotheses:", best[:4])
            # was another state evaluated?  distance of the GPU row to the expected rows of every node of the batch (and the roots)
            d_all = np.abs(lg - gs[None, :]).max(1)
            j = int(np.argmin(d_all))
            print("   nearest expected row among all nodes:", (j // n, j % n + 1), float(d_all[j]), "| own", float(d_all[i]))
            root_emb = got["embeddings"][:, 0]
            evr = O.mlp_forward_states(net, env, H.uncompact(env, root_emb))["exploit_logits"]
            evr = evr - evr.max(1, keepdims=True)
            dr = np.abs(evr - gs[None, :]).max(1)
            print("   nearest root row:", int(np.argmin(dr)), float(dr.min()))
            # the same state with one memory word / the step count changed
            base = {k: v[i : i + 1].copy() for k, v in st.items()}
            cand = []
            for w in range(ws):
                for val in range(ws):
                    c2 = {k: v.copy() for k, v in base.items()}
                    c2["memory"][0, w] = val
                    cand.append((w, val, c2))
            allc = {k: np.concatenate([c[2][k] for c in cand]) for k in base}
            evc = O.mlp_forward_states(net, env, allc)["exploit_logits"]
            evc = evc - evc.max(1, keepdims=True)
            dc = np.abs(evc - gs[None, :]).max(1)
            jj = int(np.argmin(dc))
            print("   nearest one-word variant: memory[%d] = %d" % (cand[jj][0], cand[jj][1]), float(dc[jj]))
            same_state = [int(j) for j in range(B * n) if (emb[j] == emb[i]).all()]
            print("   nodes with the same state:", [(j // n, j % n + 1, float(dl[j])) for j in same_state[:10]])
print(f"{bad_runs} of {reps} runs had mismatching nodes")
