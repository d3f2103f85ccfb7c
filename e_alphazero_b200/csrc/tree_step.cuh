// tree_step.cuh -- the per-simulation tree kernel of the search (included by search.cu after the
// lane-group helpers).  ONE WARP PER TREE.  One launch does, for its tree:
//   1. expand  (mctx search.py expand / update_tree_node + the glue of context.py:132-154) for the leaf chosen
//              by the previous launch, using the network outputs;
//   2. backward (A.5) along the recorded path: lane l owns level l; the value / variance recurrences run as a
//              short uniform loop over levels (exact op order), the per-level running means (the divisions)
//              are computed by all levels in parallel;
//   3. action refresh: every node whose statistics just changed (the path + the new leaf) gets its NEXT action
//              selection recomputed NOW -- qtransform + seq-halving root score / interior score (A.3, A.6, A.7) --
//              32/G nodes at a time, G lanes per node, and cached in the node record together with the child it
//              leads to.  This is the same arithmetic on the same inputs the descent would evaluate, but all
//              levels are independent here, so they run in parallel instead of one after the other;
//   4. simulate (A.3) for the next simulation: a pure pointer chase over the cached (action, child) pairs, one
//              16-byte load per level, recording the path;  [+ the DeepSea transition, context.py:127].
#pragma once

namespace eaz {

// cached selection in NodeRec.pad0: bits 0..7 action, bits 8.. child index + 1 (0 = unvisited)
__device__ __forceinline__ int pack_next(int action, int child) { return action | ((child + 1) << 8); }

template <int G, int J>
__global__ void __launch_bounds__(128) tree_step_kernel(Tree t, SearchParams sp, EnvDesc env, int sim, int do_backward, int do_select,
                                                         const float* __restrict__ beta_in, const uint8_t* __restrict__ invalid,
                                                         unsigned long long* tl) {
  unsigned long long t_entry = 0;
  if (tl && blockIdx.x == 0 && threadIdx.x == 0) t_entry = globaltimer_ns();
  pdl_trigger();  // the next kernel (network / Subleq step) may begin its prologue now
  pdl_wait();     // ... and this one starts only once the previous kernel's results are visible
  unsigned long long t_wait = 0;
  if (tl && blockIdx.x == 0 && threadIdx.x == 0) t_wait = globaltimer_ns();
  struct TlExit {  // records at every exit path of the first warp
    unsigned long long *tl, a, b;
    __device__ ~TlExit() {
      if (tl && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long i = atomicAdd(tl, 1ull);
        if (i < 2000) { tl[8 + 4 * i] = a; tl[9 + 4 * i] = b; tl[10 + 4 * i] = globaltimer_ns(); tl[11 + 4 * i] = 0; }
      }
    }
  } tl_exit{tl, t_entry, t_wait};
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per tree
  if (b >= t.B) return;                                        // warp-uniform
  const unsigned uB = (unsigned)t.B, uA = (unsigned)t.A, ub = (unsigned)b;
  constexpr int kLevelsPerRound = 32 / G;
  const int gl = lane & (G - 1), glev = lane / G;
  bool valid[J];
#pragma unroll
  for (int j = 0; j < J; ++j) valid[j] = (gl + G * j) < t.A;

  int L = 0, leaf = 0;
  if (do_backward) {
    // ------------------------------------------------------------------ 1. expand
    leaf = t.leaf[b];
    L = t.path_len[b];
    const unsigned lslot = (unsigned)leaf * uB + ub;
    float m = -INFINITY;
    for (int a = lane; a < t.A; a += 32) m = fmaxf(m, t.net_logits[ub * uA + a]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));  // context.py:135
    for (int a = lane; a < t.A; a += 32)
      t.edges[(size_t)(lslot * uA + a)].pl = __fsub_rn(t.net_logits[ub * uA + a], m);  // legal_action_mask is all True (:137)
    int term;
    if (env.kind == EAZ_ENV_DEEPSEA) term = EAZ_DS_TERM(reinterpret_cast<const uint32_t*>(t.states)[lslot]);
    else term = t.states[(size_t)lslot * t.S + 35] & EAZ_SQ_FLAG_TERM;
    const float value = term ? 0.0f : t.net_value[b];  // :140
    const float var = term ? 0.0f : t.net_ube[b];      // :141
    float disc = sp.discount;
    if (sp.two_players) disc = __fmul_rn(disc, -1.0f);  // :142-143
    if (term) disc = 0.0f;                              // :144
    const int2 last = t.path[(unsigned)(L - 1) * uB + ub];
    EdgeRec* pe0 = t.edges + (size_t)(((unsigned)last.x * uB + ub) * uA + (unsigned)last.y);
    if (lane == 0) {
      // update_tree_node: the leaf's record (visits + 1: a leaf can be re-expanded under a max_depth cut-off)
      uint4* ln = reinterpret_cast<uint4*>(t.nodes + lslot);
      const int old_visits = (int)ln[0].x;
      ln[0] = make_uint4((unsigned)(old_visits + 1), __float_as_uint(value), __float_as_uint(var), 0u);
      ln[1] = make_uint4(__float_as_uint(value), __float_as_uint(var), (unsigned)(last.x + 1), (unsigned)(last.y + 1));
      pe0->ci1 = leaf + 1;
      pe0->rew = t.reward[b];  // :139
      pe0->dis = disc;
    }
    __syncwarp();

    // ------------------------------------------------------------------ 2. backward (levels L-1 .. 0)
    const bool std_backup = (sp.flags & EAZ_FLAG_BACKUP_STD) != 0;
    float lv = value, lvar = std_backup ? __fsqrt_rn(var) : var;  // running leaf_value / leaf variance (or std)
    float below_val = value, below_var = var;                     // updated mean / variance of the node one level deeper
    for (int hi = L; hi > 0; hi -= 32) {                          // rounds of 32 levels, deepest first
      const int lo = max(hi - 32, 0), cnt = hi - lo;
      const int lev = lo + lane;                                  // this lane's level
      const bool have = lane < cnt;
      int2 pa = make_int2(0, 0);
      uint4 nrec = make_uint4(0u, 0u, 0u, 0u);
      float rr = 0.0f, dd = 0.0f;
      int cvis = 0;
      unsigned pslot = 0;
      EdgeRec* pe = nullptr;
      if (have) {
        pa = t.path[(unsigned)lev * uB + ub];
        pslot = (unsigned)pa.x * uB + ub;
        pe = t.edges + (size_t)(pslot * uA + (unsigned)pa.y);
        nrec = reinterpret_cast<const uint4*>(t.nodes + pslot)[0];
        cvis = pe->vis;
        rr = pe->rew;
        dd = pe->dis;
      }
      // the recurrences: a uniform loop over this round's levels, deepest first, in the reference's op order
      float my_lv = 0.0f, my_lvar = 0.0f;
      for (int i = cnt - 1; i >= 0; --i) {
        const float r = __shfl_sync(0xffffffffu, rr, i), d = __shfl_sync(0xffffffffu, dd, i);
        lv = __fadd_rn(r, __fmul_rn(d, lv));
        lvar = std_backup ? __fadd_rn(0.0f, __fmul_rn(fabsf(d), lvar)) : __fadd_rn(0.0f, __fmul_rn(__fmul_rn(d, d), lvar));
        if (lane == i) { my_lv = lv; my_lvar = lvar; }
      }
      // running means, all levels in parallel
      const int nvis = (int)nrec.x;
      const float nval = __uint_as_float(nrec.y), nvar = __uint_as_float(nrec.z);
      const float count = (float)nvis;
      const float pv = __fdiv_rn(__fadd_rn(__fmul_rn(nval, count), my_lv), __fadd_rn(count, 1.0f));
      float pvar;
      if (std_backup) {
        const float ps = __fdiv_rn(__fadd_rn(__fmul_rn(__fsqrt_rn(nvar), count), my_lvar), __fadd_rn(count, 1.0f));
        pvar = __fmul_rn(ps, ps);
      } else {
        pvar = __fdiv_rn(__fadd_rn(__fmul_rn(nvar, count), my_lvar), __fadd_rn(count, 1.0f));
      }
      // children_values[parent, a] = the child's CURRENT (already updated) mean: one level deeper
      float cval = __shfl_down_sync(0xffffffffu, pv, 1), cvarr = __shfl_down_sync(0xffffffffu, pvar, 1);
      if (lane == cnt - 1) { cval = below_val; cvarr = below_var; }
      if (have) {
        reinterpret_cast<uint4*>(t.nodes + pslot)[0] = make_uint4((unsigned)(nvis + 1), __float_as_uint(pv), __float_as_uint(pvar), 0u);
        pe->vis = cvis + 1;
        *reinterpret_cast<float2*>(&pe->val) = make_float2(cval, cvarr);
      }
      below_val = __shfl_sync(0xffffffffu, pv, 0);
      below_var = __shfl_sync(0xffffffffu, pvar, 0);
    }
    __syncwarp();
  }
  if (!do_select) return;

  // -------------------------------------------------------------------- 3. refresh the cached selections
  // nodes: path levels 0..L-1 and the new leaf (index L); before the first simulation only the root.
  {
    const float beta = beta_in ? beta_in[b] : 0.0f;
    const int nrefresh = do_backward ? L + 1 : 1;
    for (int base = 0; base < nrefresh; base += kLevelsPerRound) {
      const int lev = base + glev;
      const bool act_on = lev < nrefresh;
      int node = 0;
      if (act_on && do_backward) node = lev < L ? t.path[(unsigned)lev * uB + ub].x : leaf;
      const unsigned slot = (unsigned)node * uB + ub;
      Edge<G, J> e;
      float raw, raw_var;
      load_edges<G, J>(t, slot, gl, act_on, e);
      load_node_raw(t, slot, act_on, raw, raw_var);
      float cq[J];
      int sumN, maxN, act;
      const bool is_root = act_on && node == 0;
      const bool use_beta = is_root || (sp.flags & EAZ_FLAG_BETA_INTERIOR) != 0;
      qtransform<G, J>(sp, e, valid, raw, raw_var, beta, use_beta, cq, sumN, maxN);
      if (__any_sync(0xffffffffu, is_root)) {  // gumbel_muzero_root_action_selection (node 0 is level 0 of round 0)
        float gum[J];
        bool inval[J];
        int num_valid = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int a = gl + G * j;
          gum[j] = (is_root && valid[j]) ? t.gumbel[ub * uA + a] : 0.0f;
          inval[j] = (is_root && valid[j] && invalid) ? (invalid[ub * uA + a] != 0) : false;
          num_valid += (valid[j] && !inval[j]) ? 1 : 0;
        }
        num_valid = group_sum_i<G>(num_valid);
        const int num_considered = min(sp.max_considered, num_valid);
        const int considered_visit = is_root ? t.table[num_considered * sp.n + min(sumN, sp.n - 1)] : 0;
        act = root_argmax<G, J>(e, valid, gum, inval, cq, considered_visit, gl);
      }
      int act_i;
      {  // gumbel_muzero_interior_action_selection
        float x[J], p[J];
#pragma unroll
        for (int j = 0; j < J; ++j) x[j] = __fadd_rn(e.pl[j], cq[j]);
        group_softmax<G, J>(x, valid, p);
        const float den = (float)(1 + sumN);
        float best = -INFINITY;
        int besti = 1 << 30;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float s = valid[j] ? __fsub_rn(p[j], __fdiv_rn((float)e.vis[j], den)) : -INFINITY;
          const int ia = valid[j] ? gl + G * j : (1 << 30);
          if (s > best || (s == best && ia < besti)) { best = s; besti = ia; }
        }
        act_i = group_argmax<G>(best, besti);
      }
      if (!is_root) act = act_i;
      // children_index[node, act]: owned by lane act % G of the group, slot act / G
      int ci_sel = -1;
#pragma unroll
      for (int j = 0; j < J; ++j) if (j == act / G) ci_sel = e.ci[j];
      const int child = __shfl_sync(0xffffffffu, ci_sel, (lane & ~(G - 1)) + (act & (G - 1)));
      if (act_on && gl == 0) t.nodes[slot].pad0 = pack_next(act, child);
    }
    __syncwarp();
  }

  // -------------------------------------------------------------------- 4. simulate: follow the cached selections
  // Nodes 0..sim exist.  For trees of up to 32*kChaseSlots nodes every cached (action, child) word is fetched once
  // (independent loads, lane i holds nodes i, i+32, ...) and the chase itself runs over registers with warp
  // shuffles; larger trees chase through memory, one dependent 16-byte load per level.
  constexpr int kChaseSlots = 9;
  int node = 0, depth = 0, action = 0, child = -1;
  if (sim < 32 * kChaseSlots) {
    int nxw[kChaseSlots];
#pragma unroll
    for (int s = 0; s < kChaseSlots; ++s) {
      const int nd = lane + 32 * s;
      nxw[s] = (32 * s <= sim && nd <= sim) ? t.nodes[(unsigned)nd * uB + ub].pad0 : 0;
    }
    while (true) {
      const int slot = node >> 5;
      int v = 0;
#pragma unroll
      for (int s = 0; s < kChaseSlots; ++s)
        if (s == slot) v = nxw[s];
      const int nx = __shfl_sync(0xffffffffu, v, node & 31);
      action = nx & 0xff;
      child = (nx >> 8) - 1;
      if (lane == 0) t.path[(unsigned)depth * uB + ub] = make_int2(node, action);
      depth += 1;
      if (child < 0 || depth >= sp.max_depth) break;
      node = child;
    }
  } else {
    while (true) {
      const int nx = t.nodes[(unsigned)node * uB + ub].pad0;
      action = nx & 0xff;
      child = (nx >> 8) - 1;
      if (lane == 0) t.path[(unsigned)depth * uB + ub] = make_int2(node, action);
      depth += 1;
      if (child < 0 || depth >= sp.max_depth) break;
      node = child;
    }
  }
  if (lane == 0) {
    const int new_leaf = child < 0 ? sim + 1 : child;  // search.py: node first expanded on simulation i gets index i+1
    t.path_len[b] = depth;
    t.parent[b] = node;
    t.action[b] = action;
    t.leaf[b] = new_leaf;
    if (env.kind == EAZ_ENV_DEEPSEA) {  // context.py:127 env.step fused here
      uint32_t* st = reinterpret_cast<uint32_t*>(t.states);
      float reward;
      const uint32_t ns = deepsea_step(st[(unsigned)node * uB + ub], action, env.size, env.action_map, &reward);
      st[(unsigned)new_leaf * uB + ub] = ns;
      t.reward[b] = reward;
      t.cell[b] = deepsea_obs_index(ns, env.size);
    }
  }
}

}  // namespace eaz
