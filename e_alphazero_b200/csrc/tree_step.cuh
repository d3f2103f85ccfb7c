// tree_step.cuh -- the per-simulation tree kernel of the search (included by search.cu after the
// lane-group helpers).  One launch does, for every tree:
//   part 1  expand (mctx search.py expand / update_tree_node, glue of context.py:132-154) and backward
//           (A.5) for the leaf selected in the previous launch, using the network outputs;
//   part 2  simulate (A.3) for the next simulation [+ the DeepSea transition, context.py:127].
// The descent records its path (node, action per level) so that backward does not chase parent links:
// all levels' operands are independent loads issued together, and only the short value / variance
// recurrences are sequential.  For A == 2 (DeepSea) the descent prefetches both children of the current
// node while the node's scores are being computed, hiding the dependent-load latency of each level.
#pragma once

namespace eaz {

constexpr int kBackChunk = 8;

template <int G, int J>
__global__ void __launch_bounds__(128) tree_step_kernel(Tree t, SearchParams sp, EnvDesc env, int sim, int do_backward, int do_select,
                                                         const float* __restrict__ beta_in, const uint8_t* __restrict__ invalid) {
  EAZ_GROUP_PROLOGUE();

  // ======================================================================== part 1: expand + backward
  if (do_backward) {
    float lg[J];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      lg[j] = (in_range && valid[j]) ? t.net_logits[(size_t)b * t.A + gl + G * j] : 0.0f;
      if (valid[j]) m = fmaxf(m, lg[j]);
    }
    m = group_max<G>(m);  // context.py:135
    if (in_range) {
      const int leaf = t.leaf[b];
      const size_t lslot = (size_t)leaf * t.B + b;
#pragma unroll
      for (int j = 0; j < J; ++j)
        if (valid[j]) t.prior[lslot * t.A + gl + G * j] = __fsub_rn(lg[j], m);  // legal_action_mask is all True (:137)
      if (gl == 0) {
        int term;
        if (env.kind == EAZ_ENV_DEEPSEA) term = EAZ_DS_TERM(reinterpret_cast<const uint32_t*>(t.states)[lslot]);
        else term = t.states[lslot * t.S + 35] & EAZ_SQ_FLAG_TERM;
        const float value = term ? 0.0f : t.net_value[b];  // :140
        const float var = term ? 0.0f : t.net_ube[b];      // :141
        float disc = sp.discount;
        if (sp.two_players) disc = __fmul_rn(disc, -1.0f);  // :142-143
        if (term) disc = 0.0f;                              // :144
        const int L = t.path_len[b];
        const int2 last = t.path[(size_t)(L - 1) * t.B + b];
        // update_tree_node + edge (parent, action)
        t.raw_values[lslot] = value;
        t.node_values[lslot] = value;
        t.raw_var[lslot] = var;
        t.node_var[lslot] = var;
        t.node_visits[lslot] = t.node_visits[lslot] + 1;
        t.link[lslot] = last;
        const size_t pe0 = ((size_t)last.x * t.B + b) * t.A + last.y;
        t.children_index[pe0] = leaf;
        t.rewards[pe0] = t.reward[b];  // :139
        t.discounts[pe0] = disc;
        // backward: levels L-1 .. 0, kBackChunk levels of independent loads at a time
        const bool std_backup = (sp.flags & EAZ_FLAG_BACKUP_STD) != 0;
        float leaf_value = value, leaf_var = std_backup ? __fsqrt_rn(var) : var;
        float cur_val = value, cur_var = var;
        for (int hi = L; hi > 0; hi -= kBackChunk) {
          const int cnt = min(kBackChunk, hi);
          int2 pa[kBackChunk];
#pragma unroll
          for (int i = 0; i < kBackChunk; ++i)
            if (i < cnt) pa[i] = t.path[(size_t)(hi - 1 - i) * t.B + b];
          float nval[kBackChunk], nvar[kBackChunk], rr[kBackChunk], dd[kBackChunk];
          int nvis[kBackChunk], cvis[kBackChunk];
#pragma unroll
          for (int i = 0; i < kBackChunk; ++i)
            if (i < cnt) {
              const size_t pslot = (size_t)pa[i].x * t.B + b, pe = pslot * t.A + pa[i].y;
              nvis[i] = t.node_visits[pslot];
              nval[i] = t.node_values[pslot];
              nvar[i] = t.node_var[pslot];
              cvis[i] = t.children_visits[pe];
              rr[i] = t.rewards[pe];
              dd[i] = t.discounts[pe];
            }
#pragma unroll
          for (int i = 0; i < kBackChunk; ++i)
            if (i < cnt) {
              const size_t pslot = (size_t)pa[i].x * t.B + b, pe = pslot * t.A + pa[i].y;
              const float count = (float)nvis[i], d = dd[i];
              leaf_value = __fadd_rn(rr[i], __fmul_rn(d, leaf_value));
              const float pv = __fdiv_rn(__fadd_rn(__fmul_rn(nval[i], count), leaf_value), __fadd_rn(count, 1.0f));
              float pvar;
              if (std_backup) {
                leaf_var = __fadd_rn(0.0f, __fmul_rn(fabsf(d), leaf_var));
                const float ps = __fdiv_rn(__fadd_rn(__fmul_rn(__fsqrt_rn(nvar[i]), count), leaf_var), __fadd_rn(count, 1.0f));
                pvar = __fmul_rn(ps, ps);
              } else {
                leaf_var = __fadd_rn(0.0f, __fmul_rn(__fmul_rn(d, d), leaf_var));  // reward variance == 0 (context.py:149)
                pvar = __fdiv_rn(__fadd_rn(__fmul_rn(nvar[i], count), leaf_var), __fadd_rn(count, 1.0f));
              }
              t.node_values[pslot] = pv;
              t.node_var[pslot] = pvar;
              t.node_visits[pslot] = nvis[i] + 1;
              t.values[pe] = cur_val;  // the child's CURRENT mean (already updated)
              t.values_var[pe] = cur_var;
              t.children_visits[pe] = cvis[i] + 1;
              cur_val = pv;
              cur_var = pvar;
            }
        }
      }
    }
    __syncwarp();  // lane 0's tree updates are visible to the group's lanes below
  }
  if (!do_select) return;

  // ======================================================================== part 2: simulate
  const float beta = (in_range && beta_in) ? beta_in[b] : 0.0f;
  float gum[J];
  bool inval[J];
  int num_valid = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int a = gl + G * j;
    gum[j] = (in_range && valid[j]) ? t.gumbel[(size_t)b * t.A + a] : 0.0f;
    inval[j] = (in_range && valid[j] && invalid) ? (invalid[(size_t)b * t.A + a] != 0) : false;
    num_valid += (valid[j] && !inval[j]) ? 1 : 0;
  }
  num_valid = group_sum_i<G>(num_valid);
  const int num_considered = min(sp.max_considered, num_valid);
  constexpr bool kPrefetch = (G == 2 && J == 1);

  int node = 0, parent = 0, action = 0, next = 0, depth = 0, mylen = 0;
  bool cont = in_range;
  Edge<G, J> e;
  float raw = 0.0f, raw_var = 0.0f;
  if (kPrefetch) {
    load_edges<G, J>(t, (size_t)b, gl, cont, e);
    raw = cont ? t.raw_values[b] : 0.0f;
    raw_var = cont ? t.raw_var[b] : 0.0f;
  }
  while (__any_sync(0xffffffffu, cont)) {
    Edge<G, J> ec[2];
    float craw[2] = {0.0f, 0.0f}, cvar[2] = {0.0f, 0.0f};
    int child[2] = {-1, -1};
    if (kPrefetch) {  // both children of `node`, fetched while its scores are computed
      const int other = __shfl_xor_sync(0xffffffffu, e.ci[0], 1);
      child[0] = gl == 0 ? e.ci[0] : other;
      child[1] = gl == 0 ? other : e.ci[0];
      const bool deeper = cont && (depth + 1 < sp.max_depth);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool go = deeper && child[c] != -1;
        const size_t cslot = (size_t)(go ? child[c] : 0) * t.B + b;
        load_edges<G, J>(t, cslot, gl, go, ec[c]);
        craw[c] = go ? t.raw_values[cslot] : 0.0f;
        cvar[c] = go ? t.raw_var[cslot] : 0.0f;
      }
    } else {
      const size_t slot = (size_t)node * t.B + b;
      load_edges<G, J>(t, slot, gl, cont, e);
      raw = cont ? t.raw_values[slot] : 0.0f;
      raw_var = cont ? t.raw_var[slot] : 0.0f;
    }
    float cq[J];
    int sumN, maxN, act;
    if (depth == 0) {  // gumbel_muzero_root_action_selection (uniform: all trees start at the root together)
      qtransform<G, J>(sp, e, valid, raw, raw_var, beta, true, cq, sumN, maxN);
      const int considered_visit = cont ? t.table[(size_t)num_considered * sp.n + min(sumN, sp.n - 1)] : 0;
      act = root_argmax<G, J>(e, valid, gum, inval, cq, considered_visit, gl);
    } else {  // gumbel_muzero_interior_action_selection
      qtransform<G, J>(sp, e, valid, raw, raw_var, beta, (sp.flags & EAZ_FLAG_BETA_INTERIOR) != 0, cq, sumN, maxN);
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
      act = group_argmax<G>(best, besti);
    }
    int nxt;
    if (kPrefetch) {
      nxt = child[act & 1];
    } else {  // children_index[node, act]: owned by lane act % G, slot act / G
      int ci_sel = -1;
#pragma unroll
      for (int j = 0; j < J; ++j) if (j == act / G) ci_sel = e.ci[j];
      nxt = __shfl_sync(0xffffffffu, ci_sel, (lane & ~(G - 1)) + (act & (G - 1)));
    }
    if (cont) {
      if (gl == 0) t.path[(size_t)depth * t.B + b] = make_int2(node, act);
      parent = node;
      action = act;
      next = nxt;
      mylen = depth + 1;
    }
    depth += 1;
    if (cont) {
      cont = (nxt != -1) && (depth < sp.max_depth);
      if (cont) {
        node = nxt;
        if (kPrefetch) {
          e = ec[act & 1];
          raw = craw[act & 1];
          raw_var = cvar[act & 1];
        }
      }
    }
  }
  if (!in_range || gl != 0) return;
  const int leaf = next == -1 ? sim + 1 : next;  // search.py: node first expanded on simulation i gets index i+1
  t.path_len[b] = mylen;  // levels recorded for this tree; the last one is the leaf's parent
  t.parent[b] = parent;
  t.action[b] = action;
  t.leaf[b] = leaf;
  if (env.kind == EAZ_ENV_DEEPSEA) {  // context.py:127 env.step fused here
    uint32_t* st = reinterpret_cast<uint32_t*>(t.states);
    float reward;
    st[(size_t)leaf * t.B + b] = deepsea_step(st[(size_t)parent * t.B + b], action, env.size, env.action_map, &reward);
    t.reward[b] = reward;
  }
}

}  // namespace eaz
