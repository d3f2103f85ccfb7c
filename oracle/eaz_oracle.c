/*
 * eaz_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's algorithm for the self-play search
 * hot path, used only as the checker in tests/, __graft_entry__.smoke() and as
 * the timed CPU arm of bench.py.  Nothing under e_alphazero_b200/ may call it.
 *
 * PARITY STATUS
 *   - envs, hash, FC network, recurrent_fn glue: restated from the in-tree
 *     reference sources cited per function (paths under /root/reference/src)
 *     and pinned against golden vectors produced by executing those reference
 *     sources (tests/golden/, oracle/make_golden.py).
 *   - emctx search: "PARITY UNPINNED".  emctx is an un-vendored, un-versioned
 *     dependency (Pipfile:7, floating main.zip of YanivO1123/emctx, a fork of
 *     google-deepmind/mctx); its source is not in /root/reference and JAX is
 *     not installable here.  The search follows mctx's published algorithm
 *     (search.py / action_selection.py / qtransforms.py / seq_halving.py /
 *     policies.py / tree.py) plus the epistemic extension reconstructed in
 *     SURVEY.md Appendix A; every assumption there is a named EAZ_FLAG_*.
 *
 * It shares two headers with the product: include/eaz_b200.h (the POD structs,
 * here filled with HOST pointers) and include/eaz_math.h (the fp32 op-order
 * contract).  Reductions over the action axis use the fixed order documented
 * at orc_tree_sum() so results are bit-comparable with the CUDA kernels.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/eaz_b200.h"
#include "../include/eaz_math.h"

#define ORC_MAX_A 256
#define ORC_NEG_INF (-INFINITY)

/* ------------------------------------------------------------------------ */
/* threads                                                                   */

int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* Subleq test cases: get_test_cases, envs/subleq.py:398-501                  */

static int floormod(int x, int m) {
  int r = x % m;
  return r < 0 ? r + m : r;
}

typedef struct {
  int len;
  int v[8];
} orc_vec;

/* rows: [task 0..5][test 0..2] */
static const orc_vec k_inputs[6][3] = {
    /* NEGATION_POSITIVE :406-410 */
    {{7, {1, 2, 3, 4, 5, 6, 7}}, {8, {5, 4, 4, 5, 1, 2, 3, 1}}, {8, {1, 1, 6, 2, 4, 4, 5, 3}}},
    /* NEGATION :413-417 */
    {{8, {-4, -3, -2, -1, 0, 1, 2, 3}}, {7, {1, 2, 3, 4, 5, 6, 7}}, {5, {0, -1, 2, -3, 4}}},
    /* IDENTITY :420 */
    {{8, {-4, -3, -2, -1, 0, 1, 2, 3}}, {7, {1, 2, 3, 4, 5, 6, 7}}, {5, {0, -1, 2, -3, 4}}},
    /* SUBTRACTION :423-427 */
    {{6, {1, 1, 5, 4, 0, -3}}, {6, {2, 3, 0, 0, -1, -2}}, {8, {1, 2, 3, 4, 4, 3, 2, 1}}},
    /* ADDITION :434-438 */
    {{6, {1, 1, 5, 4, 0, -3}}, {6, {2, 3, 0, 0, -1, -2}}, {8, {1, 2, 3, 4, 4, 3, 2, 1}}},
    /* MULTIPLICATION (last switch branch) :445-449 */
    {{6, {1, 2, 2, 3, -7, 4}}, {6, {-1, -5, 5, 2, 0, 1}}, {6, {0, 0, 1, 1, -3, 3}}},
};
static const orc_vec k_outputs[6][3] = {
    {{7, {-1, -2, -3, -4, -5, -6, -7}}, {8, {-5, -4, -4, -5, -1, -2, -3, -1}}, {8, {-1, -1, -6, -2, -4, -4, -5, -3}}},
    {{8, {4, 3, 2, 1, 0, -1, -2, -3}}, {7, {-1, -2, -3, -4, -5, -6, -7}}, {5, {0, 1, -2, 3, -4}}},
    {{8, {-4, -3, -2, -1, 0, 1, 2, 3}}, {7, {1, 2, 3, 4, 5, 6, 7}}, {5, {0, -1, 2, -3, 4}}},
    {{3, {0, 1, 3}}, {3, {-1, 0, 1}}, {4, {-1, -1, 1, 1}}},
    {{3, {2, 9, -3}}, {3, {5, 0, -3}}, {4, {3, 7, 7, 3}}},
    {{3, {2, 6, -28}}, {3, {5, 10, 0}}, {3, {0, 1, -9}}},
};

/* prepare(): x % word_size, pad with word_size (:402-404, pad :101-107).
 * lax.switch clamps its index (:459-461). */
int orc_subleq_test_cases(int32_t task, int32_t ws, int32_t* inputs, int32_t* outputs) {
  if (ws < 16 || ws > 256) return EAZ_ERR_INVALID_ARG;
  int t = task - 1;
  if (t < 0) t = 0;
  if (t > 5) t = 5;
  for (int k = 0; k < 3; ++k) {
    for (int i = 0; i < 8; ++i) {
      inputs[k * 8 + i] = i < k_inputs[t][k].len ? floormod(k_inputs[t][k].v[i], ws) : ws;
      outputs[k * 8 + i] = i < k_outputs[t][k].len ? floormod(k_outputs[t][k].v[i], ws) : ws;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Subleq interpreter: simulate(), envs/subleq.py:156-395                     */

typedef struct {
  int in_after[8];
  int out_after[8];
  int bytes_used;
  int cycles;
  int correct;
} orc_sim;

static void orc_simulate(int ws, const int* memory, const int* test_in, const int* test_out, orc_sim* r) {
  const int AMAX = ws - 4, AIN = ws - 3, AOUT = ws - 2; /* :162-165 */
  int mem[256], in[8], out[8];
  memcpy(mem, memory, sizeof(int) * (size_t)ws);
  memcpy(in, test_in, sizeof(in));
  for (int i = 0; i < 8; ++i) out[i] = ws; /* :382 */
  int out_cur = 0, cur = 0, bytes = 0, cycles = 0, halt = 0, err = 0;
  while (!err && !halt && cycles < EAZ_SUBLEQ_MAX_CYCLES) { /* cond_fn :297-299 */
    if (cur + 2 >= ws) { /* :364-370 */
      cycles += 1;
      err = 1;
      continue;
    }
    cycles += 1;
    if (cur + 3 > bytes) bytes = cur + 3; /* :311 */
    const int a = mem[cur], b = mem[cur + 1], c = mem[cur + 2];
    /* read_memory :187-232 for a then b (both see the same input_state) */
    int va = 0, vb = 0, acc_a = 0, acc_b = 0, err_a = 0, err_b = 0;
    if (a <= AMAX) va = mem[a];
    else if (a == AIN) { if (in[0] >= ws) err_a = 1; else { va = in[0]; acc_a = 1; } }
    if (b <= AMAX) vb = mem[b];
    else if (b == AIN) { if (in[0] >= ws) err_b = 1; else { vb = in[0]; acc_b = 1; } }
    const int value = floormod(va - vb, ws); /* :322 */
    /* write_memory :243-295 */
    int modified = 0, err_w = 0;
    if (a <= AMAX) mem[a] = value;
    else if (a == AOUT) {
      if (out_cur >= 8) err_w = 1;
      else { out[out_cur] = value; out_cur += 1; modified = 1; }
    }
    const int jump = (value == 0) || (2 * value >= ws); /* :329, float compare value >= ws/2 */
    cur = jump ? c : cur + 3;
    if (acc_a || acc_b) { /* :333-338 */
      for (int i = 0; i < 7; ++i) in[i] = in[i + 1];
      in[7] = ws;
    }
    int all_eq = 1;
    for (int i = 0; i < 8; ++i) all_eq &= (out[i] == test_out[i]);
    halt = (((jump ? 1 : 0) & c) > AMAX) | all_eq; /* :340-345 -- precedence as written */
    err = (err_a | err_b | err_w) | (modified && (out[out_cur - 1] != test_out[out_cur - 1])); /* :346-350 */
  }
  memcpy(r->in_after, in, sizeof(in));
  memcpy(r->out_after, out, sizeof(out));
  r->bytes_used = bytes;
  r->cycles = cycles;
  int all_eq = 1;
  for (int i = 0; i < 8; ++i) all_eq &= (out[i] == test_out[i]);
  r->correct = !err && all_eq; /* :394 */
}

typedef struct {
  int solved;
  int in_after[8], out_after[8];
  int bytes_used, cycles_used;
} orc_tests;

/* run_tests(), envs/subleq.py:504-532 */
static void orc_run_tests(int ws, const int* memory, const int32_t* tin, const int32_t* tout, orc_tests* t) {
  t->solved = 1;
  t->bytes_used = 0;
  t->cycles_used = 0;
  for (int k = 0; k < 3; ++k) {
    orc_sim r;
    orc_simulate(ws, memory, tin + 8 * k, tout + 8 * k, &r);
    t->solved &= r.correct;
    if (r.bytes_used > t->bytes_used) t->bytes_used = r.bytes_used;
    if (r.cycles > t->cycles_used) t->cycles_used = r.cycles;
    if (k == 0) {
      memcpy(t->in_after, r.in_after, sizeof(r.in_after));
      memcpy(t->out_after, r.out_after, sizeof(r.out_after));
    }
  }
}

/* debugging / golden-vector entry: one simulate() call */
void orc_subleq_simulate(int32_t ws, const int32_t* memory, const int32_t* test_in, const int32_t* test_out,
                         int32_t* in_after, int32_t* out_after, int32_t* bytes_cycles_correct) {
  orc_sim r;
  orc_simulate(ws, memory, test_in, test_out, &r);
  memcpy(in_after, r.in_after, sizeof(r.in_after));
  memcpy(out_after, r.out_after, sizeof(r.out_after));
  bytes_cycles_correct[0] = r.bytes_used;
  bytes_cycles_correct[1] = r.cycles;
  bytes_cycles_correct[2] = r.correct;
}

static float orc_subleq_reward(int reward_fn, const orc_tests* t) {
  if (reward_fn == EAZ_SUBLEQ_REWARD_LOWEST_BYTES) return (float)t->solved / (float)(1 + t->bytes_used); /* :540-542 */
  return (float)t->solved; /* :535-537 */
}

/* ------------------------------------------------------------------------ */
/* Single-env state (what one pgx.State row carries)                          */

typedef struct {
  int step_count, terminated, truncated;
  float reward;
  int col;           /* DeepSea */
  int memory[256];   /* Subleq */
  int task, solved;
  int in_after[8], out_after[8];
} orc_env_state;

static void orc_load(const eaz_env* env, const eaz_state* s, int b, orc_env_state* e) {
  memset(e, 0, sizeof(*e));
  e->step_count = s->step_count[b];
  e->terminated = s->terminated[b] != 0;
  e->truncated = s->truncated ? (s->truncated[b] != 0) : 0;
  e->reward = s->rewards ? s->rewards[b] : 0.0f;
  if (env->kind == EAZ_ENV_DEEPSEA) {
    e->col = s->col[b];
  } else {
    const int ws = env->word_size;
    for (int i = 0; i < ws; ++i) e->memory[i] = s->memory[(size_t)b * ws + i];
    e->task = s->task[b];
    e->solved = s->solved[b] != 0;
    for (int i = 0; i < 8; ++i) {
      e->in_after[i] = s->input_after[(size_t)b * 8 + i];
      e->out_after[i] = s->output_after[(size_t)b * 8 + i];
    }
  }
}

static void orc_store(const eaz_env* env, eaz_state* s, int b, const orc_env_state* e) {
  s->step_count[b] = e->step_count;
  s->terminated[b] = (uint8_t)e->terminated;
  if (s->truncated) s->truncated[b] = (uint8_t)e->truncated;
  if (s->rewards) s->rewards[b] = e->reward;
  if (env->kind == EAZ_ENV_DEEPSEA) {
    s->col[b] = e->col;
  } else {
    const int ws = env->word_size;
    for (int i = 0; i < ws; ++i) s->memory[(size_t)b * ws + i] = e->memory[i];
    s->task[b] = e->task;
    s->solved[b] = (uint8_t)e->solved;
    for (int i = 0; i < 8; ++i) {
      s->input_after[(size_t)b * 8 + i] = e->in_after[i];
      s->output_after[(size_t)b * 8 + i] = e->out_after[i];
    }
  }
}

/* DeepSea._init deep_sea.py:54-57; Subleq._init subleq.py:623-646 */
static void orc_init_one(const eaz_env* env, int task, orc_env_state* e) {
  memset(e, 0, sizeof(*e));
  if (env->kind == EAZ_ENV_SUBLEQ) {
    const int ws = env->word_size;
    int32_t tin[24], tout[24];
    orc_subleq_test_cases(task, ws, tin, tout);
    orc_tests t;
    orc_run_tests(ws, e->memory, tin, tout, &t); /* empty program :632 */
    e->task = task;
    e->reward = orc_subleq_reward(env->reward_fn, &t);
    e->solved = t.solved;
    memcpy(e->in_after, t.in_after, sizeof(t.in_after));
    memcpy(e->out_after, t.out_after, sizeof(t.out_after));
  }
}

/* pgx.Env.step (pgx core.py) around
 * DeepSea._step deep_sea.py:59-81 / Subleq._step subleq.py:648-677. */
static void orc_step_one(const eaz_env* env, const uint8_t* action_map, orc_env_state* e, int action) {
  if (e->terminated || e->truncated) { /* absorbing: same state, zero rewards */
    e->reward = 0.0f;
    return;
  }
  e->step_count += 1; /* incremented BEFORE _step */
  if (env->kind == EAZ_ENV_DEEPSEA) {
    const int N = env->size;
    int row = e->step_count - 1, colc = e->col; /* jax gathers clamp out-of-range indices */
    if (row < 0) row = 0;
    if (row > N - 1) row = N - 1;
    if (colc < 0) colc = 0;
    if (colc > N - 1) colc = N - 1;
    const int flip = action_map ? (action_map[row * N + colc] != 0) : 0; /* :62 */
    const int shift = ((action == 0) ^ flip) ? -1 : 1;                  /* :63 */
    int c = e->col + shift;                                              /* :64 */
    if (c < 0) c = 0;
    if (c > N - 1) c = N - 1;
    e->col = c;
    e->terminated = e->step_count >= N - 1;                              /* :72 */
    e->reward = (e->terminated && c == N - 1) ? 1.0f : 0.0f;             /* :74-78 */
  } else {
    const int ws = env->word_size;
    if (e->step_count >= ws - 3 || e->solved) { /* :671-673 */
      e->terminated = 1;
      e->reward = 0.0f;
      return;
    }
    int32_t tin[24], tout[24];
    orc_subleq_test_cases(e->task, ws, tin, tout);
    e->memory[e->step_count - 1] = action; /* :654 */
    orc_tests t;
    orc_run_tests(ws, e->memory, tin, tout, &t);
    e->reward = orc_subleq_reward(env->reward_fn, &t);
    memcpy(e->in_after, t.in_after, sizeof(t.in_after));
    memcpy(e->out_after, t.out_after, sizeof(t.out_after));
    e->solved = t.solved;
  }
  /* illegal-action branch is dead: legal_action_mask is all True for both envs
   * (deep_sea.py:19, subleq.py:626). */
}

static int orc_check_env(const eaz_env* env) {
  if (env->kind == EAZ_ENV_DEEPSEA) return env->size >= 1 && env->size <= 4095 ? 0 : EAZ_ERR_INVALID_ARG;
  if (env->kind == EAZ_ENV_SUBLEQ) return (env->word_size >= 16 && env->word_size <= 256) ? 0 : EAZ_ERR_INVALID_ARG;
  return EAZ_ERR_INVALID_ARG;
}

int32_t orc_binary_width(int ws) { /* binary_encoding_width subleq.py:59-60 */
  int x = ws - 1, n = 0;
  while (x > 0) { n++; x >>= 1; }
  return n + 1;
}

int32_t orc_env_num_actions(const eaz_env* env) { return env->kind == EAZ_ENV_DEEPSEA ? 2 : env->word_size; }
int32_t orc_env_obs_cols(const eaz_env* env) {
  if (env->kind == EAZ_ENV_DEEPSEA) return env->size;
  return env->binary_encoding ? orc_binary_width(env->word_size) : env->word_size + 1;
}
int32_t orc_env_obs_dim(const eaz_env* env) {
  if (env->kind == EAZ_ENV_DEEPSEA) return env->size * env->size;
  return (env->word_size + 32) * orc_env_obs_cols(env); /* subleq.py:618-621 */
}
int32_t orc_env_hash_dim(const eaz_env* env, int32_t hash_io) {
  if (env->kind == EAZ_ENV_SUBLEQ && hash_io) return 32 * orc_env_obs_cols(env);
  return orc_env_obs_dim(env);
}

/* observation of one state into bool bytes [obs_dim] */
static void orc_observe_one(const eaz_env* env, const orc_env_state* e, uint8_t* obs) {
  const int D = orc_env_obs_dim(env);
  memset(obs, 0, (size_t)D);
  if (env->kind == EAZ_ENV_DEEPSEA) {
    const int N = env->size;
    int row = e->step_count < N - 1 ? e->step_count : N - 1; /* deep_sea.py:68-70; init cell (0,0) :56 */
    if (row < 0) row = 0;
    obs[row * N + e->col] = 1;
    return;
  }
  const int ws = env->word_size, w = orc_env_obs_cols(env);
  int32_t tin[24], tout[24];
  orc_subleq_test_cases(e->task, ws, tin, tout);
  /* concatenation order subleq.py:697-705: memory, example in, IN-after, example out, OUT-after */
  const int* parts[5] = {e->memory, tin, e->in_after, tout, e->out_after};
  const int lens[5] = {ws, 8, 8, 8, 8};
  int row = 0;
  for (int p = 0; p < 5; ++p) {
    for (int i = 0; i < lens[p]; ++i, ++row) {
      const int v = parts[p][i];
      uint8_t* o = obs + (size_t)row * w;
      if (env->binary_encoding) { /* subleq.py:88-97 */
        const unsigned m = (unsigned)floormod(v, ws) & 0xffu; /* astype(uint8) */
        for (int bit = 0; bit < w; ++bit) o[bit] = bit < 8 ? ((m >> bit) & 1u) : 0;
        o[w - 1] = (v == ws);
      } else { /* one-hot subleq.py:51-55 */
        o[v == ws ? ws : floormod(v, ws)] = 1;
      }
    }
  }
}

int orc_env_init(const eaz_env* env, const int32_t* task_ids, eaz_state* out, int32_t B) {
  if (orc_check_env(env)) return EAZ_ERR_INVALID_ARG;
  const int D = orc_env_obs_dim(env);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    orc_env_state e;
    orc_init_one(env, task_ids ? task_ids[b] : 1, &e);
    orc_store(env, out, b, &e);
    if (out->observation) orc_observe_one(env, &e, out->observation + (size_t)b * D);
  }
  return 0;
}

int orc_env_step(const eaz_env* env, eaz_state* state, const int32_t* action, int32_t auto_reset,
                 const int32_t* task_ids, int32_t B) {
  if (orc_check_env(env)) return EAZ_ERR_INVALID_ARG;
  const int D = orc_env_obs_dim(env);
#pragma omp parallel for schedule(dynamic, 64)
  for (int b = 0; b < B; ++b) {
    orc_env_state e;
    orc_load(env, state, b, &e);
    if (auto_reset && (e.terminated || e.truncated)) { /* selfplay.py:66-71 */
      orc_init_one(env, task_ids ? task_ids[b] : e.task, &e);
    } else {
      orc_step_one(env, env->action_map, &e, action[b]);
    }
    orc_store(env, state, b, &e);
    if (state->observation) orc_observe_one(env, &e, state->observation + (size_t)b * D);
  }
  return 0;
}

int orc_env_observe(const eaz_env* env, const eaz_state* state, uint8_t* observation, int32_t B) {
  if (orc_check_env(env)) return EAZ_ERR_INVALID_ARG;
  const int D = orc_env_obs_dim(env);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    orc_env_state e;
    orc_load(env, state, b, &e);
    orc_observe_one(env, &e, observation + (size_t)b * D);
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* XXHash variant: XXHash.get_indices, network/hashes.py:162-229              */

static uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

static uint32_t orc_xxhash_row(const uint32_t* x, int D, int bits) {
  const uint32_t P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, SEED = 1u;
  const int L = D / 4; /* reshape [B,4,L]: lane l = contiguous quarter l (:211-213) */
  uint32_t acc[4] = {SEED + P1 + P2, SEED + P2, SEED + 0u, SEED - P1}; /* :217-220 */
  for (int i = 0; i < L; ++i)
    for (int l = 0; l < 4; ++l) acc[l] = rotl32(acc[l] + x[l * L + i] * P2, 13) * P1; /* round :176-181 */
  uint32_t h = rotl32(acc[0], 1) + rotl32(acc[1], 7) + rotl32(acc[2], 12) + rotl32(acc[3], 18); /* :187 */
  h += (uint32_t)L; /* :226 (stripe count, not bytes) */
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16; /* avalanche :190-197 */
  return bits >= 32 ? h : (h >> (32 - bits)); /* :229 */
}

int orc_xxhash_indices(const float* x, int32_t B, int32_t D, int32_t bits, uint32_t* indices) {
  if (D % 4 != 0 || D <= 0 || bits <= 0 || bits > 32) return EAZ_ERR_INVALID_ARG; /* :154, :210 */
  for (int b = 0; b < B; ++b) indices[b] = orc_xxhash_row((const uint32_t*)(x + (size_t)b * D), D, bits);
  return 0;
}

/* BaseHash.__call__ hashes.py:23-38 */
int orc_hash_lookup(const float* x, int32_t B, int32_t D, int32_t bits, const uint8_t* binary_set, uint8_t* seen) {
  if (D % 4 != 0 || D <= 0 || bits <= 0 || bits > 32) return EAZ_ERR_INVALID_ARG;
  for (int b = 0; b < B; ++b) {
    const uint32_t idx = orc_xxhash_row((const uint32_t*)(x + (size_t)b * D), D, bits);
    seen[b] = (binary_set[idx >> 3] & (1u << (idx & 7u))) != 0;
  }
  return 0;
}

/* BaseHash.update hashes.py:45-50 */
int orc_hash_update(const float* x, int32_t B, int32_t D, int32_t bits, uint8_t* binary_set) {
  if (D % 4 != 0 || D <= 0 || bits <= 0 || bits > 32) return EAZ_ERR_INVALID_ARG;
  for (int b = 0; b < B; ++b) {
    const uint32_t idx = orc_xxhash_row((const uint32_t*)(x + (size_t)b * D), D, bits);
    binary_set[idx >> 3] |= (uint8_t)(1u << (idx & 7u));
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Convolutional evaluators in inference mode: EpistemicResidualAZNet          */
/* (network/resnet.py:41-135) and EpistemicMinatarAZNet (network/minatar.py:   */
/* 11-114).  Contract shared with csrc/convnet.cu: hk.Conv2D (SAME, NHWC/HWIO) */
/* = acc = 0; for kh, kw, ci ascending: acc = fma(x, w, acc); + b; hk.Linear =  */
/* the chain over k ascending; hk.BatchNorm (inference) = (x - mean) * (scale * */
/* 1/sqrt(var + 1e-5)) + offset, every operation rounded separately.           */

static float orc_bn(const eaz_bn* bn, int c, float x) {
  const float inv = eaz_mul(bn->scale[c], eaz_div(1.0f, eaz_sqrt(eaz_add(bn->var[c], 1e-5f))));
  return eaz_add(eaz_mul(eaz_sub(x, bn->mean[c]), inv), bn->offset[c]);
}
/* one board: in [H,W,Cin] -> out [H,W,Cout]; pre: BatchNorm + relu on the input (the padding stays zero); post: BatchNorm on the output */
static void orc_conv3x3(const float* in, int H, int W, int Cin, int Cout, const eaz_conv* c, const eaz_bn* pre, const eaz_bn* post,
                        const float* residual, int relu_out, float* out) {
  float* x = (float*)malloc(sizeof(float) * (size_t)H * W * Cin);
  for (int i = 0; i < H * W * Cin; ++i) {
    float v = in[i];
    if (pre) v = eaz_max(orc_bn(pre, i % Cin, v), 0.0f);
    x[i] = v;
  }
  for (int y = 0; y < H; ++y)
    for (int xx = 0; xx < W; ++xx)
      for (int co = 0; co < Cout; ++co) {
        float acc = 0.0f;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int yy = y + kh - 1, xw = xx + kw - 1;
            const int inside = yy >= 0 && yy < H && xw >= 0 && xw < W;
            for (int ci = 0; ci < Cin; ++ci) {
              const float xv = inside ? x[(yy * W + xw) * Cin + ci] : 0.0f; /* SAME padding: zeros take part in the chain */
              acc = eaz_fma(xv, c->w[((size_t)(kh * 3 + kw) * Cin + ci) * Cout + co], acc);
            }
          }
        float r = eaz_add(acc, c->b[co]);
        if (post) r = orc_bn(post, co, r);
        const size_t o = (size_t)(y * W + xx) * Cout + co;
        if (residual) r = eaz_add(r, residual[o]);
        if (relu_out) r = eaz_max(r, 0.0f);
        out[o] = r;
      }
  free(x);
}
static void orc_dense(const float* x, int K, int N, const eaz_conv* l, const eaz_bn* pre, const eaz_bn* post, int relu, float* y) {
  for (int n = 0; n < N; ++n) {
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) {
      float v = x[k];
      if (pre) v = eaz_max(orc_bn(pre, k, v), 0.0f);
      acc = eaz_fma(v, l->w[(size_t)k * N + n], acc);
    }
    acc = eaz_add(acc, l->b[n]);
    if (post) acc = orc_bn(post, n, acc);
    if (relu) acc = eaz_max(acc, 0.0f);
    y[n] = acc;
  }
}

int orc_convnet_forward(const eaz_convnet_params* net, const uint8_t* observation, int32_t B, float* exploit_logits, float* explore_logits,
                        float* value, float* ube, float* novelty) {
  const int H = net->height, W = net->width, C0 = net->in_channels, C = net->num_channels, A = net->num_actions, HW = H * W, Hd = net->hidden;
  const int D = HW * C0;
  if (D % 4 != 0 || net->hash_bits <= 0 || net->hash_bits > 32) return EAZ_ERR_INVALID_ARG; /* hashes.py:154,210 */
  int fail = 0;
#pragma omp parallel for schedule(dynamic, 4)
  for (int b = 0; b < B; ++b) {
    float* x0 = (float*)malloc(sizeof(float) * (size_t)D);
    float* bufs[3];
    for (int i = 0; i < 3; ++i) bufs[i] = (float*)malloc(sizeof(float) * (size_t)HW * C);
    float* hc = (float*)malloc(sizeof(float) * (size_t)HW * 2);
    float* hf = (float*)malloc(sizeof(float) * (size_t)(C > Hd ? C : Hd) * 2);
    float* lg = (float*)malloc(sizeof(float) * (size_t)A);
    for (int i = 0; i < D; ++i) x0[i] = observation[(size_t)b * D + i] ? 1.0f : 0.0f; /* x.astype(float32) */
    float v_raw = 0.0f, u_raw = 0.0f;
    if (net->kind == EAZ_CONVNET_RESNET) {
      const int v2 = net->resnet_v2 != 0;
      orc_conv3x3(x0, H, W, C0, C, &net->stem, NULL, v2 ? NULL : &net->stem_bn, NULL, !v2, bufs[0]); /* :70-75 */
      int cur = 0;
      for (int i = 0; i < net->num_blocks; ++i) {
        const int t = (cur + 1) % 3, o = (cur + 2) % 3;
        if (v2) { /* BlockV2 :28-43 */
          orc_conv3x3(bufs[cur], H, W, C, C, &net->block_conv[i][0], &net->block_bn[i][0], NULL, NULL, 0, bufs[t]);
          orc_conv3x3(bufs[t], H, W, C, C, &net->block_conv[i][1], &net->block_bn[i][1], NULL, bufs[cur], 0, bufs[o]);
        } else { /* BlockV1 :11-24 */
          orc_conv3x3(bufs[cur], H, W, C, C, &net->block_conv[i][0], NULL, &net->block_bn[i][0], NULL, 1, bufs[t]);
          orc_conv3x3(bufs[t], H, W, C, C, &net->block_conv[i][1], NULL, &net->block_bn[i][1], bufs[cur], 1, bufs[o]);
        }
        cur = o;
      }
      for (int h = 0; h < 4; ++h) { /* :84-124 */
        const int k = h < 2 ? 2 : 1;
        for (int p = 0; p < HW; ++p) orc_dense(bufs[cur] + (size_t)p * C, C, k, &net->head_conv[h], v2 ? &net->final_bn : NULL, &net->head_bn[h], 1, hc + p * k);
        if (h < 2) {
          orc_dense(hc, HW * k, A, &net->head_fc[h], NULL, NULL, 0, lg);
          float* dst = h == 0 ? exploit_logits : explore_logits;
          if (dst) memcpy(dst + (size_t)b * A, lg, sizeof(float) * (size_t)A);
        } else {
          float o1;
          orc_dense(hc, HW * k, C, &net->head_fc[h], NULL, NULL, 1, hf);
          orc_dense(hf, C, 1, &net->head_out[h], NULL, NULL, 0, &o1);
          if (h == 2) v_raw = o1; else u_raw = o1;
        }
      }
    } else { /* minatar.py:55-95 */
      float* tower[2] = {hf, hf + Hd};
      float* f1 = (float*)malloc(sizeof(float) * (size_t)Hd);
      float* hh = (float*)malloc(sizeof(float) * (size_t)Hd);
      for (int tw = 0; tw < 2; ++tw) {
        orc_conv3x3(x0, H, W, C0, C, &net->tower_conv[tw], NULL, NULL, NULL, 1, bufs[0]);
        orc_dense(bufs[0], HW * C, Hd, &net->tower_fc[tw][0], NULL, NULL, 1, f1);
        orc_dense(f1, Hd, Hd, &net->tower_fc[tw][1], NULL, NULL, 1, tower[tw]);
      }
      for (int h = 0; h < 4; ++h) { /* [0] main policy, [1] value, [2] exploration policy, [3] ube */
        const int policy = h == 0 || h == 2;
        orc_dense(tower[h >> 1], Hd, Hd, &net->mhead_fc[h][0], NULL, NULL, 1, hh);
        if (policy) {
          orc_dense(hh, Hd, A, &net->mhead_fc[h][1], NULL, NULL, 0, lg);
          float* dst = h == 0 ? exploit_logits : explore_logits;
          if (dst) memcpy(dst + (size_t)b * A, lg, sizeof(float) * (size_t)A);
        } else {
          float o1;
          orc_dense(hh, Hd, 1, &net->mhead_fc[h][1], NULL, NULL, 0, &o1);
          if (h == 1) v_raw = o1; else u_raw = o1;
        }
      }
      free(f1);
      free(hh);
    }
    const uint32_t idx = orc_xxhash_row((const uint32_t*)x0, D, net->hash_bits);
    const int seen = net->binary_set ? ((net->binary_set[idx >> 3] >> (idx & 7u)) & 1) : 0;
    const float nov = eaz_mul(seen ? 0.0f : 1.0f, net->novelty_scale);
    float v, u;
    if (net->kind == EAZ_CONVNET_RESNET) {
      v = eaz_tanh(v_raw);                                  /* resnet.py:102 */
      u = eaz_mul(0.5f, eaz_add(eaz_tanh(u_raw), 1.0f));    /* :114 */
      u = eaz_max(nov, u);                                  /* :126-128 */
    } else {
      v = v_raw;                                            /* minatar.py:69-70 */
      u = eaz_softplus(u_raw);                              /* :91 */
      u = eaz_max(eaz_mul(nov, net->local_unc_scale), u);   /* :101-103 */
      u = eaz_min(eaz_max(u, 0.0f), net->max_u);            /* :104 */
    }
    if (value) value[b] = v;
    if (ube) ube[b] = u;
    if (novelty) novelty[b] = nov;
    free(x0);
    for (int i = 0; i < 3; ++i) free(bufs[i]);
    free(hc);
    free(hf);
    free(lg);
  }
  return fail;
}

/* ------------------------------------------------------------------------ */
/* EpistemicFullyConnectedAZNet.__call__ (is_training=False),                  */
/* network/fully_connected.py:41-101                                          */

/* hk.Linear: y = x @ w + b.  Contract (shared with the CUDA EXACT mode):
 * acc = 0; for k ascending: acc = fma(x[k], w[k][j], acc); y = acc + b[j].
 * Zero inputs are skipped (fma(0,w,acc) == acc for finite w). */
static void orc_linear(const float* x, int K, const float* w, const float* b, int Nout, int relu, float* y) {
  for (int j = 0; j < Nout; ++j) y[j] = 0.0f;
  for (int k = 0; k < K; ++k) {
    const float xk = x[k];
    if (xk == 0.0f) continue;
    const float* wr = w + (size_t)k * Nout;
    for (int j = 0; j < Nout; ++j) y[j] = eaz_fma(xk, wr[j], y[j]);
  }
  for (int j = 0; j < Nout; ++j) {
    float v = eaz_add(y[j], b[j]);
    if (relu && !(v > 0.0f)) v = 0.0f; /* jax.nn.relu = max(x, 0) */
    y[j] = v;
  }
}

static void orc_head(const eaz_fc_params* net, int head, const float* x, int nout, float* out) {
  float h1[EAZ_FC_HIDDEN_MAX], h2[EAZ_FC_HIDDEN_MAX];
  orc_linear(x, net->in_dim, net->w[head][0], net->b[head][0], net->hidden, 1, h1);
  orc_linear(h1, net->hidden, net->w[head][1], net->b[head][1], net->hidden, 1, h2);
  orc_linear(h2, net->hidden, net->w[head][2], net->b[head][2], nout, 0, out);
}

typedef struct {
  float exploit[ORC_MAX_A], explore[ORC_MAX_A];
  float value, ube, novelty;
} orc_net_out;

/* obs: bool bytes [D]; heads_mask selects which heads to evaluate */
static void orc_net_one(const eaz_fc_params* net, const uint8_t* obs, int hash_dim, int heads_mask, orc_net_out* o) {
  const int D = net->in_dim, A = net->num_actions;
  float* x = (float*)malloc(sizeof(float) * (size_t)D);
  for (int i = 0; i < D; ++i) x[i] = obs[i] ? 1.0f : 0.0f; /* x.astype(float32) :45 */
  float t;
  if (heads_mask & (1 << EAZ_HEAD_VALUE)) {
    orc_head(net, EAZ_HEAD_VALUE, x, 1, &t);
    o->value = eaz_tanh(t); /* :55 */
  }
  if (heads_mask & (1 << EAZ_HEAD_EXPLOIT)) orc_head(net, EAZ_HEAD_EXPLOIT, x, A, o->exploit);
  if (heads_mask & (1 << EAZ_HEAD_EXPLORE)) orc_head(net, EAZ_HEAD_EXPLORE, x, A, o->explore);
  if (heads_mask & (1 << EAZ_HEAD_UBE)) {
    orc_head(net, EAZ_HEAD_UBE, x, 1, &t);
    float u = eaz_mul(0.5f, eaz_add(eaz_tanh(t), 1.0f)); /* :64 */
    /* hash_io: rows word_size.. of the observation = trailing hash_dim elements (:85-89) */
    const uint32_t idx = orc_xxhash_row((const uint32_t*)(x + (D - hash_dim)), hash_dim, net->hash_bits);
    const int seen = net->binary_set ? ((net->binary_set[idx >> 3] & (1u << (idx & 7u))) != 0) : 0;
    const float novelty = eaz_mul(seen ? 0.0f : 1.0f, net->novelty_scale); /* :90 */
    u = eaz_mul(u, net->max_u);                                         /* :93 */
    u = eaz_max(novelty, u);                                            /* :95 */
    u = eaz_min(eaz_max(u, 0.0f), net->max_u);                          /* :96 */
    o->ube = u;
    o->novelty = novelty;
  }
  free(x);
}

int orc_mlp_forward(const eaz_fc_params* net, const uint8_t* observation, int32_t B, int32_t hash_dim,
                    float* exploit_logits, float* explore_logits, float* value, float* ube, float* novelty) {
  const int A = net->num_actions;
  if (hash_dim % 4 != 0 || net->hidden > EAZ_FC_HIDDEN_MAX || A > ORC_MAX_A) return EAZ_ERR_INVALID_ARG;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    orc_net_out o;
    orc_net_one(net, observation + (size_t)b * net->in_dim, hash_dim, 0xF, &o);
    if (exploit_logits) memcpy(exploit_logits + (size_t)b * A, o.exploit, sizeof(float) * (size_t)A);
    if (explore_logits) memcpy(explore_logits + (size_t)b * A, o.explore, sizeof(float) * (size_t)A);
    if (value) value[b] = o.value;
    if (ube) ube[b] = o.ube;
    if (novelty) novelty[b] = o.novelty;
  }
  return 0;
}

int orc_mlp_forward_states(const eaz_fc_params* net, const eaz_env* env, const eaz_state* state, int32_t B,
                           float* exploit_logits, float* explore_logits, float* value, float* ube, float* novelty) {
  const int D = orc_env_obs_dim(env);
  if (D != net->in_dim) return EAZ_ERR_INVALID_ARG;
  uint8_t* obs = (uint8_t*)malloc((size_t)B * D);
  orc_env_observe(env, state, obs, B);
  int rc = orc_mlp_forward(net, obs, B, orc_env_hash_dim(env, net->hash_io), exploit_logits, explore_logits, value,
                           ube, novelty);
  free(obs);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* Reductions over the action axis -- the fixed order shared with the GPU.     */
/*                                                                            */
/* G = min(32, next_pow2(A)) lanes per tree; action a lives in lane a % G,    */
/* slot a / G.  A sum is: per lane, slots in ascending order starting from    */
/* 0.0f; then an xor-butterfly over lanes with strides 1, 2, 4, ... G/2.      */
/* Missing actions (a >= A) contribute 0.0f.  max/min/argmax are order-free   */
/* (argmax = lowest index among the maxima, like jnp.argmax).                 */

static int orc_group(int A) {
  int g = 1;
  while (g < A && g < 32) g <<= 1;
  return g < 2 ? 2 : g;
}

static float orc_tree_sum(const float* x, int A) {
  const int G = orc_group(A);
  float lane[32];
  for (int l = 0; l < G; ++l) {
    float acc = 0.0f;
    for (int a = l; a < A; a += G) acc = eaz_add(acc, x[a]);
    lane[l] = acc;
  }
  for (int s = 1; s < G; s <<= 1) {
    float nxt[32];
    for (int l = 0; l < G; ++l) nxt[l] = eaz_add(lane[l], lane[l ^ s]);
    memcpy(lane, nxt, sizeof(float) * (size_t)G);
  }
  return lane[0];
}

static float orc_maxv(const float* x, int A) {
  float m = x[0];
  for (int a = 1; a < A; ++a) if (x[a] > m) m = x[a];
  return m;
}
static float orc_minv(const float* x, int A) {
  float m = x[0];
  for (int a = 1; a < A; ++a) if (x[a] < m) m = x[a];
  return m;
}
static int orc_argmax(const float* x, int A) {
  int best = 0;
  for (int a = 1; a < A; ++a) if (x[a] > x[best]) best = a;
  return best;
}

/* jax.nn.softmax: exp(x - max) / sum */
static void orc_softmax(const float* x, int A, float* p) {
  const float m = orc_maxv(x, A);
  for (int a = 0; a < A; ++a) p[a] = eaz_exp(eaz_sub(x[a], m));
  const float s = orc_tree_sum(p, A);
  for (int a = 0; a < A; ++a) p[a] = eaz_div(p[a], s);
}

/* _mask_invalid_actions: mctx policies.py, copied at reanalyze.py:16-29 */
static void orc_mask_invalid(const float* logits, const uint8_t* invalid, int A, float* out) {
  const float m = orc_maxv(logits, A);
  for (int a = 0; a < A; ++a) out[a] = (invalid && invalid[a]) ? EAZ_F32_MIN : eaz_sub(logits[a], m);
}

/* ------------------------------------------------------------------------ */
/* Sequential halving schedule: mctx seq_halving.py                           */

static void orc_considered_visits(int m, int n, int32_t* seq) {
  if (m <= 1) {
    for (int i = 0; i < n; ++i) seq[i] = i;
    return;
  }
  int log2max = 0;
  while ((1 << log2max) < m) log2max++;
  int visits[ORC_MAX_A];
  for (int i = 0; i < m; ++i) visits[i] = 0;
  int nc = m, len = 0;
  while (len < n) {
    int extra = n / (log2max * nc);
    if (extra < 1) extra = 1;
    for (int e = 0; e < extra && len < n; ++e) {
      for (int i = 0; i < nc && len < n; ++i) seq[len++] = visits[i];
      for (int i = 0; i < nc; ++i) visits[i] += 1;
    }
    nc = nc / 2 > 2 ? nc / 2 : 2;
  }
}

/* table[(max_considered+1), n] (get_table_of_considered_visits) */
int orc_seq_halving_table(int32_t max_considered, int32_t n, int32_t* table) {
  if (max_considered < 0 || max_considered > ORC_MAX_A || n < 1) return EAZ_ERR_INVALID_ARG;
  for (int m = 0; m <= max_considered; ++m) orc_considered_visits(m, n, table + (size_t)m * n);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* The search (one env at a time; envs are independent)                        */

typedef struct {
  int N, A;
  int32_t *node_visits, *parents, *action_from_parent;
  float *raw_values, *node_values, *raw_var, *node_var;
  int32_t *children_index, *children_visits;
  float *prior_logits, *rewards, *discounts, *values, *rewards_var, *values_var;
  orc_env_state* emb;
} orc_tree;

/* Replay table: network outputs recorded per node by another implementation
 * (the GPU tree), so that the search logic can be checked bit-exactly under a
 * tensor-core network whose outputs differ from the EXACT mode in the last
 * bits.  Rows are keyed by the compact state bytes. */
typedef struct orc_replay {
  const uint8_t* states; /* [B,N,S] */
  const float* logits;   /* [B,N,A] post-glue prior logits */
  const float* value;    /* [B,N] */
  const float* var;      /* [B,N] */
  int32_t S;
} orc_replay;

/* Compact in-tree state encoding (documented in DESIGN.md "data layout"). */
int32_t orc_env_compact_bytes(const eaz_env* env) {
  if (env->kind == EAZ_ENV_DEEPSEA) return 4;
  return 40 + ((env->word_size + 7) / 8) * 8;
}

static void orc_compact(const eaz_env* env, const orc_env_state* e, uint8_t* out) {
  if (env->kind == EAZ_ENV_DEEPSEA) {
    const uint32_t v = ((uint32_t)e->step_count & 0xfffu) | (((uint32_t)e->col & 0xfffu) << 12) |
                       ((uint32_t)(e->terminated != 0) << 24) | ((uint32_t)(e->truncated != 0) << 25);
    memcpy(out, &v, 4);
    return;
  }
  const int S = orc_env_compact_bytes(env);
  memset(out, 0, (size_t)S);
  uint16_t* h = (uint16_t*)out;
  for (int i = 0; i < 8; ++i) h[i] = (uint16_t)e->in_after[i];
  for (int i = 0; i < 8; ++i) h[8 + i] = (uint16_t)e->out_after[i];
  h[16] = (uint16_t)e->step_count;
  out[34] = (uint8_t)e->task;
  out[35] = (uint8_t)((e->terminated != 0) | ((e->truncated != 0) << 1) | ((e->solved != 0) << 2));
  /* bytes 36..39: reserved (0) */
  for (int i = 0; i < env->word_size; ++i) out[40 + i] = (uint8_t)e->memory[i];
}

int orc_env_compact(const eaz_env* env, const eaz_state* state, uint8_t* out, int32_t B) {
  const int S = orc_env_compact_bytes(env);
  for (int b = 0; b < B; ++b) {
    orc_env_state e;
    orc_load(env, state, b, &e);
    orc_compact(env, &e, out + (size_t)b * S);
  }
  return 0;
}

typedef struct {
  const eaz_search_config* cfg;
  const eaz_env* env;
  const eaz_fc_params* net;
  int hash_dim;
  const orc_replay* replay;
  int b;
  int replay_miss;
} orc_ctx;

/* context.py:117-155 epistemic_recurrent_fn for one env */
static void orc_recurrent(orc_ctx* cx, int action, const orc_env_state* parent, orc_env_state* child, float* logits,
                          float* value, float* var, float* reward, float* discount) {
  const eaz_search_config* cfg = cx->cfg;
  const int A = orc_env_num_actions(cx->env);
  *child = *parent;
  orc_step_one(cx->env, cx->env->action_map, child, action); /* :127 */
  *reward = child->reward;                                   /* :139, current_player == 0 */
  if (cx->replay) {
    const int S = cx->replay->S, N = cfg->num_simulations + 1;
    uint8_t key[40 + 256];
    orc_compact(cx->env, child, key);
    int hit = -1; /* node 0 holds root-fn outputs, not recurrent_fn outputs: start at 1 */
    for (int i = 1; i < N && hit < 0; ++i)
      if (memcmp(cx->replay->states + ((size_t)cx->b * N + i) * S, key, (size_t)S) == 0) hit = i;
    if (hit < 0) {
      cx->replay_miss += 1;
      for (int a = 0; a < A; ++a) logits[a] = 0.0f;
      *value = 0.0f;
      *var = 0.0f;
    } else {
      memcpy(logits, cx->replay->logits + ((size_t)cx->b * N + hit) * A, sizeof(float) * (size_t)A);
      *value = cx->replay->value[(size_t)cx->b * N + hit];
      *var = cx->replay->var[(size_t)cx->b * N + hit];
    }
  } else {
    uint8_t* obs = (uint8_t*)malloc((size_t)cx->net->in_dim);
    orc_observe_one(cx->env, child, obs);
    orc_net_out o;
    const int lhead = cfg->exploration ? EAZ_HEAD_EXPLORE : EAZ_HEAD_EXPLOIT; /* :132 */
    orc_net_one(cx->net, obs, cx->hash_dim, (1 << EAZ_HEAD_VALUE) | (1 << EAZ_HEAD_UBE) | (1 << lhead), &o);
    free(obs);
    const float* lg = cfg->exploration ? o.explore : o.exploit;
    const float m = orc_maxv(lg, A);
    for (int a = 0; a < A; ++a) logits[a] = eaz_sub(lg[a], m); /* :135; legal mask all True :137 */
    *value = child->terminated ? 0.0f : o.value;                /* :140 */
    *var = child->terminated ? 0.0f : o.ube;                    /* :141 */
  }
  float d = cfg->discount;
  if (cfg->two_players_game) d = eaz_mul(d, -1.0f); /* :142-143 */
  *discount = child->terminated ? 0.0f : d;         /* :144 */
}

/* epistemic_qtransform_completed_by_mix_value (mctx qtransforms.py + beta) */
static void orc_qtransform(const eaz_search_config* cfg, const orc_tree* t, int node, float beta, int use_beta,
                           float* out) {
  const int A = t->A;
  const size_t o = (size_t)node * A;
  float q[ORC_MAX_A], p[ORC_MAX_A], tmp[ORC_MAX_A] = {0};
  int32_t sumN = 0, maxN = 0;
  for (int a = 0; a < A; ++a) {
    const float d = t->discounts[o + a];
    q[a] = eaz_add(t->rewards[o + a], eaz_mul(d, t->values[o + a]));
    if (use_beta) {
      const float qv = eaz_add(t->rewards_var[o + a], eaz_mul(eaz_mul(d, d), t->values_var[o + a]));
      q[a] = eaz_add(q[a], eaz_mul(beta, eaz_sqrt(qv)));
    }
    const int32_t v = t->children_visits[o + a];
    sumN += v;
    if (v > maxN) maxN = v;
  }
  float raw = t->raw_values[node];
  if (use_beta && (cfg->flags & EAZ_FLAG_BETA_RAW)) raw = eaz_add(raw, eaz_mul(beta, eaz_sqrt(t->raw_var[node])));
  float value = raw;
  if (cfg->use_mixed_value) { /* _compute_mixed_value */
    orc_softmax(t->prior_logits + o, A, p);
    for (int a = 0; a < A; ++a) p[a] = eaz_max(EAZ_F32_TINY, p[a]);
    for (int a = 0; a < A; ++a) tmp[a] = t->children_visits[o + a] > 0 ? p[a] : 0.0f;
    const float sumP = orc_tree_sum(tmp, A);
    for (int a = 0; a < A; ++a)
      tmp[a] = t->children_visits[o + a] > 0 ? eaz_div(eaz_mul(p[a], q[a]), sumP) : 0.0f;
    const float wq = orc_tree_sum(tmp, A);
    value = eaz_div(eaz_add(raw, eaz_mul((float)sumN, wq)), (float)(sumN + 1));
  }
  for (int a = 0; a < A; ++a) tmp[a] = t->children_visits[o + a] > 0 ? q[a] : value; /* _complete_qvalues, reanalyze.py:32-40 */
  if (cfg->rescale_values) { /* _rescale_qvalues */
    const float mn = orc_minv(tmp, A), mx = orc_maxv(tmp, A);
    const float den = eaz_max(eaz_sub(mx, mn), cfg->epsilon);
    for (int a = 0; a < A; ++a) tmp[a] = eaz_div(eaz_sub(tmp[a], mn), den);
  }
  const float scale = eaz_mul(eaz_add(cfg->maxvisit_init, (float)maxN), cfg->value_scale);
  for (int a = 0; a < A; ++a) out[a] = eaz_mul(scale, tmp[a]);
}

/* seq_halving.score_considered + masked_argmax (mctx) */
static int orc_root_argmax(const orc_tree* t, const float* gumbel, const float* cq, const uint8_t* invalid,
                           int considered_visit) {
  const int A = t->A;
  float score[ORC_MAX_A];
  const float m = orc_maxv(t->prior_logits, A);
  for (int a = 0; a < A; ++a) {
    const float lg = eaz_sub(t->prior_logits[a], m);
    float s = eaz_max(-1e9f, eaz_add(eaz_add(gumbel[a], lg), cq[a]));
    s = eaz_add(s, t->children_visits[a] == considered_visit ? 0.0f : ORC_NEG_INF);
    score[a] = (invalid && invalid[a]) ? ORC_NEG_INF : s;
  }
  return orc_argmax(score, A);
}

/* ---- emctx.epistemic_muzero_policy (EAZ_FLAG_PUCT): mctx action_selection.muzero_action_selection with
 * qtransforms.qtransform_by_parent_and_siblings; the beta bonus enters q as in the Gumbel path (assumption). */

/* tie-break noise: mctx adds 1e-7 * uniform(rng_key, [A]); here a counter-based stream keyed by (seed, tree, node, visits of
 * the node, action) -- "the k-th selection made at this node" -- so that lazily cached and eagerly recomputed selections agree */
static uint32_t orc_xx_round(uint32_t acc, uint32_t w) { return rotl32(acc + w * 0x85EBCA77u, 13) * 0x9E3779B1u; }
static float orc_tie_noise(uint32_t seed, uint32_t b, uint32_t node, uint32_t visits, uint32_t a) {
  uint32_t h = seed + 0x9E3779B1u;
  h = orc_xx_round(h, b);
  h = orc_xx_round(h, node);
  h = orc_xx_round(h, visits);
  h = orc_xx_round(h, a);
  h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13; h *= 0xC2B2AE3Du; h ^= h >> 16;
  return eaz_mul((float)(h >> 8), 5.9604644775390625e-08f); /* 24 random bits * 2^-24: exact */
}

static int orc_puct_select(const eaz_search_config* cfg, const orc_tree* t, int node, int b, float beta, int use_beta,
                           const uint8_t* invalid_at_root) {
  const int A = t->A;
  const size_t o = (size_t)node * A;
  float q[ORC_MAX_A], safe[ORC_MAX_A] = {0}, p[ORC_MAX_A], score[ORC_MAX_A];
  float node_value = t->node_values[node];
  if (use_beta && (cfg->flags & EAZ_FLAG_BETA_RAW)) node_value = eaz_add(node_value, eaz_mul(beta, eaz_sqrt(t->node_var[node])));
  for (int a = 0; a < A; ++a) { /* tree.qvalues(node) */
    const float d = t->discounts[o + a];
    q[a] = eaz_add(t->rewards[o + a], eaz_mul(d, t->values[o + a]));
    if (use_beta) {
      const float qv = eaz_add(t->rewards_var[o + a], eaz_mul(eaz_mul(d, d), t->values_var[o + a]));
      q[a] = eaz_add(q[a], eaz_mul(beta, eaz_sqrt(qv)));
    }
    safe[a] = t->children_visits[o + a] > 0 ? q[a] : node_value;
  }
  const float mn = eaz_min(node_value, orc_minv(safe, A)), mx = eaz_max(node_value, orc_maxv(safe, A));
  const float den = eaz_max(eaz_sub(mx, mn), cfg->epsilon);
  const float nv = (float)t->node_visits[node];
  const float pb_c = eaz_add(cfg->pb_c_init, eaz_log(eaz_div(eaz_add(eaz_add(nv, cfg->pb_c_base), 1.0f), cfg->pb_c_base)));
  const float explore = eaz_mul(eaz_sqrt(nv), pb_c);
  orc_softmax(t->prior_logits + o, A, p);
  for (int a = 0; a < A; ++a) {
    const int32_t vc = t->children_visits[o + a];
    const float value_score = eaz_div(eaz_sub(vc > 0 ? q[a] : mn, mn), den);
    const float policy_score = eaz_div(eaz_mul(explore, p[a]), eaz_add((float)vc, 1.0f));
    const float noise = eaz_mul(1e-7f, orc_tie_noise(cfg->noise_seed, (uint32_t)b, (uint32_t)node, (uint32_t)t->node_visits[node], (uint32_t)a));
    score[a] = eaz_add(eaz_add(value_score, policy_score), noise);
    if (invalid_at_root && invalid_at_root[a]) score[a] = ORC_NEG_INF; /* masked_argmax(to_argmax, root_invalid_actions * (depth == 0)) */
  }
  return orc_argmax(score, A);
}

static void orc_update_node(orc_tree* t, int node, const float* logits, float value, float var,
                            const orc_env_state* emb) {
  memcpy(t->prior_logits + (size_t)node * t->A, logits, sizeof(float) * (size_t)t->A);
  t->raw_values[node] = value;
  t->node_values[node] = value;
  t->raw_var[node] = var;
  t->node_var[node] = var;
  t->node_visits[node] += 1;
  t->emb[node] = *emb;
}

static void orc_search_one(orc_ctx* cx, const eaz_search_inputs* in, eaz_search_outputs* out, const int32_t* table) {
  const eaz_search_config* cfg = cx->cfg;
  const int b = cx->b, n = cfg->num_simulations, N = n + 1, A = orc_env_num_actions(cx->env);
  const int max_depth = cfg->max_depth > 0 ? cfg->max_depth : n;
  const float beta = in->beta ? in->beta[b] : 0.0f;
  const uint8_t* invalid = in->invalid_actions ? in->invalid_actions + (size_t)b * A : NULL;

  orc_tree t;
  t.N = N; t.A = A;
  t.node_visits = (int32_t*)calloc((size_t)N, 4);
  t.parents = (int32_t*)malloc((size_t)N * 4);
  t.action_from_parent = (int32_t*)malloc((size_t)N * 4);
  t.raw_values = (float*)calloc((size_t)N, 4);
  t.node_values = (float*)calloc((size_t)N, 4);
  t.raw_var = (float*)calloc((size_t)N, 4);
  t.node_var = (float*)calloc((size_t)N, 4);
  t.children_index = (int32_t*)malloc((size_t)N * A * 4);
  t.children_visits = (int32_t*)calloc((size_t)N * A, 4);
  t.prior_logits = (float*)calloc((size_t)N * A, 4);
  t.rewards = (float*)calloc((size_t)N * A, 4);
  t.discounts = (float*)calloc((size_t)N * A, 4);
  t.values = (float*)calloc((size_t)N * A, 4);
  t.rewards_var = (float*)calloc((size_t)N * A, 4);
  t.values_var = (float*)calloc((size_t)N * A, 4);
  t.emb = (orc_env_state*)calloc((size_t)N, sizeof(orc_env_state));
  for (int i = 0; i < N; ++i) t.parents[i] = t.action_from_parent[i] = -1;
  for (int i = 0; i < N * A; ++i) t.children_index[i] = -1;

  /* policy wrapper step 1-2 (A.1): mask root logits, scale gumbel */
  float root_logits[ORC_MAX_A], gumbel[ORC_MAX_A], cq[ORC_MAX_A], x[ORC_MAX_A], p[ORC_MAX_A];
  orc_mask_invalid(in->prior_logits + (size_t)b * A, invalid, A, root_logits);
  for (int a = 0; a < A; ++a) gumbel[a] = eaz_mul(cfg->gumbel_scale, in->gumbel[(size_t)b * A + a]);
  orc_env_state root_state;
  orc_load(cx->env, in->embedding, b, &root_state);
  orc_update_node(&t, 0, root_logits, in->value[b], in->value_epistemic_variance[b], &root_state); /* A.2 */

  int num_valid = 0;
  for (int a = 0; a < A; ++a) num_valid += !(invalid && invalid[a]);
  const int num_considered = num_valid < cfg->max_num_considered_actions ? num_valid : cfg->max_num_considered_actions;

  float logits[ORC_MAX_A];
  for (int sim = 0; sim < n; ++sim) {
    /* simulate (A.3) */
    int node = 0, action = -1, depth = 0, next = 0, cont = 1;
    while (cont) {
      node = next;
      if (cfg->flags & EAZ_FLAG_PUCT) { /* muzero_action_selection at every depth */
        action = orc_puct_select(cfg, &t, node, b, beta, depth == 0 || (cfg->flags & EAZ_FLAG_BETA_INTERIOR) != 0, depth == 0 ? invalid : NULL);
      } else if (depth == 0) { /* gumbel_muzero_root_action_selection */
        orc_qtransform(cfg, &t, 0, beta, 1, cq);
        int sim_index = 0;
        for (int a = 0; a < A; ++a) sim_index += t.children_visits[a];
        const int considered_visit = table[(size_t)num_considered * n + sim_index];
        action = orc_root_argmax(&t, gumbel, cq, invalid, considered_visit);
      } else { /* gumbel_muzero_interior_action_selection */
        orc_qtransform(cfg, &t, node, beta, (cfg->flags & EAZ_FLAG_BETA_INTERIOR) != 0, cq);
        const size_t o = (size_t)node * A;
        int32_t sumN = 0;
        for (int a = 0; a < A; ++a) { x[a] = eaz_add(t.prior_logits[o + a], cq[a]); sumN += t.children_visits[o + a]; }
        orc_softmax(x, A, p);
        const float den = (float)(1 + sumN);
        for (int a = 0; a < A; ++a) x[a] = eaz_sub(p[a], eaz_div((float)t.children_visits[o + a], den));
        action = orc_argmax(x, A);
      }
      next = t.children_index[(size_t)node * A + action];
      depth += 1;
      cont = (next != -1) && (depth < max_depth);
    }
    const int parent = node;
    int leaf = t.children_index[(size_t)parent * A + action];
    if (leaf == -1) leaf = sim + 1;
    /* expand (A.4) */
    float value, var, reward, discount;
    orc_env_state child;
    orc_recurrent(cx, action, &t.emb[parent], &child, logits, &value, &var, &reward, &discount);
    orc_update_node(&t, leaf, logits, value, var, &child);
    const size_t e = (size_t)parent * A + action;
    t.children_index[e] = leaf;
    t.rewards[e] = reward;
    t.discounts[e] = discount;
    t.rewards_var[e] = 0.0f; /* context.py:149 */
    t.parents[leaf] = parent;
    t.action_from_parent[leaf] = action;
    /* backward (A.5) */
    float leaf_value = t.node_values[leaf], leaf_var = t.node_var[leaf];
    if (cfg->flags & EAZ_FLAG_BACKUP_STD) leaf_var = eaz_sqrt(leaf_var);
    int index = leaf;
    while (index != 0) {
      const int par = t.parents[index], act = t.action_from_parent[index];
      const size_t pe = (size_t)par * A + act;
      const float count = (float)t.node_visits[par];
      const float d = t.discounts[pe];
      leaf_value = eaz_add(t.rewards[pe], eaz_mul(d, leaf_value));
      const float parent_value = eaz_div(eaz_add(eaz_mul(t.node_values[par], count), leaf_value), eaz_add(count, 1.0f));
      float parent_var;
      if (cfg->flags & EAZ_FLAG_BACKUP_STD) { /* running mean of std, stored squared */
        leaf_var = eaz_add(eaz_sqrt(t.rewards_var[pe]), eaz_mul(d < 0.0f ? -d : d, leaf_var));
        const float ps = eaz_div(eaz_add(eaz_mul(eaz_sqrt(t.node_var[par]), count), leaf_var), eaz_add(count, 1.0f));
        parent_var = eaz_mul(ps, ps);
      } else {
        leaf_var = eaz_add(t.rewards_var[pe], eaz_mul(eaz_mul(d, d), leaf_var));
        parent_var = eaz_div(eaz_add(eaz_mul(t.node_var[par], count), leaf_var), eaz_add(count, 1.0f));
      }
      t.node_values[par] = parent_value;
      t.node_var[par] = parent_var;
      t.node_visits[par] += 1;
      t.values[pe] = t.node_values[index];
      t.values_var[pe] = t.node_var[index];
      t.children_visits[pe] += 1;
      index = par;
    }
  }

  if (cfg->flags & EAZ_FLAG_PUCT) {
    /* mctx policies.muzero_policy: action_weights = visit_probs; action ~ categorical(log(visit_probs) / temperature), drawn as
     * argmax(logits + gumbel) with the supplied noise */
    int32_t tot = 0;
    for (int a = 0; a < A; ++a) tot += t.children_visits[a];
    for (int a = 0; a < A; ++a) {
      p[a] = tot > 0 ? eaz_div((float)t.children_visits[a], eaz_max((float)tot, 1.0f)) : eaz_div(1.0f, (float)A);
      x[a] = eaz_log(eaz_max(p[a], EAZ_F32_TINY)); /* _get_logits_from_probs */
    }
    const float mxl = orc_maxv(x, A);
    const float tdiv = eaz_max(EAZ_F32_TINY, cfg->temperature);
    for (int a = 0; a < A; ++a) x[a] = eaz_add(eaz_div(eaz_sub(x[a], mxl), tdiv), gumbel[a]); /* _apply_temperature + Gumbel-max */
    out->action[b] = orc_argmax(x, A);
    if (out->action_weights) memcpy(out->action_weights + (size_t)b * A, p, sizeof(float) * (size_t)A);
  } else {
  /* policy wrapper step 4 (A.1) */
  int considered_visit = 0;
  for (int a = 0; a < A; ++a) if (t.children_visits[a] > considered_visit) considered_visit = t.children_visits[a];
  orc_qtransform(cfg, &t, 0, beta, (cfg->flags & EAZ_FLAG_BETA_FINAL) != 0, cq);
  out->action[b] = orc_root_argmax(&t, gumbel, cq, invalid, considered_visit);
  for (int a = 0; a < A; ++a) x[a] = eaz_add(root_logits[a], cq[a]);
  orc_mask_invalid(x, invalid, A, logits);
  orc_softmax(logits, A, p);
  if (out->action_weights) memcpy(out->action_weights + (size_t)b * A, p, sizeof(float) * (size_t)A);
  }

  /* epistemic_summary (A.7) */
  if (out->value) out->value[b] = t.node_values[0];
  if (out->value_epistemic_std) out->value_epistemic_std[b] = eaz_sqrt(t.node_var[0]);
  int32_t total = 0;
  for (int a = 0; a < A; ++a) total += t.children_visits[a];
  for (int a = 0; a < A; ++a) {
    const float vc = (float)t.children_visits[a];
    const float d = t.discounts[a];
    if (out->visit_counts) out->visit_counts[(size_t)b * A + a] = vc;
    if (out->visit_probs)
      out->visit_probs[(size_t)b * A + a] =
          total > 0 ? eaz_div(vc, eaz_max((float)total, 1.0f)) : eaz_div(1.0f, (float)A);
    if (out->qvalues) out->qvalues[(size_t)b * A + a] = eaz_add(t.rewards[a], eaz_mul(d, t.values[a]));
    if (out->qvalues_epistemic_variance)
      out->qvalues_epistemic_variance[(size_t)b * A + a] =
          eaz_add(t.rewards_var[a], eaz_mul(eaz_mul(d, d), t.values_var[a]));
  }

  /* optional tree export, emctx layout */
#define ORC_CP(dst, src, cnt) if (out->dst) memcpy(out->dst + (size_t)b * (cnt), t.src, (size_t)(cnt) * 4)
  ORC_CP(node_visits, node_visits, N);
  ORC_CP(raw_values, raw_values, N);
  ORC_CP(node_values, node_values, N);
  ORC_CP(raw_values_epistemic_variance, raw_var, N);
  ORC_CP(node_values_epistemic_variance, node_var, N);
  ORC_CP(parents, parents, N);
  ORC_CP(action_from_parent, action_from_parent, N);
  ORC_CP(children_index, children_index, N * A);
  ORC_CP(children_prior_logits, prior_logits, N * A);
  ORC_CP(children_visits, children_visits, N * A);
  ORC_CP(children_rewards, rewards, N * A);
  ORC_CP(children_discounts, discounts, N * A);
  ORC_CP(children_values, values, N * A);
  ORC_CP(children_rewards_epistemic_variance, rewards_var, N * A);
  ORC_CP(children_values_epistemic_variance, values_var, N * A);
#undef ORC_CP
  if (out->embeddings) {
    const int S = orc_env_compact_bytes(cx->env);
    for (int i = 0; i < N; ++i) {
      uint8_t* dst = out->embeddings + ((size_t)b * N + i) * S;
      if (t.node_visits[i] > 0) orc_compact(cx->env, &t.emb[i], dst);
      else memset(dst, 0, (size_t)S);
    }
  }
  free(t.node_visits); free(t.parents); free(t.action_from_parent); free(t.raw_values); free(t.node_values);
  free(t.raw_var); free(t.node_var); free(t.children_index); free(t.children_visits); free(t.prior_logits);
  free(t.rewards); free(t.discounts); free(t.values); free(t.rewards_var); free(t.values_var); free(t.emb);
}

/* emctx.epistemic_gumbel_muzero_policy for B roots. `replay` may be NULL.
 * Returns 0, or a positive count of replay misses, or <0 on bad arguments. */
int orc_search_gumbel(const eaz_search_config* cfg, const eaz_search_inputs* in, eaz_search_outputs* out,
                      const orc_replay* replay) {
  if (!cfg || !in || !out || cfg->batch < 1 || cfg->num_simulations < 1) return EAZ_ERR_INVALID_ARG;
  if (orc_check_env(in->env)) return EAZ_ERR_INVALID_ARG;
  const int A = orc_env_num_actions(in->env), n = cfg->num_simulations;
  if (A > ORC_MAX_A || cfg->max_num_considered_actions < 1 || cfg->max_num_considered_actions > ORC_MAX_A)
    return EAZ_ERR_INVALID_ARG;
  const int hash_dim = replay ? 0 : orc_env_hash_dim(in->env, in->net->hash_io);
  if (!replay && (hash_dim % 4 != 0 || in->net->in_dim != orc_env_obs_dim(in->env))) return EAZ_ERR_INVALID_ARG;
  int32_t* table = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cfg->max_num_considered_actions + 1) * n);
  orc_seq_halving_table(cfg->max_num_considered_actions, n, table);
  int misses = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : misses)
  for (int b = 0; b < cfg->batch; ++b) {
    orc_ctx cx = {cfg, in->env, in->net, hash_dim, replay, b, 0};
    orc_search_one(&cx, in, out, table);
    misses += cx.replay_miss;
  }
  free(table);
  return misses;
}

/* eaz_math.h probes for the accuracy tests */
/* ------------------------------------------------------------------------ */
/* reanalyze targets: /root/reference/src/reanalyze.py:86-129                  */
int orc_reanalyze_targets(const eaz_reanalyze_config* cfg, int32_t B, int32_t A, const int32_t* action, const float* qvalues,
                          const float* qvar, const float* visit_counts, const float* value, const float* value_std,
                          const float* next_state_value, const float* next_rewards, const uint8_t* next_terminated,
                          const uint8_t* terminated, const uint8_t* invalid, float* value_target, float* ube_target,
                          float* exploration_policy_target) {
  if (!cfg || B < 0 || A < 1 || A > 256) return EAZ_ERR_INVALID_ARG;
  for (int b = 0; b < B; ++b) {
    const float* q = qvalues + (size_t)b * A;
    const float* qv = qvar + (size_t)b * A;
    const float* vc = visit_counts + (size_t)b * A;
    const float from_tree = q[action[b]];                                                                  /* :87 */
    const float not_term_next = next_terminated[b] ? 0.0f : 1.0f;
    const float from_td = eaz_add(next_rewards[b], eaz_mul(eaz_mul(cfg->discount, next_state_value[b]), not_term_next)); /* :94-95 */
    const float vt = eaz_max(from_tree, from_td);                                                           /* :101 jnp.maximum */
    const float ut = cfg->exploration_ube_target ? orc_maxv(qv, A) : qv[action[b]];                         /* :102-106 */
    const float not_term = terminated[b] ? 0.0f : 1.0f;
    value_target[b] = eaz_mul(vt, not_term);                                                                /* :109-110 */
    ube_target[b] = eaz_mul(ut, not_term);
    float sc[256], masked[256];
    const float vfill = eaz_add(value[b], eaz_mul(cfg->exploration_beta, value_std[b]));                    /* :116 */
    for (int a = 0; a < A; ++a) {
      const float qs = eaz_add(q[a], eaz_mul(cfg->exploration_beta, eaz_sqrt(qv[a])));                      /* :114 */
      sc[a] = vc[a] > 0.0f ? qs : vfill;                                                                    /* complete_qs :32-40 */
    }
    orc_mask_invalid(sc, invalid ? invalid + (size_t)b * A : NULL, A, masked);                              /* :16-29 */
    for (int a = 0; a < A; ++a) masked[a] = eaz_mul(masked[a], cfg->exploration_policy_target_temperature); /* :121 */
    orc_softmax(masked, A, exploration_policy_target + (size_t)b * A);                                      /* :120 */
  }
  return 0;
}

float orc_expf(float x) { return eaz_exp(x); }
float orc_tanhf(float x) { return eaz_tanh(x); }
float orc_logf(float x) { return eaz_log(x); }
void orc_softmax_probe(const float* x, int32_t A, float* p) { orc_softmax(x, A, p); }
float orc_tree_sum_probe(const float* x, int32_t A) { return orc_tree_sum(x, A); }
