"""ctypes mirror of include/eaz_b200.h (field order and types must match the header).

These are plain POD descriptions; they carry no behaviour.  The product fills
them with CUDA device pointers (``_lib.py``); the CPU oracle's binding reuses the
same layouts with host pointers (``oracle/oracle.py``).
"""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 2

EAZ_OK = 0
EAZ_ERR_INVALID_ARG = -1
EAZ_ERR_WORKSPACE = -2
EAZ_ERR_CUDA = -3
EAZ_ERR_UNSUPPORTED = -4

ENV_DEEPSEA = 0
ENV_SUBLEQ = 1
SUBLEQ_REWARD_SOLVED = 0
SUBLEQ_REWARD_LOWEST_BYTES = 1

HEAD_VALUE, HEAD_UBE, HEAD_EXPLOIT, HEAD_EXPLORE = 0, 1, 2, 3

FLAG_BETA_INTERIOR = 1 << 0
FLAG_BETA_RAW = 1 << 1
FLAG_BETA_FINAL = 1 << 2
FLAG_BACKUP_STD = 1 << 3
FLAG_REUSE_PREPARED = 1 << 4
FLAG_STREAMS_SHIFT = 8
FLAG_PUCT = 1 << 5


def flag_streams(k: int) -> int:
    """EAZ_FLAG_STREAMS(k): search k sub-batches concurrently on auxiliary streams (results unchanged)."""
    return (int(k) & 0xF) << FLAG_STREAMS_SHIFT


SEARCH_DEFAULT_FLAGS = FLAG_BETA_INTERIOR | FLAG_BETA_RAW | FLAG_BETA_FINAL

MLP_EXACT = 0
MLP_TENSOR = 1

_p = C.c_void_p


class EazEnv(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("size", C.c_int32),
        ("action_map", _p),
        ("word_size", C.c_int32),
        ("binary_encoding", C.c_int32),
        ("reward_fn", C.c_int32),
    ]


class EazState(C.Structure):
    _fields_ = [
        ("step_count", _p),
        ("rewards", _p),
        ("terminated", _p),
        ("truncated", _p),
        ("observation", _p),
        ("col", _p),
        ("memory", _p),
        ("task", _p),
        ("solved", _p),
        ("input_after", _p),
        ("output_after", _p),
    ]


class EazFcParams(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int32),
        ("hidden", C.c_int32),
        ("num_actions", C.c_int32),
        ("w", (_p * 3) * 4),
        ("b", (_p * 3) * 4),
        ("binary_set", _p),
        ("hash_bits", C.c_int32),
        ("hash_io", C.c_int32),
        ("word_size", C.c_int32),
        ("max_u", C.c_float),
        ("novelty_scale", C.c_float),
    ]


class EazSearchConfig(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("num_simulations", C.c_int32),
        ("max_depth", C.c_int32),
        ("max_num_considered_actions", C.c_int32),
        ("gumbel_scale", C.c_float),
        ("discount", C.c_float),
        ("two_players_game", C.c_int32),
        ("exploration", C.c_int32),
        ("value_scale", C.c_float),
        ("maxvisit_init", C.c_float),
        ("rescale_values", C.c_int32),
        ("use_mixed_value", C.c_int32),
        ("epsilon", C.c_float),
        ("flags", C.c_int32),
        ("mlp_mode", C.c_int32),
        ("pb_c_init", C.c_float),
        ("pb_c_base", C.c_float),
        ("temperature", C.c_float),
        ("noise_seed", C.c_uint32),
    ]


class EazReanalyzeConfig(C.Structure):
    _fields_ = [
        ("discount", C.c_float),
        ("exploration_beta", C.c_float),
        ("exploration_ube_target", C.c_int32),
        ("exploration_policy_target_temperature", C.c_float),
    ]


class EazSearchInputs(C.Structure):
    _fields_ = [
        ("prior_logits", _p),
        ("value", _p),
        ("value_epistemic_variance", _p),
        ("beta", _p),
        ("embedding", C.POINTER(EazState)),
        ("invalid_actions", _p),
        ("gumbel", _p),
        ("env", C.POINTER(EazEnv)),
        ("net", C.POINTER(EazFcParams)),
    ]


SEARCH_OUTPUT_FIELDS = [
    # name, dtype, shape-kind  (B | BA | BN | BNA | BNS)
    ("action", "i32", "B"),
    ("action_weights", "f32", "BA"),
    ("value", "f32", "B"),
    ("value_epistemic_std", "f32", "B"),
    ("visit_counts", "f32", "BA"),
    ("visit_probs", "f32", "BA"),
    ("qvalues", "f32", "BA"),
    ("qvalues_epistemic_variance", "f32", "BA"),
    ("node_visits", "i32", "BN"),
    ("raw_values", "f32", "BN"),
    ("node_values", "f32", "BN"),
    ("raw_values_epistemic_variance", "f32", "BN"),
    ("node_values_epistemic_variance", "f32", "BN"),
    ("parents", "i32", "BN"),
    ("action_from_parent", "i32", "BN"),
    ("children_index", "i32", "BNA"),
    ("children_prior_logits", "f32", "BNA"),
    ("children_visits", "i32", "BNA"),
    ("children_rewards", "f32", "BNA"),
    ("children_discounts", "f32", "BNA"),
    ("children_values", "f32", "BNA"),
    ("children_rewards_epistemic_variance", "f32", "BNA"),
    ("children_values_epistemic_variance", "f32", "BNA"),
    ("embeddings", "u8", "BNS"),
    ("root_value", "f32", "B"),
    ("root_ube", "f32", "B"),
]
SUMMARY_FIELDS = [f for f in SEARCH_OUTPUT_FIELDS[:8]]
TREE_FIELDS = [f for f in SEARCH_OUTPUT_FIELDS[8:24]]
ROOT_FIELDS = [f for f in SEARCH_OUTPUT_FIELDS[24:]]  # fused-root mode


class EazSearchOutputs(C.Structure):
    _fields_ = [(name, _p) for name, _, _ in SEARCH_OUTPUT_FIELDS]


def default_search_config(**kw) -> EazSearchConfig:
    """Defaults of emctx.epistemic_gumbel_muzero_policy /
    epistemic_qtransform_completed_by_mix_value."""
    cfg = EazSearchConfig(
        batch=0,
        num_simulations=32,
        max_depth=0,
        max_num_considered_actions=16,
        gumbel_scale=1.0,
        discount=0.997,
        two_players_game=0,
        exploration=0,
        value_scale=0.1,
        maxvisit_init=50.0,
        rescale_values=1,
        use_mixed_value=1,
        epsilon=1e-8,
        flags=SEARCH_DEFAULT_FLAGS,
        mlp_mode=MLP_EXACT,
        pb_c_init=1.25,
        pb_c_base=19652.0,
        temperature=1.0,
        noise_seed=0,
    )
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown search config field {k!r}")
        setattr(cfg, k, v)
    return cfg
