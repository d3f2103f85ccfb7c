"""CPU-side checks of the boundary: the CUDA library builds for sm_100a, loads without a GPU and
exports every symbol include/eaz_b200.h declares; ctypes mirrors match the header; product code never
imports the oracle."""
import ctypes as C
import os
import re

import pytest

from e_alphazero_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "eaz_b200.h")).read()
    declared = set(re.findall(r"\b(eaz_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.eaz_abi_version() == _abi.ABI_VERSION


def test_struct_sizes_match_header():
    # pointer/int layouts: compile a tiny C program printing sizeof/offsetof and compare with ctypes
    import subprocess, tempfile, textwrap

    src = textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "eaz_b200.h"
        int main(void) {
          printf("%zu %zu %zu %zu %zu %zu ", sizeof(eaz_env), sizeof(eaz_state), sizeof(eaz_fc_params), sizeof(eaz_search_config),
                 sizeof(eaz_search_inputs), sizeof(eaz_search_outputs));
          printf("%zu %zu ", sizeof(eaz_reanalyze_config), offsetof(eaz_reanalyze_config, exploration_policy_target_temperature));
          printf("%zu %zu %zu %zu\\n", offsetof(eaz_fc_params, binary_set), offsetof(eaz_fc_params, novelty_scale),
                 offsetof(eaz_search_config, mlp_mode), offsetof(eaz_search_outputs, embeddings));
          return 0;
        }""")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        got = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(_abi.EazEnv), C.sizeof(_abi.EazState), C.sizeof(_abi.EazFcParams), C.sizeof(_abi.EazSearchConfig),
           C.sizeof(_abi.EazSearchInputs), C.sizeof(_abi.EazSearchOutputs), C.sizeof(_abi.EazReanalyzeConfig),
           _abi.EazReanalyzeConfig.exploration_policy_target_temperature.offset, _abi.EazFcParams.binary_set.offset,
           _abi.EazFcParams.novelty_scale.offset, _abi.EazSearchConfig.mlp_mode.offset, _abi.EazSearchOutputs.embeddings.offset]
    assert got == exp


def test_argument_validation_without_gpu(lib):
    # host-side checks mirror the reference asserts and never touch the device
    env = _abi.EazEnv(_abi.ENV_SUBLEQ, 0, None, 8, 1, 0)  # word_size < 16 (subleq.py:606)
    assert lib.eaz_env_num_actions(C.byref(env)) == -1
    assert b"16 <= word_size <= 256" in lib.eaz_last_error()
    env = _abi.EazEnv(_abi.ENV_SUBLEQ, 0, None, 16, 1, 0)
    assert lib.eaz_env_num_actions(C.byref(env)) == 16
    assert lib.eaz_env_obs_dim(C.byref(env)) == 48 * 5 and lib.eaz_env_hash_dim(C.byref(env), 1) == 160
    assert lib.eaz_env_compact_bytes(C.byref(env)) == 56
    env = _abi.EazEnv(_abi.ENV_DEEPSEA, 30, None, 0, 0, 0)
    assert lib.eaz_env_obs_dim(C.byref(env)) == 900 and lib.eaz_env_compact_bytes(C.byref(env)) == 4
    assert lib.eaz_xxhash_indices(None, 1, 25, 24, None, None) == _abi.EAZ_ERR_INVALID_ARG  # hashes.py:210
    cfg = _abi.default_search_config(batch=4096, num_simulations=64)
    assert lib.eaz_search_workspace_bytes(C.byref(cfg), C.byref(env)) > 4096 * 65 * (7 * 2 * 4 + 7 * 4)
    assert lib.eaz_search_num_launches(C.byref(cfg), C.byref(env)) == 1 + 4 + 2 * 64 + 2


def test_ops_fail_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from e_alphazero_b200 import ops

    with pytest.raises(_lib.EazError):
        ops.env_init(ops.subleq_spec(16), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "e_alphazero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert not re.search(r"#\s*include[^\n]*oracle", text), f
                assert "libeaz_oracle" not in text, f


def test_xla_ffi_shim_builds_and_validates(lib):
    """csrc/xla_ffi_shim.cc (the JAX-side binding) compiles with -Wall -Werror against the XLA-FFI stand-in, exports one
    handler per hot-path entry point, and its argument checks answer with ffi::Error before anything touches a device."""
    from tests.xla_ffi_stub import harness as X

    shim = X.load()
    for h in X.HANDLERS:
        assert hasattr(shim, h), h
    env_attrs = dict(env_kind=_abi.ENV_DEEPSEA, size=10, word_size=0, binary_encoding=0, reward_fn=0)
    one = X.buf("s32", None, 1)
    # wrong number of state leaves
    rc, msg = X.call(shim, "EazEnvStep", [X.buf("s32", None, 4), one, X.buf("u8", None, 10, 10)], [], dict(env_attrs, auto_reset=0))
    assert rc == 3 and "number of state leaves" in msg  # kInvalidArgument
    # a missing attribute / a dtype mismatch are rejected by the binding itself
    rc, msg = X.call(shim, "EazEnvStep", [X.buf("s32", None, 4), one, X.buf("u8", None, 10, 10)], [], env_attrs)
    assert rc == 3 and "auto_reset" in msg
    rc, msg = X.call(shim, "EazHashUpdate", [X.buf("s32", None, 4, 8), X.buf("u8", None, 16)], [X.buf("u8", None, 16)], dict(bits=24))
    assert rc == 3 and "dtype" in msg
    # library-side validation travels back as the error message (hashes.py:210: D % 4 == 0)
    rc, msg = X.call(shim, "EazHashUpdate", [X.buf("f32", None, 4, 25), X.buf("u8", None, 16)], [X.buf("u8", None, 16)], dict(bits=24))
    assert rc == 3 and "eaz_hash_update" in msg
    # Subleq word size outside 16..256 (subleq.py:606) through the search handler
    f = lambda *d: X.buf("f32", None, *d)
    args = [f(4), f(4, 8), X.buf("pred", None, 1), X.buf("u8", None, 1), X.buf("u8", None, 1 << 21), f(4, 8), f(4), f(4), X.buf("u8", None, 1)]
    rets = [X.buf("s32", None, 4), f(4, 8), f(4), f(4), f(4, 8), f(4, 8), f(4, 8), f(4, 8), f(4), f(4), X.buf("u8", None, 1)]
    attrs = dict(env_kind=_abi.ENV_SUBLEQ, size=0, word_size=8, binary_encoding=1, reward_fn=0, num_simulations=4, max_depth=0,
                 max_num_considered_actions=16, gumbel_scale=1.0, discount=0.97, two_players_game=0, exploration=0, value_scale=0.1,
                 maxvisit_init=50.0, rescale_values=1, flags=7, mlp_mode=0, fused_root=0, draw_gumbel=0, noise_seed=0, reuse_prepared=0,
                 hash_bits=24, hash_io=1, max_u=1.0, novelty_scale=1.0)
    rc, msg = X.call(shim, "EazSearch", args, rets, attrs)
    assert rc == 3 and "word_size" in msg


def test_alloc_arena_layout_on_cpu():
    """ops.alloc_arena: every tensor is a typed view of ONE flat allocation at 256-byte aligned offsets, and an arena built from the
    same specs anywhere (e.g. pinned host memory) has the same layout -- one copy of `flat` moves all of them."""
    import torch

    from e_alphazero_b200 import ops

    specs = [("a", (5,), torch.int32), ("b", (5, 3), torch.float32), ("c", (7,), torch.uint8), ("d", (2, 2, 2), torch.float32)]
    flat, views = ops.alloc_arena(specs, device="cpu")
    flat2, views2 = ops.alloc_arena(specs, device="cpu")
    assert flat.dtype == torch.uint8 and flat.numel() == flat2.numel() == 4 * 256
    base = flat.data_ptr()
    offs = [views[n].data_ptr() - base for n, _, _ in specs]
    assert offs == [0, 256, 512, 768] and all(o % 256 == 0 for o in offs)
    for name, shape, dt in specs:
        assert tuple(views[name].shape) == shape and views[name].dtype == dt and views[name].is_contiguous()
    views["b"].copy_(torch.arange(15, dtype=torch.float32).reshape(5, 3))
    views["c"].fill_(7)
    flat2.copy_(flat)  # ONE copy moves every tensor
    assert torch.equal(views2["b"], views["b"]) and torch.equal(views2["c"], views["c"]) and int(views2["a"].abs().sum()) == 0
    # the search-output arena of a plan is built from specs that name every summary / root field once
    sp = ops.search_output_specs(B=4, N=9, A=2, S=4, want_tree=True)
    names = [n for n, _, _ in sp]
    assert len(names) == len(set(names)) and {"action", "value", "root_value", "children_prior_logits"} <= set(names)
