"""The SHIPPED Subleq interpreter source (csrc/common.cuh: subleq_simulate with the exact loop shortcut and the tight common-cycle loop)
-- and subleq_simulate16, the register-resident machine used for word size 16 --
compiled for the host and held to the oracle's plain MAX_CYCLE_COUNT loop (subleq.py:156-395) on programs built to spin, count and do
IO.  The function is ordinary C++ apart from two shared-memory byte accessors, which the harness replaces by array accesses; nothing of
the oracle is linked into it.  (The GPU build of the same text is tested in test_gpu_parity.py::test_subleq_cycle_detection_exact.)"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS_HEAD = r'''
#include <cstdint>
#include <cstring>
#include <algorithm>
using std::max; using std::min;
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define EAZ_SUBLEQ_MAX_CYCLES 200
namespace eaz {
static uint8_t g_arena[1024];
inline uint32_t sq_lds8(uint32_t addr) { return g_arena[addr]; }
inline void sq_sts8(uint32_t addr, uint32_t v) { g_arena[addr] = (uint8_t)v; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const uint8_t*)p - g_arena); }
'''
HARNESS_TAIL = r'''
}  // namespace eaz
extern "C" void host_simulate(int ws, const uint8_t* program, int trow, int k, int detect, int* in_after, int* out_after, int* bcc) {
  using namespace eaz;
  uint8_t* mem = g_arena + 64;
  uint8_t* snap = g_arena + 512;
  memset(g_arena, 0xAB, sizeof(g_arena));  // garbage in the padding: the interpreter must never depend on it
  memcpy(mem, program, ws);
  memcpy(snap, program, ws);
  SubleqSim r;
  if (detect >= 2) {  // the register-resident word-size-16 machine
    uint32_t w4[4];
    memcpy(w4, program, 16);
    if (detect == 3) subleq_simulate16<true>(sq_pack_nibbles16(w4), trow, k, r);
    else subleq_simulate16<false>(sq_pack_nibbles16(w4), trow, k, r);
  } else if (detect) subleq_simulate<true>(ws, mem, snap, trow, k, r);
  else subleq_simulate<false>(ws, mem, snap, trow, k, r);
  for (int i = 0; i < 8; ++i) { in_after[i] = r.in[i]; out_after[i] = r.out[i]; }
  bcc[0] = r.bytes_used; bcc[1] = r.cycles; bcc[2] = r.correct;
}
'''


@pytest.fixture(scope="module")
def host_interp():
    src = open(os.path.join(ROOT, "e_alphazero_b200", "csrc", "common.cuh")).read()
    a = src.index("struct SubleqVec {")
    b = src.index("__device__ __forceinline__ float subleq_reward")
    body = src[a:b]
    h0 = body.index("__device__ __forceinline__ uint32_t sq_lds8")
    h1 = body.index("// `mem`, `snap`: shared memory")
    body = body[:h0] + body[h1:]  # drop the two inline-PTX accessors (replaced above)
    d = tempfile.mkdtemp()
    open(os.path.join(d, "h.cpp"), "w").write(HARNESS_HEAD + body + HARNESS_TAIL)
    so = os.path.join(d, "h.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(d, "h.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def _programs(ws, n, rng):
    out = np.zeros((n, ws), np.uint8)
    for i in range(n):
        L = int(rng.integers(0, min(ws - 3, 24)))
        style = i % 4
        if style == 0:
            out[i, :L] = rng.integers(0, ws, L)
        elif style == 1:  # operands and jump targets inside the program
            out[i, :L] = rng.integers(0, max(L, 1), L)
        elif style == 2:  # IO-heavy
            v = rng.integers(0, max(L, 1), L)
            io = rng.random(L) < 0.4
            v[io] = rng.integers(ws - 4, ws, int(io.sum()))
            out[i, :L] = v
        else:  # a counter walking through the residues
            x, y = rng.integers(3, ws - 4, 2)
            out[i, :3] = (x, y, 0)
            out[i, y] = rng.integers(1, ws)
            out[i, x] = rng.integers(0, ws)
    return out


@pytest.mark.parametrize("ws,n", [(16, 12000), (23, 3000), (256, 1500)])
def test_shipped_interpreter_equals_plain_loop(host_interp, ws, n):
    rng = np.random.default_rng(ws)
    progs = _programs(ws, n, rng)
    ia, oa, bcc = (C.c_int * 8)(), (C.c_int * 8)(), (C.c_int * 3)()
    spinners = shortcuts = 0
    for i in range(n):
        task = int(rng.integers(1, 7))
        k = int(rng.integers(0, 3))
        tin, tout = O.subleq_test_cases(task, ws)
        exp = O.subleq_simulate(ws, progs[i].astype(np.int32), tin[k], tout[k])
        for detect in ((1, 0, 3, 2) if ws == 16 else (1, 0)):  # 3 / 2: subleq_simulate16 with / without the loop shortcut
            host_interp.host_simulate(ws, progs[i].ctypes.data_as(C.c_void_p), task - 1, k, detect, ia, oa, bcc)
            got = (list(ia), list(oa), bcc[0], bcc[1], bool(bcc[2]))
            want = (exp["input_after"].tolist(), exp["output_after"].tolist(), exp["bytes_used"], exp["cycles_used"], exp["correct"])
            assert got == want, (ws, i, detect, progs[i].tolist(), task, k, got, want)
        spinners += exp["cycles_used"] == 200
    assert spinners > n // 10  # the population does exercise the cap
