"""Builds e_alphazero_b200/csrc/xla_ffi_shim.cc against the XLA-FFI stand-in of this directory (+ driver.cc) and calls its handlers
through ctypes.  TEST INFRASTRUCTURE ONLY: stands in for XLA, which is not installable in this image."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "e_alphazero_b200")
SO = os.path.join(HERE, "build", "libeaz_xla_ffi_stub.so")
HANDLERS = ("EazSearch", "EazEnvStep", "EazEnvInit", "EazMlpForwardStates", "EazReanalyzeTargets", "EazHashUpdate")
DT = dict(pred=1, s32=4, u8=6, f32=11)  # xla::ffi::DataType values of the stand-in header


class StubBuf(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("ndim", C.c_int32), ("data", C.c_void_p), ("dims", C.c_int64 * 4)]


class StubAttr(C.Structure):
    _fields_ = [("name", C.c_char_p), ("is_float", C.c_int32), ("pad", C.c_int32), ("i", C.c_int64), ("f", C.c_double)]


def build():
    srcs = [os.path.join(PKG, "csrc", "xla_ffi_shim.cc"), os.path.join(HERE, "driver.cc")]
    deps = srcs + [os.path.join(HERE, "xla", "ffi", "api", "ffi.h"), os.path.join(ROOT, "include", "eaz_b200.h")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wall", "-Werror", "-I", HERE, "-I", os.path.join(ROOT, "include"),
                    "-I", "/usr/local/cuda/include", *srcs, "-o", SO, "-L", PKG, "-leaz_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                    f"-Wl,-rpath,{PKG}"], check=True)
    return SO


def load():
    from e_alphazero_b200 import _lib

    _lib.load()
    lib = C.CDLL(build())
    lib.eaz_stub_call.restype = C.c_int
    return lib


def buf(kind, data_ptr, *dims):
    b = StubBuf(DT[kind], len(dims), data_ptr, (C.c_int64 * 4)(*dims))
    return b


def tbuf(t):
    """torch tensor -> StubBuf"""
    import torch

    kind = {torch.float32: "f32", torch.int32: "s32", torch.uint8: "u8", torch.bool: "pred"}[t.dtype]
    return buf(kind, t.data_ptr(), *t.shape)


def call(lib, handler, args, rets, attrs, stream=None):
    """attrs: dict name -> int | float (python float = float attribute).  Returns (error code, message)."""
    a = (StubBuf * max(len(args), 1))(*args)
    r = (StubBuf * max(len(rets), 1))(*rets)
    at = (StubAttr * max(len(attrs), 1))(*[StubAttr(k.encode(), int(isinstance(v, float)), 0, 0 if isinstance(v, float) else int(v),
                                                    float(v) if isinstance(v, float) else 0.0) for k, v in attrs.items()])
    msg = C.create_string_buffer(512)
    fn = C.cast(getattr(lib, handler), C.c_void_p)
    rc = lib.eaz_stub_call(fn, C.c_void_p(stream), len(args), a, len(rets), r, len(attrs), at, msg, 512)
    return rc, msg.value.decode()
