// Test driver for the XLA-FFI stand-in (tests/xla_ffi_stub/xla/ffi/api/ffi.h): builds a CallFrame from plain C arrays so that
// Python (ctypes) can invoke the handlers of e_alphazero_b200/csrc/xla_ffi_shim.cc without XLA.  TEST INFRASTRUCTURE ONLY.
#include <cstring>
#include <string>

#include "xla/ffi/api/ffi.h"

extern "C" {
struct StubBuf {
  int32_t dtype;  // xla::ffi::DataType
  int32_t ndim;
  void* data;
  int64_t dims[4];
};
struct StubAttr {
  const char* name;
  int32_t is_float;
  int32_t pad;
  int64_t i;
  double f;
};
typedef int (*StubHandler)(const xla::ffi::CallFrame*, std::string*);

int eaz_stub_call(StubHandler handler, void* stream, int nargs, const StubBuf* args, int nrets, const StubBuf* rets, int nattrs,
                  const StubAttr* attrs, char* msg, int msg_len) {
  std::vector<xla::ffi::RawBuffer> store((size_t)nargs + (size_t)nrets);
  xla::ffi::CallFrame frame;
  frame.stream = stream;
  for (int k = 0; k < nargs + nrets; ++k) {
    const StubBuf& b = k < nargs ? args[k] : rets[k - nargs];
    store[k].dtype = (xla::ffi::DataType)b.dtype;
    store[k].data = b.data;
    store[k].dims.assign(b.dims, b.dims + b.ndim);
    (k < nargs ? frame.args : frame.rets).push_back(&store[k]);
  }
  for (int k = 0; k < nattrs; ++k) frame.attrs.push_back({attrs[k].name, attrs[k].is_float != 0, attrs[k].i, attrs[k].f});
  std::string m;
  const int rc = handler(&frame, &m);
  if (msg && msg_len > 0) {
    std::strncpy(msg, m.c_str(), (size_t)msg_len - 1);
    msg[msg_len - 1] = 0;
  }
  return rc;
}
}
