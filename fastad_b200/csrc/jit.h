// jit.h — interface of the NVRTC specialisation layer (jit.cu) used by expr.cu.
#pragma once
#include <string>

namespace fadb {
struct Context;
struct JitKernel { void* fn = nullptr; int per_sm = 1; int regs = 0; int local_bytes = 0; };
bool jit_enabled();
int jit_get(const Context& c, const std::string& program, JitKernel* out);
void jit_shutdown();
int jit_launch(const JitKernel& k, unsigned grid, unsigned block, cudaStream_t st, void* args_struct);
}  // namespace fadb
