// jit.h — interface of the NVRTC specialisation layer (jit.cu) used by expr.cu.
#pragma once
#include <string>

namespace fadb {
struct Context;
struct JitKernel { void* fn = nullptr; int per_sm = 1; int regs = 0; int local_bytes = 0; unsigned long long gen = 0; };
// bumped by jit_shutdown (every module is unloaded): a handle of an older generation must be re-fetched, not launched
unsigned long long jit_generation();
bool jit_enabled();
int jit_get(const Context& c, const std::string& program, JitKernel* out);
void jit_shutdown();
int jit_launch(const JitKernel& k, unsigned grid, unsigned block, cudaStream_t st, void* args_struct);
}  // namespace fadb
