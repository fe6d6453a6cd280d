// jit.h — interface of the NVRTC specialisation layer (jit.cu) used by expr.cu.
#pragma once
#include <string>

namespace fadb {
struct Context;
struct JitKernel { void* fn = nullptr; int per_sm = 1; int regs = 0; int local_bytes = 0; unsigned long long gen = 0; };
// bumped by jit_shutdown (every module is unloaded): a handle of an older generation must be re-fetched, not launched
unsigned long long jit_generation();
bool jit_enabled();
// kind 0: the stage kernel (stage_kernel.cuh, entry fadb_stage_jit, 128 threads); kind 1: the one-pass dot -> map -> sum kernel
// (glm_kernel.cuh, entry fadb_glm_jit, 160 threads, `dyn_smem` bytes of dynamic shared memory)
int jit_get(const Context& c, const std::string& program, JitKernel* out, int kind = 0, size_t dyn_smem = 0);
void jit_shutdown();
int jit_launch(const JitKernel& k, unsigned grid, unsigned block, cudaStream_t st, void* args_struct);
// general form: `params` as cuLaunchKernel takes them, dynamic shared memory
int jit_launch_params(const JitKernel& k, unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t st, void** params);
}  // namespace fadb
