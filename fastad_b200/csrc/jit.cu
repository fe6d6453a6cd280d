// jit.cu — bind-time specialisation of stage kernels with NVRTC.
//
// The reference gets a fused evaluator for every expression from the C++ compiler (expression
// templates).  A g++-built host program has no device compiler, so the generic path
// of expr.cu would be left with the tape interpreter (~8 % of the HBM roofline).  Here the planner's
// stage program is turned into compile-time constants and the SAME kernel source (stage_kernel.cuh,
// embedded in this library at build time) is compiled for sm_100a by NVRTC on first use: loops unroll,
// switches fold, node values sit in registers.  Kernels are cached per device by program text; leaf
// pointers, sizes and seeds are launch arguments, so rebinding never recompiles.
//
// libnvrtc and the driver entry points are resolved at run time (dlopen / cudaGetDriverEntryPoint):
// the library itself links only against cudart and still loads on a machine without a driver.  If
// NVRTC is missing the stage runs on the interpreter — same CUDA path, same results, slower.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "jit.h"
#include "jit_sources.inc"

namespace fadb {

namespace {

struct Nvrtc {
    void* h = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetErrorString) err = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    int major = 0, minor = 0;
    bool ok = false;
};
struct Driver {
    CUresult (*module_load)(CUmodule*, const void*) = nullptr;
    CUresult (*module_unload)(CUmodule) = nullptr;
    CUresult (*module_func)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*occupancy)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*func_attr)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*func_set_attr)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*launch_ex)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;   // optional
    bool ok = false;
};
struct Entry { CUmodule mod = nullptr; CUfunction fn = nullptr; int per_sm = 1; int regs = 0, local_bytes = 0; bool failed = false; };

std::mutex g_mu;
std::atomic<unsigned long long> g_generation{1};      // bumped when every module is unloaded (jit_shutdown)
Nvrtc g_rtc;
Driver g_drv;
int g_state = 0;            // 0 untried, 1 ready, -1 unavailable
int g_enabled = -1;         // -1: from FADB_NO_JIT on first use
std::string g_why, g_log;
std::unordered_map<std::string, Entry> g_cache;
unsigned long long g_compiled = 0, g_hits = 0, g_failed = 0, g_launched = 0;

template <class F> bool sym(void* h, const char* name, F& out) { out = reinterpret_cast<F>(dlsym(h, name)); return out != nullptr; }
template <class F> bool drv(const char* name, F& out)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    out = reinterpret_cast<F>(p);
    return true;
}

bool init_locked()
{
    if (g_state != 0) return g_state > 0;
    g_state = -1;
    const char* env = getenv("FADB_NVRTC_PATH");
    // the toolkit's own NVRTC first: a process that imported torch already holds torch's bundled (older)
    // libnvrtc.so.12 under that name
    const char* names[] = {env, "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                           "libnvrtc.so"};
    for (const char* n : names) {
        if (!n) continue;
        g_rtc.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_rtc.h) break;
    }
    if (!g_rtc.h) { g_why = "libnvrtc not found (set FADB_NVRTC_PATH)"; return false; }
    Nvrtc& r = g_rtc;
    if (!(sym(r.h, "nvrtcCreateProgram", r.create) && sym(r.h, "nvrtcCompileProgram", r.compile) &&
          sym(r.h, "nvrtcGetCUBINSize", r.cubin_size) && sym(r.h, "nvrtcGetCUBIN", r.cubin) &&
          sym(r.h, "nvrtcGetProgramLogSize", r.log_size) && sym(r.h, "nvrtcGetProgramLog", r.log) &&
          sym(r.h, "nvrtcDestroyProgram", r.destroy) && sym(r.h, "nvrtcGetErrorString", r.err))) {
        g_why = "libnvrtc lacks a required entry point";
        return false;
    }
    if (sym(r.h, "nvrtcVersion", r.version)) r.version(&r.major, &r.minor);
    Driver& d = g_drv;
    if (!(drv("cuModuleLoadData", d.module_load) && drv("cuModuleUnload", d.module_unload) &&
          drv("cuModuleGetFunction", d.module_func) && drv("cuLaunchKernel", d.launch) &&
          drv("cuOccupancyMaxActiveBlocksPerMultiprocessor", d.occupancy) && drv("cuFuncGetAttribute", d.func_attr) &&
          drv("cuFuncSetAttribute", d.func_set_attr))) {
        g_why = "driver entry points unavailable";
        return false;
    }
    drv("cuLaunchKernelEx", d.launch_ex);         // optional: programmatic dependent launch of the specialised kernels
    g_state = 1;
    return true;
}

// cubin -> module -> kernel handle + its occupancy at the kernel's block size
bool finish_load(Entry& e, const std::vector<char>& bin, int kind, size_t dyn_smem)
{
    if (g_drv.module_load(&e.mod, bin.data()) != CUDA_SUCCESS) return false;
    if (g_drv.module_func(&e.fn, e.mod, kind == 1 ? "fadb_glm_jit" : "fadb_stage_jit") != CUDA_SUCCESS) return false;
    if (dyn_smem > 48 * 1024 &&
        g_drv.func_set_attr(e.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem) != CUDA_SUCCESS) return false;
    if (g_drv.occupancy(&e.per_sm, e.fn, kind == 1 ? 160 : 128, dyn_smem) != CUDA_SUCCESS || e.per_sm < 1) e.per_sm = 1;
    g_drv.func_attr(&e.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, e.fn);
    g_drv.func_attr(&e.local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, e.fn);
    e.failed = false;
    return true;
}

}  // namespace

bool jit_enabled()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_enabled < 0) { const char* e = getenv("FADB_NO_JIT"); g_enabled = (e && e[0] == '1') ? 0 : 1; }
    return g_enabled == 1;
}

std::string jit_source(const std::string& program, int kind = 0) { return std::string(kJitPrelude) + program + (kind == 1 ? kGlmKernel : kJitKernel); }

// Returns 0 and a launchable kernel, or a negative code (caller falls back to the interpreter).
int jit_get(const Context& c, const std::string& program, JitKernel* out, int kind, size_t dyn_smem)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!init_locked()) return FADB_ERR_STATE;
    const std::string key = std::to_string(c.device) + "|" + std::to_string(kind) + "|" + program;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        if (it->second.failed) return FADB_ERR_STATE;
        ++g_hits;
        out->fn = it->second.fn; out->per_sm = it->second.per_sm; out->regs = it->second.regs; out->local_bytes = it->second.local_bytes;
        out->gen = g_generation.load();
        return FADB_OK;
    }
    Entry e;
    e.failed = true;
    const std::string src = jit_source(program, kind);
    static const std::string pf = std::string("-DFADB_JIT_PF_DIST=") + (getenv("FADB_JIT_PF_DIST") ? getenv("FADB_JIT_PF_DIST") : "1");
    static const std::string minb = getenv("FADB_JIT_MINB") ? std::string("-DFADB_JIT_MINB=") + getenv("FADB_JIT_MINB") : std::string();
    // optional on-disk cubin cache (FADB_JIT_CACHE_DIR): keyed by a hash of the full source text
    std::string cache_file;
    if (const char* dir = getenv("FADB_JIT_CACHE_DIR")) {
        unsigned long long h = 1469598103934665603ull;                 // FNV-1a 64
        for (unsigned char ch : src + pf + minb) { h ^= ch; h *= 1099511628211ull; }
        char name[64];
        snprintf(name, sizeof name, "/fadb_sm100a_%016llx.cubin", h);
        cache_file = std::string(dir) + name;
        if (FILE* f = fopen(cache_file.c_str(), "rb")) {
            std::vector<char> bin;
            char buf[65536];
            size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) bin.insert(bin.end(), buf, buf + n);
            fclose(f);
            if (!bin.empty() && finish_load(e, bin, kind, dyn_smem)) {
                g_cache.emplace(key, e);
                ++g_hits;
                out->fn = e.fn; out->per_sm = e.per_sm; out->regs = e.regs; out->local_bytes = e.local_bytes; out->gen = g_generation.load();
                return FADB_OK;
            }
        }
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = g_rtc.create(&prog, src.c_str(), "fadb_stage_jit.cu", 0, nullptr, nullptr);
    if (r == NVRTC_SUCCESS) {
        std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", "-DFADB_JIT=1", "-diag-suppress=177", pf.c_str()};
        if (g_rtc.major * 100 + g_rtc.minor < 1209) opts.push_back("-DFADB_NO_LD256=1");   // 256-bit ld/st need PTX ISA 8.8
        if (!minb.empty()) opts.push_back(minb.c_str());
        r = g_rtc.compile(prog, (int)opts.size(), opts.data());
        size_t ls = 0;
        if (g_rtc.log_size(prog, &ls) == NVRTC_SUCCESS && ls > 1) { g_log.resize(ls); g_rtc.log(prog, &g_log[0]); }
        if (r == NVRTC_SUCCESS) {
            size_t n = 0;
            std::vector<char> bin;
            if (g_rtc.cubin_size(prog, &n) == NVRTC_SUCCESS && n > 0) {
                bin.resize(n);
                if (g_rtc.cubin(prog, bin.data()) == NVRTC_SUCCESS && finish_load(e, bin, kind, dyn_smem) && !cache_file.empty()) {
                    if (FILE* f = fopen((cache_file + ".tmp").c_str(), "wb")) {
                        const bool okw = fwrite(bin.data(), 1, bin.size(), f) == bin.size();
                        fclose(f);
                        if (okw) rename((cache_file + ".tmp").c_str(), cache_file.c_str());
                    }
                }
            }
        } else {
            g_why = std::string("NVRTC: ") + g_rtc.err(r);
        }
        g_rtc.destroy(&prog);
    }
    g_cache.emplace(key, e);
    if (e.failed) { ++g_failed; return FADB_ERR_STATE; }
    ++g_compiled;
    out->fn = e.fn; out->per_sm = e.per_sm; out->regs = e.regs; out->local_bytes = e.local_bytes; out->gen = g_generation.load();
    return FADB_OK;
}

unsigned long long jit_generation() { return g_generation.load(); }

// fadb_shutdown: unload every specialised module of the current process
void jit_shutdown()
{
    std::lock_guard<std::mutex> lk(g_mu);
    ++g_generation;
    if (g_state == 1)
        for (auto& kv : g_cache)
            if (kv.second.mod) g_drv.module_unload(kv.second.mod);
    g_cache.clear();
}

int jit_launch_params(const JitKernel& k, unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t st, void** params);
int jit_launch(const JitKernel& k, unsigned grid, unsigned block, cudaStream_t st, void* args_struct)
{
    void* params[1] = {args_struct};
    return jit_launch_params(k, grid, block, 0, st, params);
}

int jit_launch_params(const JitKernel& k, unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t st, void** params)
{
    CUresult r;
    if (g_drv.launch_ex && pdl_enabled()) {
        // the kernels start with griddepcontrol.launch_dependents / .wait: they may become resident while the previous kernel on
        // the stream drains, and read nothing before it has completed (same contract as launch_pdl, common.cuh)
        CUlaunchAttribute at[1];
        at[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        at[0].value.programmaticStreamSerializationAllowed = 1;
        CUlaunchConfig cfg = {};
        cfg.gridDimX = grid; cfg.gridDimY = 1; cfg.gridDimZ = 1;
        cfg.blockDimX = block; cfg.blockDimY = 1; cfg.blockDimZ = 1;
        cfg.sharedMemBytes = (unsigned)dyn_smem;
        cfg.hStream = reinterpret_cast<CUstream>(st);
        cfg.attrs = at; cfg.numAttrs = 1;
        r = g_drv.launch_ex(&cfg, reinterpret_cast<CUfunction>(k.fn), params, nullptr);
    } else
        r = g_drv.launch(reinterpret_cast<CUfunction>(k.fn), grid, 1, 1, block, 1, 1, (unsigned)dyn_smem, reinterpret_cast<CUstream>(st), params, nullptr);
    if (r != CUDA_SUCCESS) { set_error("fastad_b200: launch of a specialised kernel failed (CUresult " + std::to_string((int)r) + ")"); return FADB_ERR_CUDA; }
    ++g_launched;
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

extern "C" int fadb_jit_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_enabled < 0) { const char* e = getenv("FADB_NO_JIT"); g_enabled = (e && e[0] == '1') ? 0 : 1; }
    const int prev = g_enabled;
    g_enabled = on ? 1 : 0;
    return prev;
}
extern "C" int fadb_jit_stats(unsigned long long out[4])
{
    if (!out) return FADB_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    out[0] = g_compiled; out[1] = g_hits; out[2] = g_failed; out[3] = g_launched;
    return FADB_OK;
}
extern "C" const char* fadb_jit_diagnostics(void)
{
    static thread_local std::string s;
    std::lock_guard<std::mutex> lk(g_mu);
    s = (g_state < 0 ? "unavailable: " : "") + g_why + " [NVRTC " + std::to_string(g_rtc.major) + "." + std::to_string(g_rtc.minor) + "]" +
        (g_log.empty() ? "" : "\n" + g_log);
    return s.c_str();
}
// Compiles `program` (the text expr.cu generates for one stage and mode) to a cubin WITHOUT a device
// or a driver: the CPU-side build check of the specialisation path (tests, __graft_entry__.build()).
extern "C" int fadb_jit_compile_only(const char* program, size_t* cubin_bytes, char* log, size_t log_cap)
{
    if (!program) return FADB_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_state == 0) init_locked();
    if (!g_rtc.create || !g_rtc.compile) { set_error("fastad_b200: libnvrtc not available"); return FADB_ERR_STATE; }
    const std::string src = jit_source(program, strstr(program, "#define GLM_DOT_LEAF") ? 1 : 0);     // the one-pass kernel's programs carry its defines
    nvrtcProgram prog = nullptr;
    if (g_rtc.create(&prog, src.c_str(), "fadb_stage_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return FADB_ERR_STATE;
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", "-DFADB_JIT=1", "-diag-suppress=177", "--ptxas-options=-v"};
    if (g_rtc.major * 100 + g_rtc.minor < 1209) opts.push_back("-DFADB_NO_LD256=1");
    const nvrtcResult r = g_rtc.compile(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    if (g_rtc.log_size(prog, &ls) == NVRTC_SUCCESS && ls > 1 && log && log_cap) {
        std::string l(ls, '\0');
        g_rtc.log(prog, &l[0]);
        snprintf(log, log_cap, "%s", l.c_str());
    } else if (log && log_cap) log[0] = 0;
    size_t n = 0;
    if (r == NVRTC_SUCCESS) g_rtc.cubin_size(prog, &n);
    if (cubin_bytes) *cubin_bytes = n;
    if (const char* dump = getenv("FADB_JIT_DUMP_CUBIN"); dump && n && g_rtc.cubin) {   // for cuobjdump -sass (instruction mix studies)
        std::string bin(n, '\0');
        if (g_rtc.cubin(prog, &bin[0]) == NVRTC_SUCCESS)
            if (FILE* f = fopen(dump, "wb")) { fwrite(bin.data(), 1, n, f); fclose(f); }
    }
    g_rtc.destroy(&prog);
    if (r != NVRTC_SUCCESS) { set_error(std::string("fastad_b200: NVRTC: ") + g_rtc.err(r)); return FADB_ERR_STATE; }
    return FADB_OK;
}
