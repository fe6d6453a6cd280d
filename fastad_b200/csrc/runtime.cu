// runtime.cu — lifecycle, error state, device memory and stream plumbing of libfastad_b200.
// Replaces the host-heap storage side of the reference path: Var's owned value/adjoint
// (reverse/core/var.hpp:23-220), ExprBind's caches (bind.hpp:38-40), PtrPack
// (util/ptr_pack.hpp:13-25).  There is deliberately no CPU fallback anywhere.
#include "common.cuh"
#include "jit.h"

namespace fadb {

static thread_local std::string g_last_error;
static Context g_ctx[kMaxDevices];

typedef void (*ReleaseFn)(int);
static ReleaseFn* shutdown_hooks() { static ReleaseFn hooks[16]; return hooks; }
static int& shutdown_hook_count() { static int n = 0; return n; }
void on_shutdown(void (*release)(int device))
{
    if (shutdown_hook_count() < 16) shutdown_hooks()[shutdown_hook_count()++] = release;
}

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "fastad_b200: CUDA error %d (%s) at %s:%d in `%s`", (int)e,
             cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    return FADB_ERR_CUDA;
}

static int init_context(Context& c, int device)
{
    FADB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FADB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("fastad_b200: this library is built for sm_100a (B200) only; device is sm_" +
                  std::to_string(prop.major) + std::to_string(prop.minor));
        return FADB_ERR_UNSUPPORTED;
    }
    c.device = device;
    c.sms = prop.multiProcessorCount;
    FADB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.own_stream = true;
    FADB_CUDA(cudaMalloc(&c.partials, sizeof(double) * kMaxGrid * kPartialSlots));
    FADB_CUDA(cudaMalloc(&c.tickets, sizeof(unsigned int) * 64));
    FADB_CUDA(cudaMemset(c.tickets, 0, sizeof(unsigned int) * 64));
    FADB_CUDA(cudaMalloc(&c.scalars, sizeof(double) * kScalarSlots));
    FADB_CUDA(cudaMemset(c.scalars, 0, sizeof(double) * kScalarSlots));
    FADB_CUDA(cudaMallocHost(&c.host_scalars, sizeof(double) * kScalarSlots));
    c.launches = 0;
    c.ready = true;
    return FADB_OK;
}

int current_context(Context** out)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("fastad_b200: device index out of range");
        return FADB_ERR_ARG;
    }
    Context& c = g_ctx[dev];
    if (!c.ready) {
        int rc = init_context(c, dev);
        if (rc != FADB_OK) return rc;
    }
    *out = &c;
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

extern "C" {

int fadb_init(int device)
{
    int n = 0;
    FADB_CUDA(cudaGetDeviceCount(&n));
    FADB_REQUIRE(device >= 0 && device < n && device < kMaxDevices, "fadb_init: bad device index");
    FADB_CUDA(cudaSetDevice(device));
    Context* c;
    return current_context(&c);
}

void fadb_shutdown(void)
{
    jit_shutdown();           // specialised stage kernels (jit.cu)
    for (int d = 0; d < kMaxDevices; ++d) {
        Context& c = g_ctx[d];
        if (!c.ready) continue;
        cudaSetDevice(d);
        cudaStreamSynchronize(c.stream);
        for (int h = 0; h < shutdown_hook_count(); ++h) shutdown_hooks()[h](d);     // the kernels' own workspaces
        if (c.own_stream) cudaStreamDestroy(c.stream);
        cudaFree(c.partials);
        cudaFree(c.tickets);
        cudaFree(c.scalars);
        cudaFreeHost(c.host_scalars);
        c = Context();
    }
}

int fadb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int fadb_get_device(void)
{
    int dev = -1;
    FADB_CUDA(cudaGetDevice(&dev));
    return dev;
}

int fadb_set_stream(void* cuda_stream)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    if (cuda_stream == nullptr) {
        if (!c->own_stream) {
            FADB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            c->own_stream = true;
        }
        return FADB_OK;
    }
    if (c->own_stream) {
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        FADB_CUDA(cudaStreamDestroy(c->stream));
        c->own_stream = false;
    }
    c->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return FADB_OK;
}

int fadb_get_stream(void** out_cuda_stream)
{
    FADB_REQUIRE(out_cuda_stream != nullptr, "fadb_get_stream: null out pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    *out_cuda_stream = reinterpret_cast<void*>(c->stream);
    return FADB_OK;
}

int fadb_sync(void)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    return FADB_OK;
}

const char* fadb_last_error(void) { return g_last_error.c_str(); }

unsigned long long fadb_launch_count(void)
{
    Context* c;
    if (current_context(&c)) return 0;
    return c->launches;
}

const char* fadb_version(void) { return "fastad_b200 0.1 (sm_100a)"; }

int fadb_malloc(size_t nbytes, void** out_dev)
{
    FADB_REQUIRE(out_dev != nullptr, "fadb_malloc: null out pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    *out_dev = nullptr;
    if (nbytes == 0) return FADB_OK;
    FADB_CUDA(cudaMalloc(out_dev, nbytes));
    return FADB_OK;
}

int fadb_free(void* dev)
{
    if (dev == nullptr) return FADB_OK;
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    FADB_CUDA(cudaFree(dev));
    return FADB_OK;
}

int fadb_h2d(void* dev_dst, const void* host_src, size_t nbytes)
{
    if (nbytes == 0) return FADB_OK;
    FADB_REQUIRE(dev_dst && host_src, "fadb_h2d: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaMemcpyAsync(dev_dst, host_src, nbytes, cudaMemcpyHostToDevice, c->stream));
    FADB_CUDA(cudaStreamSynchronize(c->stream));  // host_src may be pageable / reused
    return FADB_OK;
}

int fadb_d2h(void* host_dst, const void* dev_src, size_t nbytes)
{
    if (nbytes == 0) return FADB_OK;
    FADB_REQUIRE(host_dst && dev_src, "fadb_d2h: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaMemcpyAsync(host_dst, dev_src, nbytes, cudaMemcpyDeviceToHost, c->stream));
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    return FADB_OK;
}

int fadb_memset_zero(void* dev, size_t nbytes)
{
    if (nbytes == 0) return FADB_OK;
    FADB_REQUIRE(dev != nullptr, "fadb_memset_zero: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaMemsetAsync(dev, 0, nbytes, c->stream));
    return FADB_OK;
}

int fadb_host_alloc_pinned(size_t nbytes, void** out_host)
{
    FADB_REQUIRE(out_host != nullptr, "fadb_host_alloc_pinned: null out pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaMallocHost(out_host, nbytes ? nbytes : 1));
    return FADB_OK;
}

int fadb_host_free_pinned(void* host)
{
    if (!host) return FADB_OK;
    FADB_CUDA(cudaFreeHost(host));
    return FADB_OK;
}

// Page-lock and map an EXISTING host array (a std::vector, an Eigen matrix): the host entry points then take
// their one-launch path over PCIe instead of staging through bounce buffers.  The caller keeps ownership
// and must call fadb_host_unregister before freeing the memory.
int fadb_host_register(void* host, size_t nbytes)
{
    FADB_REQUIRE(host != nullptr && nbytes > 0, "fadb_host_register: null pointer or empty range");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    FADB_CUDA(cudaHostRegister(host, nbytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return FADB_OK;
}

int fadb_host_unregister(void* host)
{
    if (!host) return FADB_OK;
    FADB_CUDA(cudaHostUnregister(host));
    return FADB_OK;
}

}  // extern "C"
