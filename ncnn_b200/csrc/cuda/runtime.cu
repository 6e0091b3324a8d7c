// runtime.cu -- device / memory / stream / event / graph entry points of include/ncnn_cuda.h.
// CUDA analogue of VulkanDevice + VkAllocator::fastMalloc + VkCompute submit (src/gpu.h, src/allocator.h:267-296,
// src/command.h:22-88 of the reference).
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <string.h>

namespace ncnn_cuda {

static std::mutex g_err_mutex;
static char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches(0);
std::atomic<unsigned long long> g_tc_launches(0); // launches of the tcgen05 implicit-GEMM kernel (tc_gemm.cu)

void set_last_error(const char* what, cudaError_t e, const char* file, int line)
{
    std::lock_guard<std::mutex> lk(g_err_mutex);
    snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    fprintf(stderr, "[ncnn_cuda] %s\n", g_err);
}

void set_last_error_msg(const char* msg)
{
    std::lock_guard<std::mutex> lk(g_err_mutex);
    snprintf(g_err, sizeof(g_err), "%s", msg);
    fprintf(stderr, "[ncnn_cuda] %s\n", g_err);
}

void count_launch(int n)
{
    g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed);
}

void count_tc_launch()
{
    g_tc_launches.fetch_add(1ull, std::memory_order_relaxed);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0)
    {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

int grid_for(long long work_items, int block, int max_waves)
{
    long long blocks = (work_items + block - 1) / block;
    if (blocks < 1) blocks = 1;
    // resident CTAs per SM for a light elementwise kernel: 2048 threads / block
    long long per_wave = (long long)sm_count() * (2048 / block);
    long long cap = per_wave * max_waves;
    if (blocks > cap) blocks = cap;
    return (int)blocks;
}

} // namespace ncnn_cuda

using namespace ncnn_cuda;

extern "C" {

int ncnn_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ncnn_cuda_set_device(int index)
{
    NC_CHECK(cudaSetDevice(index));
    return 0;
}

int ncnn_cuda_get_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) return -1;
    return d;
}

int ncnn_cuda_device_info(int index, char* name, int* sms, int* cc_major, int* cc_minor, size_t* total_mem)
{
    cudaDeviceProp p;
    NC_CHECK(cudaGetDeviceProperties(&p, index));
    if (name)
    {
        strncpy(name, p.name, 255);
        name[255] = 0;
    }
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return 0;
}

const char* ncnn_cuda_last_error(void)
{
    return g_err;
}

int ncnn_cuda_malloc(void** ptr, size_t size)
{
    *ptr = 0;
    NC_CHECK(cudaMalloc(ptr, size));
    return 0;
}

int ncnn_cuda_free(void* ptr)
{
    if (ptr) NC_CHECK(cudaFree(ptr));
    return 0;
}

int ncnn_cuda_malloc_host(void** ptr, size_t size)
{
    *ptr = 0;
    NC_CHECK(cudaMallocHost(ptr, size));
    return 0;
}

int ncnn_cuda_free_host(void* ptr)
{
    if (ptr) NC_CHECK(cudaFreeHost(ptr));
    return 0;
}

int ncnn_cuda_host_is_pinned(const void* ptr)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return attr.type == cudaMemoryTypeHost ? 1 : 0;
}

int ncnn_cuda_memcpy_h2d_async(void* dst, const void* src, size_t size, void* stream)
{
    NC_CHECK(cudaMemcpyAsync(dst, src, size, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

int ncnn_cuda_memcpy_d2h_async(void* dst, const void* src, size_t size, void* stream)
{
    NC_CHECK(cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

int ncnn_cuda_memcpy_d2d_async(void* dst, const void* src, size_t size, void* stream)
{
    NC_CHECK(cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return 0;
}

int ncnn_cuda_memset_async(void* dst, int value, size_t size, void* stream)
{
    NC_CHECK(cudaMemsetAsync(dst, value, size, as_stream(stream)));
    return 0;
}

int ncnn_cuda_stream_create(void** stream)
{
    cudaStream_t s;
    NC_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return 0;
}

int ncnn_cuda_stream_destroy(void* stream)
{
    if (stream) NC_CHECK(cudaStreamDestroy(as_stream(stream)));
    return 0;
}

int ncnn_cuda_stream_sync(void* stream)
{
    NC_CHECK(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int ncnn_cuda_device_sync(void)
{
    NC_CHECK(cudaDeviceSynchronize());
    return 0;
}

int ncnn_cuda_event_create(void** event)
{
    cudaEvent_t e;
    NC_CHECK(cudaEventCreate(&e));
    *event = (void*)e;
    return 0;
}

int ncnn_cuda_event_destroy(void* event)
{
    if (event) NC_CHECK(cudaEventDestroy((cudaEvent_t)event));
    return 0;
}

int ncnn_cuda_event_record(void* event, void* stream)
{
    NC_CHECK(cudaEventRecord((cudaEvent_t)event, as_stream(stream)));
    return 0;
}

int ncnn_cuda_event_sync(void* event)
{
    NC_CHECK(cudaEventSynchronize((cudaEvent_t)event));
    return 0;
}

int ncnn_cuda_event_elapsed_ms(void* start, void* stop, float* ms)
{
    NC_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}

int ncnn_cuda_graph_begin_capture(void* stream)
{
    // relaxed mode: the recorder may still query pointer attributes, and the device pool may still fall back to cudaMalloc for a
    // block the warm-up walk did not leave behind, without invalidating the capture (other threads' streams are unaffected)
    NC_CHECK(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed));
    return 0;
}

int ncnn_cuda_graph_end_capture(void* stream, void** graph_exec)
{
    cudaGraph_t g = 0;
    NC_CHECK(cudaStreamEndCapture(as_stream(stream), &g));
    cudaGraphExec_t ge = 0;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    NC_CHECK(e);
    *graph_exec = (void*)ge;
    return 0;
}

int ncnn_cuda_graph_launch(void* graph_exec, void* stream)
{
    NC_CHECK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, as_stream(stream)));
    return 0;
}

int ncnn_cuda_graph_destroy(void* graph_exec)
{
    if (graph_exec) NC_CHECK(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return 0;
}

unsigned long long ncnn_cuda_launch_count(void)
{
    return g_launches.load(std::memory_order_relaxed);
}

unsigned long long ncnn_cuda_tc_launch_count(void)
{
    return ncnn_cuda::g_tc_launches.load(std::memory_order_relaxed);
}

} // extern "C"
