// common.cuh -- shared device/host helpers for the sm_100a kernels behind include/ncnn_cuda.h
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ncnn_cuda.h"

namespace ncnn_cuda {

// ---------------------------------------------------------------- error plumbing
void set_last_error(const char* what, cudaError_t e, const char* file, int line);
void set_last_error_msg(const char* msg);
void count_launch(int n = 1);
void count_tc_launch();

#define NC_CHECK(expr)                                                   \
    do                                                                   \
    {                                                                    \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess)                                           \
        {                                                                \
            ncnn_cuda::set_last_error(#expr, _e, __FILE__, __LINE__);    \
            return -100;                                                 \
        }                                                                \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define NC_LAUNCH_CHECK()                                                            \
    do                                                                               \
    {                                                                                \
        ncnn_cuda::count_launch();                                                   \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess)                                                       \
        {                                                                            \
            ncnn_cuda::set_last_error("kernel launch", _e, __FILE__, __LINE__);      \
            return -100;                                                             \
        }                                                                            \
    } while (0)

#define NC_REQUIRE(cond, msg)                          \
    do                                                 \
    {                                                  \
        if (!(cond))                                   \
        {                                              \
            ncnn_cuda::set_last_error_msg(msg);        \
            return -1;                                 \
        }                                              \
    } while (0)

static inline cudaStream_t as_stream(void* s)
{
    return (cudaStream_t)s;
}

int sm_count();

// ---------------------------------------------------------------- element types
template<typename T>
struct ElemTraits;
template<>
struct ElemTraits<float>
{
    static const int id = NCNN_CUDA_F32;
};
template<>
struct ElemTraits<__nv_bfloat16>
{
    static const int id = NCNN_CUDA_BF16;
};
template<>
struct ElemTraits<__half>
{
    static const int id = NCNN_CUDA_F16;
};

static inline size_t elem_size(int elemtype)
{
    return elemtype == NCNN_CUDA_F32 ? 4 : 2;
}

__device__ __forceinline__ float to_f32(float v)
{
    return v;
}
__device__ __forceinline__ float to_f32(__nv_bfloat16 v)
{
    return __bfloat162float(v);
}
__device__ __forceinline__ float to_f32(__half v)
{
    return __half2float(v);
}
template<typename T>
__device__ __forceinline__ T from_f32(float v);
template<>
__device__ __forceinline__ float from_f32<float>(float v)
{
    return v;
}
template<>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v)
{
    return __float2bfloat16_rn(v);
}
template<>
__device__ __forceinline__ __half from_f32<__half>(float v)
{
    return __float2half_rn(v);
}

// VEC elements of T as one aligned vector (16 bytes for VEC*sizeof(T)==16)
template<typename T, int VEC>
struct alignas(sizeof(T) * VEC) Vec
{
    T v[VEC];
};

template<typename T, int VEC>
__device__ __forceinline__ void load_vec_f32(const T* p, float (&out)[VEC])
{
    Vec<T, VEC> t = *reinterpret_cast<const Vec<T, VEC>*>(p);
#pragma unroll
    for (int i = 0; i < VEC; i++) out[i] = to_f32(t.v[i]);
}

template<typename T, int VEC>
__device__ __forceinline__ void store_vec_f32(T* p, const float (&in)[VEC])
{
    Vec<T, VEC> t;
#pragma unroll
    for (int i = 0; i < VEC; i++) t.v[i] = from_f32<T>(in[i]);
    *reinterpret_cast<Vec<T, VEC>*>(p) = t;
}

// ---------------------------------------------------------------- tensor view helpers
struct TView
{
    int dims, w, h, d, c, n;
    int cpitch;
    long long nstep;
    int P; // pixels per sample
    int C; // channels per pixel
};

static inline TView make_view(const ncnn_cuda_tensor* t)
{
    TView v;
    v.dims = t->dims;
    v.w = t->w;
    v.h = t->h;
    v.d = t->d;
    v.c = t->c;
    v.n = t->n < 1 ? 1 : t->n;
    v.cpitch = t->cpitch;
    v.nstep = t->nstep;
    if (t->dims == 1)
    {
        v.P = 1;
        v.C = t->w;
    }
    else if (t->dims == 2)
    {
        v.P = t->h;
        v.C = t->w;
    }
    else if (t->dims == 3)
    {
        v.P = t->h * t->w;
        v.C = t->c;
    }
    else
    {
        v.P = t->d * t->h * t->w;
        v.C = t->c;
    }
    return v;
}

static inline bool same_shape(const ncnn_cuda_tensor* a, const ncnn_cuda_tensor* b)
{
    return a->dims == b->dims && a->w == b->w && a->h == b->h && a->d == b->d && a->c == b->c && a->n == b->n;
}

// ---------------------------------------------------------------- fused activation
// src/layer/fused_activation.h:10-64 (codes 0..6); 7 = swish (graph-level Conv+Swish fold)
__device__ __forceinline__ float apply_activation(float v, int type, float p0, float p1)
{
    switch (type)
    {
    case 1:
        v = fmaxf(v, 0.f);
        break;
    case 2:
        v = v > 0.f ? v : v * p0;
        break;
    case 3:
        if (v < p0) v = p0;
        if (v > p1) v = p1;
        break;
    case 4:
        v = fminf(v, 88.3762626647949f);
        v = fmaxf(v, -88.3762626647949f);
        v = 1.f / (1.f + expf(-v));
        break;
    case 5:
        v = v * tanhf(logf(expf(v) + 1.f));
        break;
    case 6:
    {
        float lower = -p1 / p0;
        float upper = (1.f / p0) + lower;
        if (v < lower)
            v = 0.f;
        else if (v > upper)
            ;
        else
            v = v * (v * p0 + p1);
        break;
    }
    case 7:
        v = v / (1.f + expf(-v));
        break;
    }
    return v;
}

// launch with the programmatic-stream-serialization attribute (see tc::pdl_wait in tc_gemm.cuh); the kernel MUST call
// griddepcontrol.wait before it touches memory the previous kernel of the stream reads or writes
template<typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// every kernel of the backend starts with NC_PDL_PROLOGUE() and is launched with NC_PDL_LAUNCH: the next kernel of the stream
// may be scheduled while this one drains (its launch latency disappears), and no kernel touches memory before its
// predecessor has completed.  The TMA kernels place the wait after their on-chip set-up instead (tc::pdl_wait).
#define NC_PDL_PROLOGUE()                                                  \
    do                                                                     \
    {                                                                      \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    \
        asm volatile("griddepcontrol.wait;" ::: "memory");                 \
    } while (0)

#define NC_PDL_LAUNCH(kernel, grid, block, smem, stream, ...)                                                                   \
    do                                                                                                                          \
    {                                                                                                                           \
        cudaError_t _le = ncnn_cuda::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__);          \
        if (_le != cudaSuccess)                                                                                                 \
        {                                                                                                                       \
            ncnn_cuda::set_last_error("kernel launch", _le, __FILE__, __LINE__);                                               \
            return -100;                                                                                                        \
        }                                                                                                                       \
    } while (0)

static inline int ceil_div(long long a, long long b)
{
    return (int)((a + b - 1) / b);
}

// grid size for a grid-stride elementwise kernel: enough CTAs to fill 148 SMs a few times, not more
int grid_for(long long work_items, int block, int max_waves = 8);

} // namespace ncnn_cuda
