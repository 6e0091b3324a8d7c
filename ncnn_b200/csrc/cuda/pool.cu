// pool.cu -- Pooling (src/layer/pooling.cpp:39-348, padding rules :350-412 of the reference).
// Bandwidth-bound: one thread owns a 16-byte channel vector of one output pixel; consecutive threads walk
// channels then pixels, so every tap is a fully coalesced row segment of the channel-innermost blob.
// No padded copy is materialised (the reference's copy_make_border sweep disappears): out-of-image taps are
// skipped for max, and contribute 0 / are excluded from the divisor for avg exactly as :255-343.
#include "common.cuh"
#include "pool_tma.cuh"

#include <string.h>

using namespace ncnn_cuda;

namespace {

struct PoolGeom
{
    int C, inw, inh, outw, outh, n;
    int kw, kh, sw, sh, pad_left, pad_top;
    int ax0, ax1, ay0, ay1;
    int type, global, include_pad, adaptive;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
};

template<typename T, int VEC>
__global__ void __launch_bounds__(256) pool_kernel(const T* __restrict__ in, T* __restrict__ out, PoolGeom g)
{
    NC_PDL_PROLOGUE();
    const int CV = (g.C + VEC - 1) / VEC;
    const long long total = (long long)g.n * g.outh * g.outw * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int cv = (int)(idx % CV);
        long long r = idx / CV;
        const int ox = (int)(r % g.outw);
        r /= g.outw;
        const int oy = (int)(r % g.outh);
        const int b = (int)(r / g.outh);
        const int c0 = cv * VEC;
        const T* inb = in + (long long)b * g.in_nstep + c0;

        int y0, y1, x0, x1; // window in input coordinates, [y0,y1) x [x0,x1), may stick out of the image
        if (g.global)
        {
            y0 = 0; y1 = g.inh; x0 = 0; x1 = g.inw;
        }
        else if (g.adaptive)
        {
            y0 = g.inh * oy / g.outh;
            y1 = (g.inh * (oy + 1) + g.outh - 1) / g.outh;
            x0 = g.inw * ox / g.outw;
            x1 = (g.inw * (ox + 1) + g.outw - 1) / g.outw;
        }
        else
        {
            y0 = oy * g.sh - g.pad_top;
            y1 = y0 + g.kh;
            x0 = ox * g.sw - g.pad_left;
            x1 = x0 + g.kw;
        }
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = g.type == 0 ? -FLT_MAX : 0.f;
        int area = 0;
        for (int iy = y0; iy < y1; iy++)
        {
            if (iy < g.ay0 || iy >= g.ay1) continue;
            const bool yin = iy >= 0 && iy < g.inh;
            for (int ix = x0; ix < x1; ix++)
            {
                if (ix < g.ax0 || ix >= g.ax1) continue;
                if (!yin || ix < 0 || ix >= g.inw)
                {
                    area++; // a counted padding tap (SAME modes, avg only): contributes 0
                    continue;
                }
                float xv[VEC];
                load_vec_f32<T, VEC>(inb + ((long long)iy * g.inw + ix) * g.in_cpitch, xv);
                if (g.type == 0)
                {
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = fmaxf(acc[v], xv[v]);
                }
                else
                {
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] += xv[v];
                }
                area++;
            }
        }
        if (g.type == 1)
        {
            if (g.global)
            {
                const float size = (float)(g.inw * g.inh);
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = acc[v] / size;
            }
            else if (g.adaptive)
            {
                const float hk = (float)(y1 - y0), wk = (float)(x1 - x0);
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = acc[v] / hk / wk;
            }
            else if (g.include_pad)
            {
                const float maxk = (float)(g.kw * g.kh);
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = acc[v] / maxk;
            }
            else
            {
                const float a = (float)area; // 0 -> 0/0 = NaN, as the reference's `sum / area`
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = acc[v] / a;
            }
        }
        T* o = out + (long long)b * g.out_nstep + ((long long)oy * g.outw + ox) * g.out_cpitch + c0;
        store_vec_f32<T, VEC>(o, acc);
    }
}

// The window IS the map (global pooling, or ResNet's 7x7 average over a 7x7 map): one output pixel per image.  pool_kernel walks
// the P taps serially in one thread per channel vector -- 64 K threads with 49 dependent steps each for ResNet-50 pool5 (38 us for
// 51 MB).  Here a CTA owns (image, 256-channel chunk): 32 channel vectors x 8 pixel groups, every thread sums its strided share of
// the pixels with all loads in flight, the eight partial sums meet in shared memory.
template<typename T, int VEC>
__global__ void __launch_bounds__(256) pool_whole_map_kernel(const T* __restrict__ in, T* __restrict__ out, int P, int C, int in_cpitch, int out_cpitch, long long in_nstep,
                                                             long long out_nstep, int chunks, int type, float divisor)
{
    NC_PDL_PROLOGUE();
    __shared__ float part[8][32][VEC + 1];
    const int b = blockIdx.x / chunks;
    const int chunk = blockIdx.x - b * chunks;
    const int cv = threadIdx.x & 31, pg = threadIdx.x >> 5;
    const int c0 = (chunk * 32 + cv) * VEC;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = type == 0 ? -FLT_MAX : 0.f;
    if (c0 < C)
    {
        const T* src = in + (long long)b * in_nstep + c0;
#pragma unroll 4
        for (int p = pg; p < P; p += 8)
        {
            float xv[VEC];
            load_vec_f32<T, VEC>(src + (long long)p * in_cpitch, xv);
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = type == 0 ? fmaxf(acc[v], xv[v]) : acc[v] + xv[v];
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) part[pg][cv][v] = acc[v];
    __syncthreads();
    if (pg == 0 && c0 < C)
    {
#pragma unroll
        for (int k = 1; k < 8; k++)
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = type == 0 ? fmaxf(acc[v], part[k][cv][v]) : acc[v] + part[k][cv][v];
        if (type == 1)
        {
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = acc[v] / divisor;
        }
        store_vec_f32<T, VEC>(out + (long long)b * out_nstep + c0, acc);
    }
}

template<typename T>
static int run_pool(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const PoolGeom& g, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    {
        const int cround0 = ((g.C + VEC - 1) / VEC) * VEC;
        const bool aligned = (g.in_cpitch % VEC == 0) && (g.out_cpitch % VEC == 0) && (g.in_nstep % VEC == 0) && (g.out_nstep % VEC == 0) &&
                             (((uintptr_t)bottom->data & 15) == 0) && (((uintptr_t)top->data & 15) == 0) && cround0 <= g.in_cpitch && cround0 <= g.out_cpitch;
        // the window covers exactly the map: global pooling, or an unpadded kernel of the map's size with every tap counted
        const bool whole = g.global || (!g.adaptive && g.outw == 1 && g.outh == 1 && g.kw == g.inw && g.kh == g.inh && g.pad_left == 0 && g.pad_top == 0 && g.ax0 <= 0 &&
                                        g.ay0 <= 0 && g.ax1 >= g.inw && g.ay1 >= g.inh);
        if (aligned && whole && g.inw * g.inh >= 16 && g.n > 0)
        {
            const int chunks = (cround0 / VEC + 31) / 32;
            NC_PDL_LAUNCH((pool_whole_map_kernel<T, VEC>), g.n * chunks, 256, 0, stream, (const T*)bottom->data, (T*)top->data, g.inw * g.inh, g.C, g.in_cpitch, g.out_cpitch,
                          g.in_nstep, g.out_nstep, chunks, g.type, (float)(g.inw * g.inh));
            NC_LAUNCH_CHECK();
            return 0;
        }
    }
    if (g.type == 0 && !g.global && !g.adaptive && g.kw == g.kh && g.sw == g.sh && g.n > 0)
    {
        // the bandwidth path: TMA-staged tiles with NaN out-of-bounds fill (pool_tma.cuh); what it declines runs below
        plt::Call c;
        c.in = bottom->data;
        c.out = top->data;
        c.elemtype = bottom->elemtype;
        c.C = g.C;
        c.inw = g.inw;
        c.inh = g.inh;
        c.outw = g.outw;
        c.outh = g.outh;
        c.n = g.n;
        c.kernel = g.kw;
        c.stride = g.sw;
        c.pad_left = g.pad_left;
        c.pad_top = g.pad_top;
        c.in_cpitch = g.in_cpitch;
        c.out_cpitch = g.out_cpitch;
        c.in_nstep = g.in_nstep;
        c.out_nstep = g.out_nstep;
        int r = plt::forward<T>(c, stream);
        if (r <= 0) return r;
    }
    const int cround = ((g.C + VEC - 1) / VEC) * VEC;
    const bool vec_ok = (g.in_cpitch % VEC == 0) && (g.out_cpitch % VEC == 0) && (g.in_nstep % VEC == 0) && (g.out_nstep % VEC == 0)
                        && (((uintptr_t)bottom->data & 15) == 0) && (((uintptr_t)top->data & 15) == 0) && cround <= g.in_cpitch && cround <= g.out_cpitch;
    if (vec_ok)
    {
        long long total = (long long)g.n * g.outh * g.outw * (cround / VEC);
        NC_PDL_LAUNCH((pool_kernel<T, VEC>), grid_for(total, 256, 16), 256, 0, stream, (const T*)bottom->data, (T*)top->data, g);
    }
    else
    {
        long long total = (long long)g.n * g.outh * g.outw * g.C;
        NC_PDL_LAUNCH((pool_kernel<T, 1>), grid_for(total, 256, 16), 256, 0, stream, (const T*)bottom->data, (T*)top->data, g);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" int ncnn_cuda_pool2d_forward(const ncnn_cuda_pool2d_desc* d, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(d && bottom && top && bottom->dims == 3, "pool2d_forward: a 3-D bottom blob is required");
    NC_REQUIRE(bottom->elemtype == top->elemtype, "pool2d_forward: element types differ");
    TView bv = make_view(bottom), tv = make_view(top);
    NC_REQUIRE(bv.n == tv.n && tv.C == bottom->c, "pool2d_forward: batch/channel mismatch");
    PoolGeom g;
    g.C = bottom->c;
    g.inw = bottom->w;
    g.inh = bottom->h;
    g.n = bv.n;
    g.global = d->global_pooling;
    if (g.global)
    {
        g.outw = 1;
        g.outh = 1;
    }
    else
    {
        NC_REQUIRE(top->dims == 3, "pool2d_forward: a 3-D top blob is required");
        g.outw = top->w;
        g.outh = top->h;
    }
    g.kw = d->kernel_w;
    g.kh = d->kernel_h;
    g.sw = d->stride_w;
    g.sh = d->stride_h;
    g.pad_left = d->pad_left;
    g.pad_top = d->pad_top;
    if (d->pooling_type == 1 && !d->global_pooling && !d->adaptive_pooling && !d->avgpool_count_include_pad)
    {
        g.ax0 = d->area_x0; g.ax1 = d->area_x1; g.ay0 = d->area_y0; g.ay1 = d->area_y1;
    }
    else
    {
        g.ax0 = 0; g.ax1 = bottom->w; g.ay0 = 0; g.ay1 = bottom->h;
    }
    g.type = d->pooling_type;
    g.include_pad = d->avgpool_count_include_pad;
    g.adaptive = d->adaptive_pooling;
    g.in_cpitch = bottom->cpitch;
    g.out_cpitch = top->cpitch;
    g.in_nstep = bottom->nstep;
    g.out_nstep = top->nstep;
    NC_REQUIRE(g.type == 0 || g.type == 1, "pool2d_forward: pooling_type must be 0 (max) or 1 (avg)");
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32: return run_pool<float>(bottom, top, g, stream);
    case NCNN_CUDA_BF16: return run_pool<__nv_bfloat16>(bottom, top, g, stream);
    case NCNN_CUDA_F16: return run_pool<__half>(bottom, top, g, stream);
    }
    return -1;
}
