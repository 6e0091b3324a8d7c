// dwconv.cu -- ConvolutionDepthWise (src/layer/convolutiondepthwise.cpp:146-270 of the reference).
// Depthwise branch (:181-214): bandwidth-bound; channel-innermost blobs make every tap a 16-byte
// vector load per thread, consecutive threads cover consecutive channels then consecutive pixels, so a
// warp's request is one contiguous run of the input row.  Each thread produces OW_PER_THREAD horizontally
// adjacent outputs so that the kw + (OW-1)*stride input columns it loads are reused from registers; vertical
// reuse comes from L1/L2 (a CTA covers whole output rows).  Accumulation is fp32 in the reference's tap order.
// Grouped branch (:216-267): plain CUDA-core kernel (not on the named models' path).
#include "common.cuh"
#include "dwconv_tma.cuh"

#include <string.h>
#include <vector>

using namespace ncnn_cuda;

struct ncnn_cuda_dwconv2d
{
    ncnn_cuda_dwconv2d_desc desc;
    int taps;
    bool depthwise;
    float* w_dev;    // depthwise: [taps][cpad] fp32 ; grouped: reference order
    int cpad;
    float* bias_dev; // [outch] or NULL
};

namespace {

struct DwGeom
{
    int C, inw, inh, outw, outh, n;
    int kw, kh, dw, dh, sw, sh, pad_left, pad_top;
    float pad_value;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
    int cpad;
    int act_type;
    float act_p0, act_p1;
};

// VEC channels x OWT outputs per thread
template<typename T, int VEC, int OWT>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out, DwGeom g)
{
    NC_PDL_PROLOGUE();
    const int CV = (g.C + VEC - 1) / VEC;
    const int OWB = (g.outw + OWT - 1) / OWT;
    const long long total = (long long)g.n * g.outh * OWB * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int cv = (int)(idx % CV);
        long long r = idx / CV;
        const int owb = (int)(r % OWB);
        r /= OWB;
        const int oy = (int)(r % g.outh);
        const int b = (int)(r / g.outh);
        const int c0 = cv * VEC;
        const int ox0 = owb * OWT;

        float acc[OWT][VEC];
#pragma unroll
        for (int o = 0; o < OWT; o++)
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[o][v] = 0.f;

        const T* inb = in + (long long)b * g.in_nstep + c0;
        for (int ky = 0; ky < g.kh; ky++)
        {
            const int iy = oy * g.sh - g.pad_top + ky * g.dh;
            const bool yok = iy >= 0 && iy < g.inh;
            for (int kx = 0; kx < g.kw; kx++)
            {
                float wv[VEC];
                {
                    const float* wp = w + (long long)(ky * g.kw + kx) * g.cpad + c0;
                    if constexpr (VEC == 4)
                    {
                        float4 t = *reinterpret_cast<const float4*>(wp);
                        wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
                    }
                    else if constexpr (VEC == 8)
                    {
                        float4 t0 = *reinterpret_cast<const float4*>(wp), t1 = *reinterpret_cast<const float4*>(wp + 4);
                        wv[0] = t0.x; wv[1] = t0.y; wv[2] = t0.z; wv[3] = t0.w;
                        wv[4] = t1.x; wv[5] = t1.y; wv[6] = t1.z; wv[7] = t1.w;
                    }
                    else
                    {
#pragma unroll
                        for (int v = 0; v < VEC; v++) wv[v] = wp[v];
                    }
                }
#pragma unroll
                for (int o = 0; o < OWT; o++)
                {
                    const int ox = ox0 + o;
                    const int ix = ox * g.sw - g.pad_left + kx * g.dw;
                    float xv[VEC];
                    if (yok && ix >= 0 && ix < g.inw && ox < g.outw)
                    {
                        load_vec_f32<T, VEC>(inb + ((long long)iy * g.inw + ix) * g.in_cpitch, xv);
                    }
                    else
                    {
#pragma unroll
                        for (int v = 0; v < VEC; v++) xv[v] = g.pad_value;
                    }
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[o][v] = fmaf(xv[v], wv[v], acc[o][v]);
                }
            }
        }
        float bv[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) bv[v] = bias ? bias[c0 + v < g.C ? c0 + v : g.C - 1] : 0.f;
#pragma unroll
        for (int o = 0; o < OWT; o++)
        {
            const int ox = ox0 + o;
            if (ox >= g.outw) continue;
            float ov[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) ov[v] = apply_activation(acc[o][v] + bv[v], g.act_type, g.act_p0, g.act_p1);
            store_vec_f32<T, VEC>(out + (long long)b * g.out_nstep + ((long long)oy * g.outw + ox) * g.out_cpitch + c0, ov);
        }
    }
}

struct GroupGeom
{
    int inch_g, outch_g, group;
    int inw, inh, outw, outh, n;
    int kw, kh, dw, dh, sw, sh, pad_left, pad_top;
    float pad_value;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
    int act_type;
    float act_p0, act_p1;
};

template<typename T>
__global__ void grouped_conv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out, GroupGeom g)
{
    NC_PDL_PROLOGUE();
    const int outch = g.outch_g * g.group;
    const long long total = (long long)g.n * g.outh * g.outw * outch;
    const int taps = g.kw * g.kh;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int oc = (int)(idx % outch);
        long long r = idx / outch;
        const int ox = (int)(r % g.outw);
        r /= g.outw;
        const int oy = (int)(r % g.outh);
        const int b = (int)(r / g.outh);
        const int grp = oc / g.outch_g;
        const float* wk = w + (long long)oc * g.inch_g * taps;
        float sum = bias ? bias[oc] : 0.f;
        for (int q = 0; q < g.inch_g; q++)
        {
            const int ci = grp * g.inch_g + q;
            for (int ky = 0; ky < g.kh; ky++)
            {
                const int iy = oy * g.sh - g.pad_top + ky * g.dh;
                for (int kx = 0; kx < g.kw; kx++)
                {
                    const int ix = ox * g.sw - g.pad_left + kx * g.dw;
                    float x = g.pad_value;
                    if (iy >= 0 && iy < g.inh && ix >= 0 && ix < g.inw) x = to_f32(in[(long long)b * g.in_nstep + ((long long)iy * g.inw + ix) * g.in_cpitch + ci]);
                    sum = fmaf(x, wk[q * taps + ky * g.kw + kx], sum);
                }
            }
        }
        out[(long long)b * g.out_nstep + ((long long)oy * g.outw + ox) * g.out_cpitch + oc] = from_f32<T>(apply_activation(sum, g.act_type, g.act_p0, g.act_p1));
    }
}

template<typename T>
static int run_depthwise(const ncnn_cuda_dwconv2d* conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const DwGeom& g, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    const bool vec_ok = (g.in_cpitch % VEC == 0) && (g.out_cpitch % VEC == 0) && (g.in_nstep % VEC == 0) && (g.out_nstep % VEC == 0)
                        && (((uintptr_t)bottom->data & 15) == 0) && (((uintptr_t)top->data & 15) == 0) && (g.cpad % VEC == 0)
                        && (((g.C + VEC - 1) / VEC) * VEC <= g.in_cpitch) && (((g.C + VEC - 1) / VEC) * VEC <= g.out_cpitch);
    const T* in = (const T*)bottom->data;
    T* out = (T*)top->data;
    if (g.kw == 3 && g.kh == 3 && g.dw == 1 && g.dh == 1 && g.sw == g.sh && g.pad_value == 0.f && g.n > 0)
    {
        // the bandwidth path: TMA-staged halo tiles (dwconv_tma.cuh); anything it declines runs on the generic kernel below
        dwt::Call c;
        c.in = in;
        c.out = out;
        c.elemtype = bottom->elemtype;
        c.C = g.C;
        c.inw = g.inw;
        c.inh = g.inh;
        c.outw = g.outw;
        c.outh = g.outh;
        c.n = g.n;
        c.stride = g.sw;
        c.pad_left = g.pad_left;
        c.pad_top = g.pad_top;
        c.in_cpitch = g.in_cpitch;
        c.out_cpitch = g.out_cpitch;
        c.in_nstep = g.in_nstep;
        c.out_nstep = g.out_nstep;
        c.w = conv->w_dev;
        c.bias = conv->bias_dev;
        c.cpad = g.cpad;
        c.act_type = g.act_type;
        c.act_p0 = g.act_p0;
        c.act_p1 = g.act_p1;
        int r = dwt::forward<T>(c, stream);
        if (r <= 0) return r;
    }
    if (vec_ok)
    {
        const int CV = (g.C + VEC - 1) / VEC;
        if (g.outw >= 8)
        {
            constexpr int OWT = 4;
            long long total = (long long)g.n * g.outh * ((g.outw + OWT - 1) / OWT) * CV;
            NC_PDL_LAUNCH((dwconv_kernel<T, VEC, OWT>), grid_for(total, 256, 16), 256, 0, stream, in, conv->w_dev, conv->bias_dev, out, g);
        }
        else
        {
            long long total = (long long)g.n * g.outh * g.outw * CV;
            NC_PDL_LAUNCH((dwconv_kernel<T, VEC, 1>), grid_for(total, 256, 16), 256, 0, stream, in, conv->w_dev, conv->bias_dev, out, g);
        }
    }
    else
    {
        long long total = (long long)g.n * g.outh * g.outw * g.C;
        NC_PDL_LAUNCH((dwconv_kernel<T, 1, 1>), grid_for(total, 256, 16), 256, 0, stream, in, conv->w_dev, conv->bias_dev, out, g);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

template<typename T>
static int run_grouped(const ncnn_cuda_dwconv2d* conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const GroupGeom& g, cudaStream_t stream)
{
    long long total = (long long)g.n * g.outh * g.outw * g.outch_g * g.group;
    NC_PDL_LAUNCH((grouped_conv_kernel<T>), grid_for(total, 256, 16), 256, 0, stream, (const T*)bottom->data, conv->w_dev, conv->bias_dev, (T*)top->data, g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" {

int ncnn_cuda_dwconv2d_create(ncnn_cuda_dwconv2d_t* out, const ncnn_cuda_dwconv2d_desc* desc, const float* weight, const float* bias, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    *out = 0;
    NC_REQUIRE(desc->group > 0 && desc->inch % desc->group == 0 && desc->outch % desc->group == 0, "dwconv2d_create: channels not divisible by group");
    ncnn_cuda_dwconv2d* c = new ncnn_cuda_dwconv2d;
    memset(c, 0, sizeof(*c));
    c->desc = *desc;
    c->taps = desc->kernel_w * desc->kernel_h;
    c->depthwise = desc->group == desc->inch && desc->group == desc->outch;
    std::vector<float> host;
    if (c->depthwise)
    {
        c->cpad = ((desc->inch + 7) / 8) * 8;
        host.assign((size_t)c->taps * c->cpad, 0.f);
        for (int ch = 0; ch < desc->inch; ch++)
            for (int t = 0; t < c->taps; t++) host[(size_t)t * c->cpad + ch] = weight[(size_t)ch * c->taps + t];
    }
    else
    {
        size_t count = (size_t)desc->outch * (desc->inch / desc->group) * c->taps;
        host.assign(weight, weight + count);
    }
    cudaError_t e = cudaMalloc((void**)&c->w_dev, host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->w_dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && desc->bias_term && bias)
    {
        e = cudaMalloc((void**)&c->bias_dev, sizeof(float) * desc->outch);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->bias_dev, bias, sizeof(float) * desc->outch, cudaMemcpyHostToDevice, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess)
    {
        set_last_error("dwconv2d_create upload", e, __FILE__, __LINE__);
        ncnn_cuda_dwconv2d_destroy(c);
        return -100;
    }
    *out = c;
    return 0;
}

int ncnn_cuda_dwconv2d_destroy(ncnn_cuda_dwconv2d_t c)
{
    if (!c) return 0;
    if (c->w_dev) cudaFree(c->w_dev);
    if (c->bias_dev) cudaFree(c->bias_dev);
    delete c;
    return 0;
}

int ncnn_cuda_dwconv2d_forward(ncnn_cuda_dwconv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int pad_left, int pad_top, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && bottom && top && bottom->dims == 3 && top->dims == 3, "dwconv2d_forward: 3-D blobs required");
    NC_REQUIRE(bottom->c == conv->desc.inch && top->c == conv->desc.outch && bottom->elemtype == top->elemtype, "dwconv2d_forward: blob does not match the layer");
    const ncnn_cuda_dwconv2d_desc& d = conv->desc;
    TView bv = make_view(bottom), tv = make_view(top);
    NC_REQUIRE(bv.n == tv.n, "dwconv2d_forward: batch mismatch");
    if (conv->depthwise)
    {
        DwGeom g;
        g.C = d.inch;
        g.inw = bottom->w;
        g.inh = bottom->h;
        g.outw = top->w;
        g.outh = top->h;
        g.n = bv.n;
        g.kw = d.kernel_w;
        g.kh = d.kernel_h;
        g.dw = d.dilation_w;
        g.dh = d.dilation_h;
        g.sw = d.stride_w;
        g.sh = d.stride_h;
        g.pad_left = pad_left;
        g.pad_top = pad_top;
        g.pad_value = d.pad_value;
        g.in_cpitch = bottom->cpitch;
        g.out_cpitch = top->cpitch;
        g.in_nstep = bottom->nstep;
        g.out_nstep = top->nstep;
        g.cpad = conv->cpad;
        g.act_type = d.act.type;
        g.act_p0 = d.act.p0;
        g.act_p1 = d.act.p1;
        switch (bottom->elemtype)
        {
        case NCNN_CUDA_F32: return run_depthwise<float>(conv, bottom, top, g, stream);
        case NCNN_CUDA_BF16: return run_depthwise<__nv_bfloat16>(conv, bottom, top, g, stream);
        case NCNN_CUDA_F16: return run_depthwise<__half>(conv, bottom, top, g, stream);
        }
        return -1;
    }
    GroupGeom g;
    g.group = d.group;
    g.inch_g = d.inch / d.group;
    g.outch_g = d.outch / d.group;
    g.inw = bottom->w;
    g.inh = bottom->h;
    g.outw = top->w;
    g.outh = top->h;
    g.n = bv.n;
    g.kw = d.kernel_w;
    g.kh = d.kernel_h;
    g.dw = d.dilation_w;
    g.dh = d.dilation_h;
    g.sw = d.stride_w;
    g.sh = d.stride_h;
    g.pad_left = pad_left;
    g.pad_top = pad_top;
    g.pad_value = d.pad_value;
    g.in_cpitch = bottom->cpitch;
    g.out_cpitch = top->cpitch;
    g.in_nstep = bottom->nstep;
    g.out_nstep = top->nstep;
    g.act_type = d.act.type;
    g.act_p0 = d.act.p0;
    g.act_p1 = d.act.p1;
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32: return run_grouped<float>(conv, bottom, top, g, stream);
    case NCNN_CUDA_BF16: return run_grouped<__nv_bfloat16>(conv, bottom, top, g, stream);
    case NCNN_CUDA_F16: return run_grouped<__half>(conv, bottom, top, g, stream);
    }
    return -1;
}

} // extern "C"
