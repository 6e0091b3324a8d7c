// deconv.cu -- Deconvolution / DeconvolutionDepthWise (src/layer/deconvolution.cpp:68-146,
// src/layer/deconvolutiondepthwise.cpp:68-208 of the reference), SURVEY.md 8(f3).
// The reference scatters every input pixel into a bordered output of size (w-1)*stride + extent + output_pad and
// then cuts the pads away (cut_padding, deconvolution.cpp:364-392).  On the device the same sum is a GATHER: one
// thread owns one output element of the already-cut blob, walks the kernel taps whose source pixel
// (oy + cut_top - ky*dilation) / stride exists, and accumulates over the group's input channels in fp32, so there is
// neither a bordered scratch blob nor atomics.  Channel-innermost blobs make the input read of a pixel a broadcast
// across the threads of that pixel; weights are re-packed once to [group][tap][inch_g][outch_g] so that consecutive
// threads (consecutive output channels) read consecutive weights.  Not on the named models' path: HBM/L2-bound
// CUDA-core kernel, no tensor-core variant yet.
#include "common.cuh"

#include <vector>

using namespace ncnn_cuda;

struct ncnn_cuda_deconv2d
{
    ncnn_cuda_deconv2d_desc desc;
    int taps;
    float* w_dev;    // [group][tap][inch_g][outch_g] fp32
    float* bias_dev; // [outch] or NULL
};

namespace {

struct DeconvGeom
{
    int inch_g, outch_g, group;
    int inw, inh, outw, outh, n;
    int kw, kh, dw, dh, sw, sh, cut_left, cut_top;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
    int act_type;
    float act_p0, act_p1;
};

template<typename T>
__global__ void __launch_bounds__(256) deconv_gather_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out, DeconvGeom g)
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
        const int p = oc - grp * g.outch_g;
        const float* wg = w + (long long)grp * taps * g.inch_g * g.outch_g + p;
        const T* inb = in + (long long)b * g.in_nstep + grp * g.inch_g;
        float sum = bias ? bias[oc] : 0.f;
        const int fy = oy + g.cut_top, fx = ox + g.cut_left; // position in the reference's bordered output
        for (int ky = 0; ky < g.kh; ky++)
        {
            const int ty = fy - ky * g.dh;
            if (ty < 0) break; // taps are visited in increasing ky, ty only decreases
            const int iy = ty / g.sh;
            if (iy * g.sh != ty || iy >= g.inh) continue;
            for (int kx = 0; kx < g.kw; kx++)
            {
                const int tx = fx - kx * g.dw;
                if (tx < 0) break;
                const int ix = tx / g.sw;
                if (ix * g.sw != tx || ix >= g.inw) continue;
                const T* px = inb + ((long long)iy * g.inw + ix) * g.in_cpitch;
                const float* wk = wg + (long long)(ky * g.kw + kx) * g.inch_g * g.outch_g;
                for (int q = 0; q < g.inch_g; q++) sum = fmaf(to_f32(px[q]), wk[(long long)q * g.outch_g], sum);
            }
        }
        out[(long long)b * g.out_nstep + ((long long)oy * g.outw + ox) * g.out_cpitch + oc] = from_f32<T>(apply_activation(sum, g.act_type, g.act_p0, g.act_p1));
    }
}

template<typename T>
static int run_deconv(const ncnn_cuda_deconv2d* conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const DeconvGeom& g, cudaStream_t stream)
{
    long long total = (long long)g.n * g.outh * g.outw * g.outch_g * g.group;
    NC_PDL_LAUNCH((deconv_gather_kernel<T>), grid_for(total, 256, 16), 256, 0, stream, (const T*)bottom->data, conv->w_dev, conv->bias_dev, (T*)top->data, g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" {

int ncnn_cuda_deconv2d_create(ncnn_cuda_deconv2d_t* out, const ncnn_cuda_deconv2d_desc* desc, const float* weight, const float* bias, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    *out = 0;
    NC_REQUIRE(desc->group > 0 && desc->inch % desc->group == 0 && desc->outch % desc->group == 0, "deconv2d_create: channels not divisible by group");
    NC_REQUIRE(desc->kernel_w > 0 && desc->kernel_h > 0 && desc->stride_w > 0 && desc->stride_h > 0 && desc->dilation_w > 0 && desc->dilation_h > 0, "deconv2d_create: bad geometry");
    ncnn_cuda_deconv2d* c = new ncnn_cuda_deconv2d;
    memset(c, 0, sizeof(*c));
    c->desc = *desc;
    c->taps = desc->kernel_w * desc->kernel_h;
    const int inch_g = desc->inch / desc->group, outch_g = desc->outch / desc->group;
    // reference order [group][outch_g][inch_g][tap] (deconvolutiondepthwise.cpp:160-181) -> [group][tap][inch_g][outch_g]
    std::vector<float> host((size_t)desc->group * c->taps * inch_g * outch_g);
    for (int g = 0; g < desc->group; g++)
        for (int p = 0; p < outch_g; p++)
            for (int q = 0; q < inch_g; q++)
                for (int t = 0; t < c->taps; t++)
                    host[(((size_t)g * c->taps + t) * inch_g + q) * outch_g + p] = weight[(((size_t)g * outch_g + p) * inch_g + q) * c->taps + t];
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
        set_last_error("deconv2d_create upload", e, __FILE__, __LINE__);
        ncnn_cuda_deconv2d_destroy(c);
        return -100;
    }
    *out = c;
    return 0;
}

int ncnn_cuda_deconv2d_destroy(ncnn_cuda_deconv2d_t c)
{
    if (!c) return 0;
    if (c->w_dev) cudaFree(c->w_dev);
    if (c->bias_dev) cudaFree(c->bias_dev);
    delete c;
    return 0;
}

int ncnn_cuda_deconv2d_forward(ncnn_cuda_deconv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int cut_left, int cut_top, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && bottom && top && bottom->dims == 3 && top->dims == 3, "deconv2d_forward: 3-D blobs required");
    NC_REQUIRE(bottom->c == conv->desc.inch && top->c == conv->desc.outch && bottom->elemtype == top->elemtype, "deconv2d_forward: blob does not match the layer");
    NC_REQUIRE(cut_left >= 0 && cut_top >= 0, "deconv2d_forward: negative cut");
    const ncnn_cuda_deconv2d_desc& d = conv->desc;
    TView bv = make_view(bottom), tv = make_view(top);
    NC_REQUIRE(bv.n == tv.n, "deconv2d_forward: batch mismatch");
    // the cut blob must lie inside the reference's bordered output (deconvolution.cpp:156-157)
    const int full_w = (bottom->w - 1) * d.stride_w + d.dilation_w * (d.kernel_w - 1) + 1 + d.output_pad_right;
    const int full_h = (bottom->h - 1) * d.stride_h + d.dilation_h * (d.kernel_h - 1) + 1 + d.output_pad_bottom;
    NC_REQUIRE(cut_left + top->w <= full_w && cut_top + top->h <= full_h, "deconv2d_forward: top blob exceeds the bordered output");
    DeconvGeom g;
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
    g.cut_left = cut_left;
    g.cut_top = cut_top;
    g.in_cpitch = bottom->cpitch;
    g.out_cpitch = top->cpitch;
    g.in_nstep = bottom->nstep;
    g.out_nstep = top->nstep;
    g.act_type = d.act.type;
    g.act_p0 = d.act.p0;
    g.act_p1 = d.act.p1;
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32: return run_deconv<float>(conv, bottom, top, g, stream);
    case NCNN_CUDA_BF16: return run_deconv<__nv_bfloat16>(conv, bottom, top, g, stream);
    case NCNN_CUDA_F16: return run_deconv<__half>(conv, bottom, top, g, stream);
    }
    return -1;
}

} // extern "C"
