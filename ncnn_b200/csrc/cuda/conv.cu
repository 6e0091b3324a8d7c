// conv.cu -- Convolution / InnerProduct entry points of include/ncnn_cuda.h.
// Reference semantics: src/layer/convolution.cpp:113-184 (loop), :328-372 (padding),
// src/layer/innerproduct.cpp:84-165.  InnerProduct over a (w,h,c) blob is a convolution whose kernel
// covers the whole blob: the reference's flattened weight order [outch][c][h][w] (innerproduct.cpp:141-162)
// is exactly a conv weight [outch][inch][kh][kw].
#include "common.cuh"
#include "conv_simt.cuh"
#include "tc_gemm.cuh"

#include <string.h>
#include <vector>

using namespace ncnn_cuda;

struct ncnn_cuda_conv2d
{
    ncnn_cuda_conv2d_desc desc;
    int taps;
    int K;
    float* wp_simt; // [Kpad][wp_ld] fp32, k = tap*inch + ci
    int wp_ld;
    float* bias_dev; // [outch] or NULL
    TcPlan tc;
    bool has_tc;
    // folded projection shortcut (ncnn_cuda_conv2d_fuse_shortcut): one GEMM over K = [this layer's channels | the shortcut's]
    TcPlan tc_dual;
    bool has_dual;
    ncnn_cuda_conv2d_desc sdesc;
};

namespace {

struct Geom2
{
    int inw, inh, inch, outw, outh, outch, n;
};

static int geom_of(const ncnn_cuda_conv2d* conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, Geom2* g)
{
    TView b = make_view(bottom), t = make_view(top);
    if (bottom->dims >= 3)
    {
        g->inw = bottom->w;
        g->inh = bottom->h * (bottom->dims == 4 ? bottom->d : 1);
    }
    else
    {
        g->inw = 1;
        g->inh = b.P;
    }
    if (top->dims >= 3)
    {
        g->outw = top->w;
        g->outh = top->h * (top->dims == 4 ? top->d : 1);
    }
    else
    {
        g->outw = 1;
        g->outh = t.P;
    }
    g->inch = b.C;
    g->outch = t.C;
    g->n = b.n;
    if (g->inch != conv->desc.inch || g->outch != conv->desc.outch || t.n != b.n) return -1;
    return 0;
}

static void fill_call(const ncnn_cuda_conv2d* conv, const Geom2& g, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int pad_left, int pad_top,
                      const ncnn_cuda_tensor* residual, const ncnn_cuda_activation& act, TcConvCall* c)
{
    const ncnn_cuda_conv2d_desc& d = conv->desc;
    c->in = bottom->data;
    c->n = g.n;
    c->inh = g.inh;
    c->inw = g.inw;
    c->inch = g.inch;
    c->in_cpitch = bottom->cpitch;
    c->outh = g.outh;
    c->outw = g.outw;
    c->kernel_w = d.kernel_w;
    c->kernel_h = d.kernel_h;
    c->stride_w = d.stride_w;
    c->stride_h = d.stride_h;
    c->dil_w = d.dilation_w;
    c->dil_h = d.dilation_h;
    c->pad_left = pad_left;
    c->pad_top = pad_top;
    // effective far-side padding implied by the output size (may be negative when the stride leaves a remainder)
    c->pad_right = (g.outw - 1) * d.stride_w + (d.kernel_w - 1) * d.dilation_w + 1 - g.inw - pad_left;
    c->pad_bottom = (g.outh - 1) * d.stride_h + (d.kernel_h - 1) * d.dilation_h + 1 - g.inh - pad_top;
    c->out = top->data;
    c->out_cpitch = top->cpitch;
    c->residual = residual ? residual->data : 0;
    c->res_cpitch = residual ? residual->cpitch : 0;
    c->act_type = act.type;
    c->act_p0 = act.p0;
    c->act_p1 = act.p1;
    c->workspace = 0;
    c->workspace_size = 0;
    c->in2 = 0;
    c->in2_h = c->in2_w = c->in2_ch = c->in2_cpitch = 0;
    c->in2_stride_w = c->in2_stride_h = 1;
    c->tiled = (d.kernel_w == 1 && d.kernel_h == 1 && d.stride_w == 1 && d.stride_h == 1 && pad_left == 0 && pad_top == 0 && g.outw == g.inw && g.outh == g.inh) ? 1 : 0;
}

static bool dense(const ncnn_cuda_tensor* t)
{
    TView v = make_view(t);
    return v.n == 1 || v.nstep == (long long)v.P * v.cpitch;
}

// 0 SIMT, 1 tcgen05 tiled, 2 tcgen05 im2col
static int pick_algo(const ncnn_cuda_conv2d* conv, const Geom2& g, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int pad_left, int pad_top,
                     const ncnn_cuda_tensor* residual, TcConvCall* call)
{
    if (!conv->has_tc) return 0;
    if (bottom->elemtype != conv->desc.elemtype || top->elemtype != conv->desc.elemtype) return 0;
    if (!dense(bottom) || !dense(top) || (residual && !dense(residual))) return 0;
    if (conv->desc.pad_value != 0.f && (pad_left || pad_top || call->pad_right > 0 || call->pad_bottom > 0)) return 0;
    (void)g;
    if (!tc_conv_supported(&conv->tc, call)) return 0;
    return call->tiled ? 1 : 2;
}

} // namespace

extern "C" {

int ncnn_cuda_conv2d_create(ncnn_cuda_conv2d_t* out, const ncnn_cuda_conv2d_desc* desc, const float* weight, const float* bias, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    *out = 0;
    NC_REQUIRE(desc->inch > 0 && desc->outch > 0 && desc->kernel_w > 0 && desc->kernel_h > 0, "conv2d_create: bad shape");
    NC_REQUIRE(desc->stride_w > 0 && desc->stride_h > 0 && desc->dilation_w > 0 && desc->dilation_h > 0, "conv2d_create: bad stride/dilation");
    ncnn_cuda_conv2d* c = new ncnn_cuda_conv2d;
    memset(c, 0, sizeof(*c));
    c->desc = *desc;
    c->taps = desc->kernel_w * desc->kernel_h;
    c->K = c->taps * desc->inch;
    const int inch = desc->inch, outch = desc->outch, taps = c->taps;

    // tap-major, channel-innermost copy of the reference's [outch][inch][taps] weights
    std::vector<float> wt((size_t)outch * taps * inch);
    for (int oc = 0; oc < outch; oc++)
        for (int ci = 0; ci < inch; ci++)
        {
            const float* src = weight + ((size_t)oc * inch + ci) * taps;
            float* dst = wt.data() + (size_t)oc * taps * inch + ci;
            for (int t = 0; t < taps; t++) dst[(size_t)t * inch] = src[t];
        }

    int ret = 0;
    c->has_tc = false;
    if (desc->elemtype != NCNN_CUDA_F32 && tc_available())
    {
        // the A_ROWS stem variant needs the explicit left padding; SAME modes resolve per input size -> -1 disables it
        const int pad_left_known = (desc->pad_left >= 0 && desc->pad_right >= 0 && desc->pad_top >= 0 && desc->pad_bottom >= 0) ? desc->pad_left : -1;
        if (tc_plan_create(&c->tc, desc->elemtype, inch, outch, desc->kernel_w, desc->kernel_h, desc->stride_w, desc->dilation_w, pad_left_known, wt.data(),
                           desc->bias_term ? bias : 0, stream) == 0)
            c->has_tc = true;
    }

    // SIMT pack: [Kpad][wp_ld], k-major rows so a CTA's B tile is contiguous float4 loads
    {
        int Kpad = ((c->K + 15) / 16) * 16;
        c->wp_ld = ((outch + 127) / 128) * 128;
        std::vector<float> wp((size_t)Kpad * c->wp_ld, 0.f);
        for (int oc = 0; oc < outch; oc++)
        {
            const float* src = wt.data() + (size_t)oc * c->K;
            for (int k = 0; k < c->K; k++) wp[(size_t)k * c->wp_ld + oc] = src[k];
        }
        cudaError_t e = cudaMalloc((void**)&c->wp_simt, wp.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->wp_simt, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess && desc->bias_term && bias)
        {
            e = cudaMalloc((void**)&c->bias_dev, sizeof(float) * outch);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c->bias_dev, bias, sizeof(float) * outch, cudaMemcpyHostToDevice, stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess)
        {
            set_last_error("conv2d_create upload", e, __FILE__, __LINE__);
            ret = -100;
        }
    }
    if (ret != 0)
    {
        ncnn_cuda_conv2d_destroy(c);
        return ret;
    }
    *out = c;
    return 0;
}

int ncnn_cuda_conv2d_destroy(ncnn_cuda_conv2d_t c)
{
    if (!c) return 0;
    if (c->wp_simt) cudaFree(c->wp_simt);
    if (c->bias_dev) cudaFree(c->bias_dev);
    if (c->has_tc) tc_plan_destroy(&c->tc);
    if (c->has_dual) tc_plan_destroy(&c->tc_dual);
    delete c;
    return 0;
}

size_t ncnn_cuda_conv2d_workspace_size(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top)
{
    // only the small-channel stem variant keeps a (zero-padded, 4/8-channel) copy of its input; everything else is an
    // implicit GEMM straight from the blob
    if (!conv || !conv->has_tc || !bottom || !top || bottom->elemtype != conv->desc.elemtype) return 0;
    Geom2 g;
    if (geom_of(conv, bottom, top, &g) != 0) return 0;
    TcConvCall call;
    fill_call(conv, g, bottom, top, conv->desc.pad_left, conv->desc.pad_top, 0, conv->desc.act, &call);
    return tc_conv_workspace(&conv->tc, &call);
}

int ncnn_cuda_conv2d_algo(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom)
{
    return (conv->has_tc && bottom->elemtype == conv->desc.elemtype) ? 2 : 0;
}

int ncnn_cuda_conv2d_forward(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int pad_left, int pad_top,
                             const ncnn_cuda_tensor* residual, const ncnn_cuda_activation* act_override, void* workspace, size_t workspace_size, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && bottom && top && bottom->data && top->data, "conv2d_forward: null argument");
    Geom2 g;
    NC_REQUIRE(geom_of(conv, bottom, top, &g) == 0, "conv2d_forward: blob shape does not match the layer");
    NC_REQUIRE(bottom->elemtype == top->elemtype, "conv2d_forward: bottom/top element types differ");
    if (residual) NC_REQUIRE(same_shape(residual, top) && residual->elemtype == top->elemtype, "conv2d_forward: residual shape/type mismatch");
    const ncnn_cuda_activation act = act_override ? *act_override : conv->desc.act;

    TcConvCall call;
    fill_call(conv, g, bottom, top, pad_left, pad_top, residual, act, &call);
    call.workspace = workspace;
    call.workspace_size = workspace_size;
    int algo = pick_algo(conv, g, bottom, top, pad_left, pad_top, residual, &call);
    if (algo != 0)
    {
        int r = tc_conv_forward(&conv->tc, &call, stream);
        if (r == 0) return 0;
        if (r != -1) return r;
        // -1: descriptor could not be built for this geometry -> CUDA-core path below
    }

    ConvGeom cg;
    const ncnn_cuda_conv2d_desc& d = conv->desc;
    cg.inch = g.inch;
    cg.outch = g.outch;
    cg.kw = d.kernel_w;
    cg.kh = d.kernel_h;
    cg.dw = d.dilation_w;
    cg.dh = d.dilation_h;
    cg.sw = d.stride_w;
    cg.sh = d.stride_h;
    cg.pad_left = pad_left;
    cg.pad_top = pad_top;
    cg.pad_value = d.pad_value;
    cg.inw = g.inw;
    cg.inh = g.inh;
    cg.outw = g.outw;
    cg.outh = g.outh;
    cg.n = g.n;
    cg.in_cpitch = bottom->cpitch;
    cg.out_cpitch = top->cpitch;
    cg.res_cpitch = residual ? residual->cpitch : 0;
    cg.in_nstep = bottom->nstep;
    cg.out_nstep = top->nstep;
    cg.res_nstep = residual ? residual->nstep : 0;
    cg.K = conv->K;
    cg.wp_ld = conv->wp_ld;
    cg.act_type = act.type;
    cg.act_p0 = act.p0;
    cg.act_p1 = act.p1;
    const void* res = residual ? residual->data : 0;
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32:
        return launch_conv_simt<float>((const float*)bottom->data, conv->wp_simt, conv->bias_dev, (const float*)res, (float*)top->data, cg, stream);
    case NCNN_CUDA_BF16:
        return launch_conv_simt<__nv_bfloat16>((const __nv_bfloat16*)bottom->data, conv->wp_simt, conv->bias_dev, (const __nv_bfloat16*)res,
                                               (__nv_bfloat16*)top->data, cg, stream);
    case NCNN_CUDA_F16:
        return launch_conv_simt<__half>((const __half*)bottom->data, conv->wp_simt, conv->bias_dev, (const __half*)res, (__half*)top->data, cg, stream);
    }
    return -1;
}

// top = act(W * bottom + Ws * bottom2 + b + bs): a projection shortcut (1x1, possibly strided) folded into the 1x1 layer whose
// output it is added to -- one tcgen05 GEMM whose K axis runs over both inputs, fp32 accumulation across the sum; the shortcut's
// output blob never exists (ResNet-50 res2a: 411 MB written + 411 MB re-read per batch of 256 less)
int ncnn_cuda_conv2d_fuse_shortcut(ncnn_cuda_conv2d_t conv, const float* weight, const float* bias, const ncnn_cuda_conv2d_desc* sdesc, const float* sweight,
                                   const float* sbias, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && weight && sdesc && sweight, "conv2d_fuse_shortcut: null argument");
    if (conv->has_dual)
    {
        tc_plan_destroy(&conv->tc_dual);
        conv->has_dual = false;
    }
    const ncnn_cuda_conv2d_desc& d = conv->desc;
    if (!conv->has_tc) return -1; // fp32 storage: no tensor-core plan, the caller keeps the two layers apart
    const bool main_ok = d.kernel_w == 1 && d.kernel_h == 1 && d.stride_w == 1 && d.stride_h == 1 && d.pad_left == 0 && d.pad_right == 0 && d.pad_top == 0 && d.pad_bottom == 0;
    const bool sc_ok = sdesc->kernel_w == 1 && sdesc->kernel_h == 1 && sdesc->pad_left == 0 && sdesc->pad_right == 0 && sdesc->pad_top == 0 && sdesc->pad_bottom == 0 &&
                       sdesc->outch == d.outch && sdesc->elemtype == d.elemtype && sdesc->act.type == 0 && d.act.type == 0;
    if (!main_ok || !sc_ok || d.inch <= 32 || sdesc->inch <= 0) return -1;
    const int k1 = (d.inch + 63) / 64 * 64;
    const int K = k1 + sdesc->inch;
    std::vector<float> w((size_t)d.outch * K, 0.f), b((size_t)d.outch, 0.f);
    for (int oc = 0; oc < d.outch; oc++)
    {
        memcpy(w.data() + (size_t)oc * K, weight + (size_t)oc * d.inch, sizeof(float) * d.inch);
        memcpy(w.data() + (size_t)oc * K + k1, sweight + (size_t)oc * sdesc->inch, sizeof(float) * sdesc->inch);
        b[oc] = (d.bias_term && bias ? bias[oc] : 0.f) + (sdesc->bias_term && sbias ? sbias[oc] : 0.f);
    }
    if (tc_plan_create(&conv->tc_dual, d.elemtype, K, d.outch, 1, 1, 1, 1, 0, w.data(), b.data(), stream) != 0) return -1;
    if (conv->tc_dual.block_k != 64)
    {
        tc_plan_destroy(&conv->tc_dual);
        return -1;
    }
    conv->tc_dual.dual_k1_blocks = k1 / 64;
    conv->tc_dual.dual_inch2 = sdesc->inch;
    conv->sdesc = *sdesc;
    conv->has_dual = true;
    return 0;
}

int ncnn_cuda_conv2d_forward_shortcut(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* bottom2, const ncnn_cuda_tensor* top,
                                      const ncnn_cuda_activation* act, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && bottom && bottom2 && top && bottom->data && bottom2->data && top->data, "conv2d_forward_shortcut: null argument");
    if (!conv->has_dual) return -1;
    Geom2 g;
    NC_REQUIRE(geom_of(conv, bottom, top, &g) == 0, "conv2d_forward_shortcut: blob shape does not match the layer");
    if (bottom->elemtype != conv->desc.elemtype || bottom2->elemtype != conv->desc.elemtype || top->elemtype != conv->desc.elemtype) return -1;
    if (bottom2->dims != 3 || bottom->dims != 3 || !dense(bottom) || !dense(bottom2) || !dense(top)) return -1;
    TView b2 = make_view(bottom2);
    if (b2.C != conv->sdesc.inch || b2.n != g.n) return -1;
    ncnn_cuda_activation a;
    a.type = 0;
    a.p0 = a.p1 = 0.f;
    if (act) a = *act;
    TcConvCall call;
    fill_call(conv, g, bottom, top, 0, 0, 0, a, &call);
    call.in2 = bottom2->data;
    call.in2_w = bottom2->w;
    call.in2_h = bottom2->h;
    call.in2_ch = b2.C;
    call.in2_cpitch = bottom2->cpitch;
    call.in2_stride_w = conv->sdesc.stride_w;
    call.in2_stride_h = conv->sdesc.stride_h;
    if (!call.tiled || !tc_conv_supported(&conv->tc_dual, &call)) return -1;
    return tc_conv_forward(&conv->tc_dual, &call, stream);
}

// Convolution (+bias, ReLU) and the 3x3 stride-2 max pooling behind it in one kernel (stem_pool.cuh): small-channel stride-2 stems
// whose output row fits one 128-column tile.  `top` is the POOLED blob; the conv map (conv_outw x conv_outh) is never written.
static int maxpool_fold_call(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, int conv_outw, int conv_outh, int pad_left, int pad_top, int pw, int ph,
                             int pool_pad_left, int pool_pad_top, TcConvCall* call, TcPoolCall* pc)
{
    if (!conv || !bottom || !conv->has_tc || bottom->dims != 3) return -1;
    if (bottom->elemtype != conv->desc.elemtype || !dense(bottom) || conv->desc.pad_value != 0.f) return -1;
    // the conv map's shape on a stand-in tensor (never addressed)
    ncnn_cuda_tensor ct = *bottom;
    ct.w = conv_outw;
    ct.h = conv_outh;
    ct.c = conv->desc.outch;
    ct.cpitch = (conv->desc.outch + 7) & ~7;
    Geom2 g;
    if (geom_of(conv, bottom, &ct, &g) != 0) return -1;
    fill_call(conv, g, bottom, &ct, pad_left, pad_top, 0, conv->desc.act, call);
    pc->pad_left = pool_pad_left;
    pc->pad_top = pool_pad_top;
    pc->pw = pw;
    pc->ph = ph;
    pc->out = bottom->data; // (aligned stand-in for the support check; the caller sets the real blob)
    pc->out_cpitch = ct.cpitch;
    call->out = bottom->data;
    call->out_cpitch = ct.cpitch;
    return tc_stem_pool_supported(&conv->tc, call, pc) ? 0 : -1;
}

int ncnn_cuda_conv2d_maxpool3x3s2_supported(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, int conv_outw, int conv_outh, int pad_left, int pad_top, int pw, int ph,
                                            int pool_pad_left, int pool_pad_top)
{
    TcConvCall call;
    TcPoolCall pc;
    return maxpool_fold_call(conv, bottom, conv_outw, conv_outh, pad_left, pad_top, pw, ph, pool_pad_left, pool_pad_top, &call, &pc) == 0 ? 1 : 0;
}

int ncnn_cuda_conv2d_forward_maxpool3x3s2(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, int conv_outw, int conv_outh, int pad_left, int pad_top,
                                          const ncnn_cuda_tensor* top, int pool_pad_left, int pool_pad_top, void* workspace, size_t workspace_size, void* stream_)
{
    cudaStream_t stream = as_stream(stream_);
    NC_REQUIRE(conv && bottom && top && bottom->data && top->data, "conv2d_forward_maxpool: null argument");
    if (top->dims != 3 || top->elemtype != conv->desc.elemtype || !dense(top) || top->c != conv->desc.outch || top->n != bottom->n) return -1;
    TcConvCall call;
    TcPoolCall pc;
    if (maxpool_fold_call(conv, bottom, conv_outw, conv_outh, pad_left, pad_top, top->w, top->h, pool_pad_left, pool_pad_top, &call, &pc) != 0) return -1;
    call.workspace = workspace;
    call.workspace_size = workspace_size;
    pc.out = top->data;
    pc.out_cpitch = top->cpitch;
    if (!tc_stem_pool_supported(&conv->tc, &call, &pc)) return -1;
    return tc_stem_pool_forward(&conv->tc, &call, &pc, stream);
}

int ncnn_cuda_linear_create(ncnn_cuda_linear_t* fc, const ncnn_cuda_linear_desc* d, const float* weight, const float* bias, void* stream)
{
    ncnn_cuda_conv2d_desc cd;
    memset(&cd, 0, sizeof(cd));
    if (d->in_w > 0 && d->in_h > 0 && d->in_c > 0)
    {
        NC_REQUIRE(d->in_w * d->in_h * d->in_c == d->num_input, "linear_create: in_w*in_h*in_c != num_input");
        cd.inch = d->in_c;
        cd.kernel_w = d->in_w;
        cd.kernel_h = d->in_h;
    }
    else
    {
        cd.inch = d->num_input;
        cd.kernel_w = 1;
        cd.kernel_h = 1;
    }
    cd.outch = d->num_output;
    cd.dilation_w = cd.dilation_h = cd.stride_w = cd.stride_h = 1;
    cd.bias_term = d->bias_term;
    cd.act = d->act;
    cd.elemtype = d->elemtype;
    return ncnn_cuda_conv2d_create(fc, &cd, weight, bias, stream);
}

int ncnn_cuda_linear_destroy(ncnn_cuda_linear_t fc)
{
    return ncnn_cuda_conv2d_destroy(fc);
}

int ncnn_cuda_linear_forward(ncnn_cuda_linear_t fc, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream)
{
    return ncnn_cuda_conv2d_forward(fc, bottom, top, 0, 0, 0, 0, 0, 0, stream);
}

} // extern "C"
