// conv_layers.cpp -- Convolution, ConvolutionDepthWise, InnerProduct, Pooling, Gemm: the dense-contraction and
// bandwidth layers of the hot path.  Host side only: parameter ids, weight loading and shape rules follow the
// reference (file:line cited per function); the arithmetic runs in the sm_100a kernels behind include/ncnn_cuda.h.
#include "cuda_layers.h"

#include <mutex>

namespace ncnn {

// resolved (left, right, top, bottom) per Convolution::make_padding, src/layer/convolution.cpp:333-372
static void resolve_conv_pads(int w, int h, int kernel_w, int kernel_h, int dilation_w, int dilation_h, int stride_w, int stride_h, int pad_left, int pad_right,
                              int pad_top, int pad_bottom, int out[4])
{
    const int kernel_extent_w = dilation_w * (kernel_w - 1) + 1;
    const int kernel_extent_h = dilation_h * (kernel_h - 1) + 1;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (pad_left > 0 || pad_right > 0 || pad_top > 0 || pad_bottom > 0)
    {
        out[0] = pad_left;
        out[1] = pad_right;
        out[2] = pad_top;
        out[3] = pad_bottom;
    }
    else if (pad_left == -233 && pad_right == -233 && pad_top == -233 && pad_bottom == -233)
    {
        int wpad = kernel_extent_w + (w - 1) / stride_w * stride_w - w;
        int hpad = kernel_extent_h + (h - 1) / stride_h * stride_h - h;
        if (wpad > 0 || hpad > 0)
        {
            out[0] = wpad / 2;
            out[1] = wpad - wpad / 2;
            out[2] = hpad / 2;
            out[3] = hpad - hpad / 2;
        }
    }
    else if (pad_left == -234 && pad_right == -234 && pad_top == -234 && pad_bottom == -234)
    {
        int wpad = kernel_extent_w + (w - 1) / stride_w * stride_w - w;
        int hpad = kernel_extent_h + (h - 1) / stride_h * stride_h - h;
        if (wpad > 0 || hpad > 0)
        {
            out[0] = wpad - wpad / 2;
            out[1] = wpad / 2;
            out[2] = hpad - hpad / 2;
            out[3] = hpad / 2;
        }
    }
}

// ------------------------------------------------------------------------------------------------ Convolution
Convolution::Convolution()
{
    one_blob_only = true;
    support_inplace = false;
    fused_residual = false;
    fused_post_activation = -1;
    shortcut = 0;
    shortcut_fused = false;
    fused_pool = 0;
    handle = 0;
    handle_elemtype = -1;
}

Convolution::~Convolution()
{
    if (handle) ncnn_cuda_conv2d_destroy(handle);
    delete shortcut;
    delete fused_pool;
}

// src/layer/convolution.cpp:18-56
int Convolution::load_param(const ParamDict& pd)
{
    num_output = pd.get(0, 0);
    kernel_w = pd.get(1, 0);
    kernel_h = pd.get(11, kernel_w);
    dilation_w = pd.get(2, 1);
    dilation_h = pd.get(12, dilation_w);
    stride_w = pd.get(3, 1);
    stride_h = pd.get(13, stride_w);
    pad_left = pd.get(4, 0);
    pad_right = pd.get(15, pad_left);
    pad_top = pd.get(14, pad_left);
    pad_bottom = pd.get(16, pad_top);
    pad_value = pd.get(18, 0.f);
    bias_term = pd.get(5, 0);
    weight_data_size = pd.get(6, 0);
    int8_scale_term = pd.get(8, 0);
    activation_type = pd.get(9, 0);
    activation_params = pd.get(10, Mat());
    dynamic_weight = pd.get(19, 0);
    if (dynamic_weight)
    {
        NCNN_LOGE("Convolution: dynamic_weight is not supported by the CUDA backend");
        return -1;
    }
    if (int8_scale_term)
    {
        NCNN_LOGE("Convolution: int8 models are outside the CUDA backend's scope");
        return -1;
    }
    if (num_output <= 0 || kernel_w <= 0 || kernel_h <= 0 || weight_data_size <= 0 || weight_data_size % (num_output * kernel_w * kernel_h) != 0) return -1;
    return 0;
}

// src/layer/convolution.cpp:58-111
int Convolution::load_model(const ModelBin& mb)
{
    weight_data = mb.load(weight_data_size, 0);
    if (weight_data.empty()) return -100;
    if (bias_term)
    {
        bias_data = mb.load(num_output, 1);
        if (bias_data.empty()) return -100;
    }
    return 0;
}

int Convolution::create_pipeline(const Option& opt)
{
    if (handle)
    {
        ncnn_cuda_conv2d_destroy(handle);
        handle = 0;
    }
    ncnn_cuda_conv2d_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.outch = num_output;
    desc.inch = weight_data_size / (num_output * kernel_w * kernel_h);
    desc.kernel_w = kernel_w;
    desc.kernel_h = kernel_h;
    desc.dilation_w = dilation_w;
    desc.dilation_h = dilation_h;
    desc.stride_w = stride_w;
    desc.stride_h = stride_h;
    // SAME_UPPER / SAME_LOWER (-233 / -234) resolve per input size at forward time: the descriptor carries -1 ("not known yet"),
    // so the kernel plan does not build the fixed-padding stem variant for a padding the layer will never use
    const bool pads_known = pad_left >= 0 && pad_right >= 0 && pad_top >= 0 && pad_bottom >= 0;
    desc.pad_left = pads_known ? pad_left : -1;
    desc.pad_right = pads_known ? pad_right : -1;
    desc.pad_top = pads_known ? pad_top : -1;
    desc.pad_bottom = pads_known ? pad_bottom : -1;
    desc.pad_value = pad_value;
    desc.bias_term = bias_term;
    desc.act = make_activation(activation_type, activation_params);
    desc.elemtype = opt.cuda_elemtype();
    handle_elemtype = desc.elemtype;
    int ret = ncnn_cuda_conv2d_create(&handle, &desc, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
    if (ret != 0) return ret;
    shortcut_fused = false;
    if (shortcut)
    {
        ncnn_cuda_conv2d_desc sd;
        memset(&sd, 0, sizeof(sd));
        sd.outch = shortcut->num_output;
        sd.inch = shortcut->weight_data_size / shortcut->num_output;
        sd.kernel_w = sd.kernel_h = 1;
        sd.dilation_w = sd.dilation_h = 1;
        sd.stride_w = shortcut->stride_w;
        sd.stride_h = shortcut->stride_h;
        sd.bias_term = shortcut->bias_term;
        sd.elemtype = desc.elemtype;
        shortcut_fused = ncnn_cuda_conv2d_fuse_shortcut(handle, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, &sd,
                                                        (const float*)shortcut->weight_data.data, shortcut->bias_term ? (const float*)shortcut->bias_data.data : 0, 0) == 0;
        // the shortcut keeps its own pipeline as well: fp32 storage / small channel counts are not folded at kernel level, and a
        // blob the two-operand GEMM cannot address (a strided view) falls back to shortcut + residual at forward time
        ret = shortcut->create_pipeline(opt);
        if (ret != 0) return ret;
    }
    if (opt.lightmode)
    {
        weight_data.release();
        bias_data.release();
    }
    return 0;
}

int Convolution::destroy_pipeline(const Option& opt)
{
    if (handle) ncnn_cuda_conv2d_destroy(handle);
    handle = 0;
    if (shortcut) shortcut->destroy_pipeline(opt);
    return 0;
}

int Convolution::forward_impl(const CudaMat& bottom_blob, const CudaMat* residual, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (!handle) return -1;
    int w = bottom_blob.w, h = bottom_blob.h;
    if (bottom_blob.dims != 3)
    {
        NCNN_LOGE("Convolution %s: a 3-D bottom blob is required (use InnerProduct for flattened input)", name.c_str());
        return -1;
    }
    int pads[4];
    resolve_conv_pads(w, h, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h, pad_left, pad_right, pad_top, pad_bottom, pads);
    const int kernel_extent_w = dilation_w * (kernel_w - 1) + 1;
    const int kernel_extent_h = dilation_h * (kernel_h - 1) + 1;
    // src/layer/convolution.cpp:262-263
    const int outw = (w + pads[0] + pads[1] - kernel_extent_w) / stride_w + 1;
    const int outh = (h + pads[2] + pads[3] - kernel_extent_h) / stride_h + 1;
    if (outw <= 0 || outh <= 0) return -1;
    top_blob.create(outw, outh, num_output, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view(), r;
    if (residual) r = residual->view();
    ncnn_cuda_activation act;
    const ncnn_cuda_activation* actp = 0;
    if (residual)
    {
        // conv (no activation of its own) + residual, then the folded post activation
        act.type = fused_post_activation < 0 ? 0 : fused_post_activation;
        act.p0 = act.p1 = 0.f;
        actp = &act;
    }
    // scratch for the stem variant (zero-padded small-channel copy of the input); returned to the pool right after the
    // launch is enqueued -- reuse is ordered by the stream
    CudaMat workspace;
    size_t wsize = residual ? 0 : ncnn_cuda_conv2d_workspace_size(handle, &b, &t);
    if (wsize > 0)
    {
        workspace.create((int)((wsize + 3) / 4), NCNN_CUDA_F32, 1, cmd.workspace_allocator(opt));
        if (workspace.empty()) wsize = 0;
    }
    return ncnn_cuda_conv2d_forward(handle, &b, &t, pads[0], pads[2], residual ? &r : 0, actp, workspace.data, wsize, cmd.stream());
}

int Convolution::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (!fused_pool) return forward_impl(bottom_blob, 0, top_blob, cmd, opt);
    // stem fold: conv (+ReLU) + max pooling 3x3 s2 in one kernel when the geometry allows
    if (handle && bottom_blob.dims == 3)
    {
        const int w = bottom_blob.w, h = bottom_blob.h;
        int pads[4];
        resolve_conv_pads(w, h, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h, pad_left, pad_right, pad_top, pad_bottom, pads);
        const int cw = (w + pads[0] + pads[1] - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
        const int ch = (h + pads[2] + pads[3] - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
        int al, at, pw, ph;
        ncnn_cuda_tensor b = bottom_blob.view();
        if (cw > 0 && ch > 0 && fused_pool->window_geometry(cw, ch, al, at, pw, ph) &&
                ncnn_cuda_conv2d_maxpool3x3s2_supported(handle, &b, cw, ch, pads[0], pads[2], pw, ph, al, at))
        {
            top_blob.create(pw, ph, num_output, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
            if (top_blob.empty()) return -100;
            ncnn_cuda_tensor t = top_blob.view();
            ncnn_cuda_tensor ct = t;
            ct.w = cw;
            ct.h = ch;
            CudaMat workspace;
            size_t wsize = ncnn_cuda_conv2d_workspace_size(handle, &b, &ct);
            if (wsize > 0) workspace.create((int)((wsize + 3) / 4), NCNN_CUDA_F32, 1, cmd.workspace_allocator(opt));
            if (workspace.empty()) return -100;
            return ncnn_cuda_conv2d_forward_maxpool3x3s2(handle, &b, cw, ch, pads[0], pads[2], &t, al, at, workspace.data, wsize, cmd.stream());
        }
    }
    // two launches; the conv map is a scratch blob of THIS layer: never placed into a Concat buffer (opt may carry a placement
    // allocator meant for the pooled top)
    Option opt_scratch = opt;
    if (opt.blob_cuda_allocator) opt_scratch.blob_cuda_allocator = opt.blob_cuda_allocator->real();
    CudaMat conv_top;
    int r = forward_impl(bottom_blob, 0, conv_top, cmd, opt_scratch);
    if (r != 0) return r;
    return fused_pool->forward(conv_top, top_blob, cmd, opt);
}

int Convolution::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blobs.size() == 2 && shortcut)
    {
        const CudaMat& x = bottom_blobs[0];
        const CudaMat& x2 = bottom_blobs[1];
        if (shortcut_fused && handle && x.dims == 3 && x2.dims == 3)
        {
            CudaMat& top_blob = top_blobs[0];
            top_blob.create(x.w, x.h, num_output, x.elemtype, x.n, cmd.blob_allocator(opt));
            if (top_blob.empty()) return -100;
            ncnn_cuda_tensor b = x.view(), b2 = x2.view(), t = top_blob.view();
            ncnn_cuda_activation act;
            act.type = fused_post_activation < 0 ? 0 : fused_post_activation;
            act.p0 = act.p1 = 0.f;
            int r = ncnn_cuda_conv2d_forward_shortcut(handle, &b, &b2, &t, &act, cmd.stream());
            if (r != -1) return r;
            top_blob.release();
        }
        // two launches: the shortcut into a scratch blob, then this layer with that blob as the fused residual
        if (!shortcut->handle)
        {
            NCNN_LOGE("Convolution %s: folded shortcut %s has no pipeline for this blob layout", name.c_str(), shortcut->name.c_str());
            return -1;
        }
        Option opt_scratch = opt;
        if (opt.blob_cuda_allocator) opt_scratch.blob_cuda_allocator = opt.blob_cuda_allocator->real();
        CudaMat sc;
        int r = shortcut->forward_impl(x2, 0, sc, cmd, opt_scratch);
        if (r != 0) return r;
        return forward_impl(x, &sc, top_blobs[0], cmd, opt);
    }
    if (bottom_blobs.size() == 2 && fused_residual) return forward_impl(bottom_blobs[0], &bottom_blobs[1], top_blobs[0], cmd, opt);
    if (bottom_blobs.size() == 1) return forward_impl(bottom_blobs[0], 0, top_blobs[0], cmd, opt);
    return -1;
}

// ------------------------------------------------------------------------------------------------ ConvolutionDepthWise
ConvolutionDepthWise::ConvolutionDepthWise()
{
    one_blob_only = true;
    support_inplace = false;
    handle = 0;
    dense_handle = 0;
    handle_elemtype = -1;
}

ConvolutionDepthWise::~ConvolutionDepthWise()
{
    if (handle) ncnn_cuda_dwconv2d_destroy(handle);
    if (dense_handle) ncnn_cuda_conv2d_destroy(dense_handle);
}

// src/layer/convolutiondepthwise.cpp:18-60
int ConvolutionDepthWise::load_param(const ParamDict& pd)
{
    num_output = pd.get(0, 0);
    kernel_w = pd.get(1, 0);
    kernel_h = pd.get(11, kernel_w);
    dilation_w = pd.get(2, 1);
    dilation_h = pd.get(12, dilation_w);
    stride_w = pd.get(3, 1);
    stride_h = pd.get(13, stride_w);
    pad_left = pd.get(4, 0);
    pad_right = pd.get(15, pad_left);
    pad_top = pd.get(14, pad_left);
    pad_bottom = pd.get(16, pad_top);
    pad_value = pd.get(18, 0.f);
    bias_term = pd.get(5, 0);
    weight_data_size = pd.get(6, 0);
    group = pd.get(7, 1);
    int8_scale_term = pd.get(8, 0);
    activation_type = pd.get(9, 0);
    activation_params = pd.get(10, Mat());
    dynamic_weight = pd.get(19, 0);
    if (dynamic_weight || int8_scale_term)
    {
        NCNN_LOGE("ConvolutionDepthWise: dynamic_weight / int8 are not supported by the CUDA backend");
        return -1;
    }
    if (group <= 0 || num_output % group != 0) return -1; // reference: "num_output % group != 0" -> -100
    // a malformed param must fail here, not divide by zero in create_pipeline (weight_data_size / maxk) or loop forever in a kernel
    if (num_output <= 0 || kernel_w <= 0 || kernel_h <= 0 || stride_w <= 0 || stride_h <= 0 || dilation_w <= 0 || dilation_h <= 0 || weight_data_size <= 0)
    {
        NCNN_LOGE("ConvolutionDepthWise: invalid num_output / kernel / stride / dilation / weight_data_size");
        return -1;
    }
    return 0;
}

int ConvolutionDepthWise::load_model(const ModelBin& mb)
{
    weight_data = mb.load(weight_data_size, 0);
    if (weight_data.empty()) return -100;
    if (bias_term)
    {
        bias_data = mb.load(num_output, 1);
        if (bias_data.empty()) return -100;
    }
    return 0;
}

int ConvolutionDepthWise::create_pipeline(const Option& opt)
{
    const int maxk = kernel_w * kernel_h;
    const int channels = (weight_data_size / group) / maxk / (num_output / group) * group; // convolutiondepthwise.cpp:276
    handle_elemtype = opt.cuda_elemtype();
    int ret;
    if (group == 1)
    {
        ncnn_cuda_conv2d_desc desc;
        memset(&desc, 0, sizeof(desc));
        desc.outch = num_output;
        desc.inch = channels;
        desc.kernel_w = kernel_w;
        desc.kernel_h = kernel_h;
        desc.dilation_w = dilation_w;
        desc.dilation_h = dilation_h;
        desc.stride_w = stride_w;
        desc.stride_h = stride_h;
        desc.pad_value = pad_value;
        desc.bias_term = bias_term;
        desc.act = make_activation(activation_type, activation_params);
        desc.elemtype = handle_elemtype;
        ret = ncnn_cuda_conv2d_create(&dense_handle, &desc, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
    }
    else
    {
        ncnn_cuda_dwconv2d_desc desc;
        memset(&desc, 0, sizeof(desc));
        desc.inch = channels;
        desc.outch = num_output;
        desc.group = group;
        desc.kernel_w = kernel_w;
        desc.kernel_h = kernel_h;
        desc.dilation_w = dilation_w;
        desc.dilation_h = dilation_h;
        desc.stride_w = stride_w;
        desc.stride_h = stride_h;
        desc.pad_value = pad_value;
        desc.bias_term = bias_term;
        desc.act = make_activation(activation_type, activation_params);
        desc.elemtype = handle_elemtype;
        ret = ncnn_cuda_dwconv2d_create(&handle, &desc, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
    }
    if (ret != 0) return ret;
    if (opt.lightmode)
    {
        weight_data.release();
        bias_data.release();
    }
    return 0;
}

int ConvolutionDepthWise::destroy_pipeline(const Option&)
{
    if (handle) ncnn_cuda_dwconv2d_destroy(handle);
    if (dense_handle) ncnn_cuda_conv2d_destroy(dense_handle);
    handle = 0;
    dense_handle = 0;
    return 0;
}

int ConvolutionDepthWise::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blob.dims != 3) return -1;
    int w = bottom_blob.w, h = bottom_blob.h;
    if (bottom_blob.c % group != 0 || num_output % group != 0) return -100; // convolutiondepthwise.cpp:287-291
    int pads[4];
    resolve_conv_pads(w, h, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h, pad_left, pad_right, pad_top, pad_bottom, pads);
    const int kernel_extent_w = dilation_w * (kernel_w - 1) + 1;
    const int kernel_extent_h = dilation_h * (kernel_h - 1) + 1;
    const int outw = (w + pads[0] + pads[1] - kernel_extent_w) / stride_w + 1;
    const int outh = (h + pads[2] + pads[3] - kernel_extent_h) / stride_h + 1;
    if (outw <= 0 || outh <= 0) return -1;
    top_blob.create(outw, outh, num_output, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view();
    if (dense_handle) return ncnn_cuda_conv2d_forward(dense_handle, &b, &t, pads[0], pads[2], 0, 0, 0, 0, cmd.stream());
    if (!handle) return -1;
    return ncnn_cuda_dwconv2d_forward(handle, &b, &t, pads[0], pads[2], cmd.stream());
}

// ------------------------------------------------------------------------------------------------ InnerProduct
InnerProduct::InnerProduct()
{
    one_blob_only = true;
    support_inplace = false;
    elemtype = NCNN_CUDA_F32;
    pipes_lock = new std::mutex;
}

InnerProduct::~InnerProduct()
{
    for (size_t i = 0; i < pipes.size(); i++) ncnn_cuda_linear_destroy(pipes[i].handle);
    delete pipes_lock;
}

// src/layer/innerproduct.cpp:18-40
int InnerProduct::load_param(const ParamDict& pd)
{
    num_output = pd.get(0, 0);
    bias_term = pd.get(1, 0);
    weight_data_size = pd.get(2, 0);
    int8_scale_term = pd.get(8, 0);
    activation_type = pd.get(9, 0);
    activation_params = pd.get(10, Mat());
    if (int8_scale_term)
    {
        NCNN_LOGE("InnerProduct: int8 models are outside the CUDA backend's scope");
        return -1;
    }
    if (num_output <= 0 || weight_data_size <= 0 || weight_data_size % num_output != 0) return -1;
    return 0;
}

int InnerProduct::load_model(const ModelBin& mb)
{
    weight_data = mb.load(weight_data_size, 0);
    if (weight_data.empty()) return -100;
    if (bias_term)
    {
        bias_data = mb.load(num_output, 1);
        if (bias_data.empty()) return -100;
    }
    return 0;
}

int InnerProduct::create_pipeline(const Option& opt)
{
    elemtype = opt.cuda_elemtype();
    // shape hints (param id 30 -> bottom_shapes) let the packed weights be built at load time
    if (!bottom_shapes.empty() && bottom_shapes[0].dims == 3)
    {
        const Mat& s = bottom_shapes[0];
        if (s.w * s.h * s.c == weight_data_size / num_output)
        {
            Pipe p;
            p.in_w = s.w;
            p.in_h = s.h;
            p.in_c = s.c;
            ncnn_cuda_linear_desc d;
            memset(&d, 0, sizeof(d));
            d.num_input = weight_data_size / num_output;
            d.num_output = num_output;
            d.bias_term = bias_term;
            d.act = make_activation(activation_type, activation_params);
            d.elemtype = elemtype;
            d.in_w = s.w;
            d.in_h = s.h;
            d.in_c = s.c;
            if (s.w == 1 && s.h == 1) d.in_w = d.in_h = d.in_c = 0, p.in_w = p.in_h = p.in_c = 0;
            int ret = ncnn_cuda_linear_create(&p.handle, &d, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
            if (ret != 0) return ret;
            pipes.push_back(p);
            if (opt.lightmode)
            {
                weight_data.release();
                bias_data.release();
            }
        }
    }
    return 0;
}

int InnerProduct::destroy_pipeline(const Option&)
{
    for (size_t i = 0; i < pipes.size(); i++) ncnn_cuda_linear_destroy(pipes[i].handle);
    pipes.clear();
    return 0;
}

int InnerProduct::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    const int num_input = weight_data_size / num_output;
    CudaMat bottom = bottom_blob;
    int in_w = 0, in_h = 0, in_c = 0;
    bool rowwise = false;
    if (bottom.dims == 2 && bottom.w == num_input)
    {
        rowwise = true; // src/layer/innerproduct.cpp:102-134
    }
    else if (bottom.dims == 3 && bottom.w * bottom.h * bottom.c == num_input)
    {
        if (bottom.w * bottom.h > 1)
        {
            in_w = bottom.w;
            in_h = bottom.h;
            in_c = bottom.c;
        }
    }
    else if (bottom.dims == 1 && bottom.w == num_input)
    {
    }
    else
    {
        // any other rank with the right element count: flatten in the reference's logical order first
        if ((size_t)bottom.w * bottom.h * bottom.d * bottom.c != (size_t)num_input) return -1;
        CudaMat flat;
        flat.create(num_input, bottom.elemtype, bottom.n, cmd.blob_allocator(opt));
        if (flat.empty()) return -100;
        ncnn_cuda_tensor s = bottom.view(), f = flat.view();
        int ret = ncnn_cuda_reshape(&s, &f, cmd.stream());
        if (ret != 0) return ret;
        bottom = flat;
    }

    ncnn_cuda_linear_t handle = 0;
    {
        std::lock_guard<std::mutex> lk(*pipes_lock);
        for (size_t i = 0; i < pipes.size(); i++)
            if (pipes[i].in_w == in_w && pipes[i].in_h == in_h && pipes[i].in_c == in_c) handle = pipes[i].handle;
        if (!handle)
        {
            if (weight_data.empty())
            {
                NCNN_LOGE("InnerProduct %s: weights were released (lightmode) before the packed form for a %dx%dx%d bottom was built", name.c_str(), in_w, in_h, in_c);
                return -1;
            }
            Pipe p;
            p.in_w = in_w;
            p.in_h = in_h;
            p.in_c = in_c;
            ncnn_cuda_linear_desc d;
            memset(&d, 0, sizeof(d));
            d.num_input = num_input;
            d.num_output = num_output;
            d.bias_term = bias_term;
            d.act = make_activation(activation_type, activation_params);
            d.elemtype = bottom.elemtype;
            d.in_w = in_w;
            d.in_h = in_h;
            d.in_c = in_c;
            int ret = ncnn_cuda_linear_create(&p.handle, &d, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
            if (ret != 0) return ret;
            pipes.push_back(p);
            handle = p.handle;
        }
    }

    if (rowwise)
        top_blob.create(num_output, bottom.h, bottom.elemtype, bottom.n, cmd.blob_allocator(opt));
    else
        top_blob.create(num_output, bottom.elemtype, bottom.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom.view(), t = top_blob.view();
    return ncnn_cuda_linear_forward(handle, &b, &t, cmd.stream());
}

// ------------------------------------------------------------------------------------------------ Pooling
Pooling::Pooling()
{
    one_blob_only = true;
    support_inplace = false;
}

// src/layer/pooling.cpp:18-37
int Pooling::load_param(const ParamDict& pd)
{
    pooling_type = pd.get(0, 0);
    kernel_w = pd.get(1, 0);
    kernel_h = pd.get(11, kernel_w);
    stride_w = pd.get(2, 1);
    stride_h = pd.get(12, stride_w);
    pad_left = pd.get(3, 0);
    pad_right = pd.get(14, pad_left);
    pad_top = pd.get(13, pad_left);
    pad_bottom = pd.get(15, pad_top);
    global_pooling = pd.get(4, 0);
    pad_mode = pd.get(5, 0);
    avgpool_count_include_pad = pd.get(6, 0);
    adaptive_pooling = pd.get(7, 0);
    out_w = pd.get(8, 0);
    out_h = pd.get(18, out_w);
    if (pooling_type != 0 && pooling_type != 1) return -1;
    // windowed pooling needs a real window and step (the output-size formula divides by the stride); global and adaptive
    // pooling take their geometry from the blob
    if (!global_pooling && !adaptive_pooling && (kernel_w <= 0 || kernel_h <= 0 || stride_w <= 0 || stride_h <= 0))
    {
        NCNN_LOGE("Pooling: invalid kernel / stride");
        return -1;
    }
    if (adaptive_pooling && (out_w < 0 || out_h < 0)) return -1;
    return 0;
}

// make_padding, pooling.cpp:350-412: the effective pads of a windowed pooling over a (w, h) map (no padded copy is made)
void Pooling::resolve_pads(int w, int h, int& al, int& ar, int& at, int& ab, int& wtail, int& htail) const
{
    al = ar = at = ab = wtail = htail = 0;
    if (pad_mode == 0)
    {
        int wt = (w + pad_left + pad_right - kernel_w) % stride_w;
        int ht = (h + pad_top + pad_bottom - kernel_h) % stride_h;
        if (wt != 0) wtail = stride_w - wt;
        if (ht != 0) htail = stride_h - ht;
        al = pad_left;
        ar = pad_right + wtail;
        at = pad_top;
        ab = pad_bottom + htail;
    }
    else if (pad_mode == 1)
    {
        al = pad_left;
        ar = pad_right;
        at = pad_top;
        ab = pad_bottom;
    }
    else if (pad_mode == 2 || pad_mode == 3)
    {
        int wpad = kernel_w + (w - 1) / stride_w * stride_w - w;
        int hpad = kernel_h + (h - 1) / stride_h * stride_h - h;
        if (wpad > 0 || hpad > 0)
        {
            if (pad_mode == 2)
            {
                at = hpad / 2;
                ab = hpad - hpad / 2;
                al = wpad / 2;
                ar = wpad - wpad / 2;
            }
            else
            {
                at = hpad - hpad / 2;
                ab = hpad / 2;
                al = wpad - wpad / 2;
                ar = wpad / 2;
            }
        }
    }
}

// output size and leading pads of the window walk (pooling.cpp:197-198); false for global / adaptive pooling
bool Pooling::window_geometry(int w, int h, int& al, int& at, int& outw, int& outh) const
{
    if (global_pooling || adaptive_pooling) return false;
    int ar, ab, wtail, htail;
    resolve_pads(w, h, al, ar, at, ab, wtail, htail);
    outw = (w + al + ar - kernel_w) / stride_w + 1;
    outh = (h + at + ab - kernel_h) / stride_h + 1;
    return outw > 0 && outh > 0;
}

int Pooling::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blob.dims != 3) return -1;
    const int w = bottom_blob.w, h = bottom_blob.h, channels = bottom_blob.c;
    ncnn_cuda_pool2d_desc d;
    memset(&d, 0, sizeof(d));
    d.pooling_type = pooling_type;
    d.kernel_w = kernel_w;
    d.kernel_h = kernel_h;
    d.stride_w = stride_w;
    d.stride_h = stride_h;
    d.global_pooling = global_pooling;
    d.avgpool_count_include_pad = avgpool_count_include_pad;
    d.adaptive_pooling = adaptive_pooling;
    d.area_x0 = 0;
    d.area_x1 = w;
    d.area_y0 = 0;
    d.area_y1 = h;
    if (global_pooling)
    {
        top_blob.create(channels, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt)); // 1-D, pooling.cpp:52
    }
    else if (adaptive_pooling)
    {
        // pooling.cpp:96-100: -233 keeps the input extent
        int ow = out_w == -233 ? w : out_w;
        int oh = out_h == -233 ? h : out_h;
        if (ow == w && oh == h)
        {
            top_blob = bottom_blob;
            return 0;
        }
        top_blob.create(ow, oh, channels, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    }
    else
    {
        int al, ar, at, ab, wtail, htail;
        resolve_pads(w, h, al, ar, at, ab, wtail, htail);
        const int bw = w + al + ar, bh = h + at + ab;
        const int outw = (bw - kernel_w) / stride_w + 1; // pooling.cpp:197-198
        const int outh = (bh - kernel_h) / stride_h + 1;
        if (outw <= 0 || outh <= 0) return -1;
        d.pad_left = al;
        d.pad_top = at;
        // divisor region of avg without count_include_pad, pooling.cpp:283-300 (tests the MEMBER pads)
        d.area_x0 = pad_left - al;
        d.area_x1 = (bw - pad_right - (pad_mode == 0 ? wtail : 0)) - al;
        d.area_y0 = pad_top - at;
        d.area_y1 = (bh - pad_bottom - (pad_mode == 0 ? htail : 0)) - at;
        top_blob.create(outw, outh, channels, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    }
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_pool2d_forward(&d, &b, &t, cmd.stream());
}

} // namespace ncnn
