// deconv_layers.cpp -- Deconvolution and DeconvolutionDepthWise.  Host side only: parameter ids,
// weight loading, the bordered-output size and the cut rules follow the reference (file:line cited per function);
// the arithmetic runs in csrc/cuda/deconv.cu behind ncnn_cuda_deconv2d_*.
#include "cuda_layers.h"

namespace ncnn {

Deconvolution::Deconvolution()
{
    one_blob_only = true;
    support_inplace = false;
    group = 1;
    handle = 0;
}

Deconvolution::~Deconvolution()
{
    if (handle) ncnn_cuda_deconv2d_destroy(handle);
}

// src/layer/deconvolution.cpp:16-46, src/layer/deconvolutiondepthwise.cpp:16-47 (id 7 = group there)
int Deconvolution::load_param(const ParamDict& pd)
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
    output_pad_right = pd.get(18, 0);
    output_pad_bottom = pd.get(19, output_pad_right);
    output_w = pd.get(20, 0);
    output_h = pd.get(21, output_w);
    bias_term = pd.get(5, 0);
    weight_data_size = pd.get(6, 0);
    group = reads_group() ? pd.get(7, 1) : 1;
    activation_type = pd.get(9, 0);
    activation_params = pd.get(10, Mat());
    dynamic_weight = pd.get(28, 0);
    if (dynamic_weight)
    {
        NCNN_LOGE("Deconvolution: dynamic_weight is not supported by the CUDA backend");
        return -1;
    }
    if (num_output <= 0 || kernel_w <= 0 || kernel_h <= 0 || stride_w <= 0 || stride_h <= 0 || dilation_w <= 0 || dilation_h <= 0) return -1;
    if (group <= 0 || num_output % group != 0) return -1;
    return 0;
}

// src/layer/deconvolution.cpp:48-66
int Deconvolution::load_model(const ModelBin& mb)
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

int Deconvolution::create_pipeline(const Option& opt)
{
    const int maxk = kernel_w * kernel_h;
    const int channels = (weight_data_size / group) / maxk / (num_output / group) * group; // as deconvolutiondepthwise_x86.cpp does
    if (channels <= 0 || (long long)channels / group * (num_output / group) * maxk * group != (long long)weight_data_size) return -1;
    ncnn_cuda_deconv2d_desc desc;
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
    desc.output_pad_right = output_pad_right;
    desc.output_pad_bottom = output_pad_bottom;
    desc.bias_term = bias_term;
    desc.act = make_activation(activation_type, activation_params);
    desc.elemtype = opt.cuda_elemtype();
    int ret = ncnn_cuda_deconv2d_create(&handle, &desc, (const float*)weight_data.data, bias_term ? (const float*)bias_data.data : 0, 0);
    if (ret != 0) return ret;
    if (opt.lightmode)
    {
        weight_data.release();
        bias_data.release();
    }
    return 0;
}

int Deconvolution::destroy_pipeline(const Option&)
{
    if (handle) ncnn_cuda_deconv2d_destroy(handle);
    handle = 0;
    return 0;
}

// forward: src/layer/deconvolution.cpp:146-180; the cut: cut_padding :364-392 (copy_cut_border removes
// top/bottom/left/right rows and columns of the bordered output)
int Deconvolution::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blob.dims != 3 || !handle) return -1;
    if (bottom_blob.c % group != 0) return -100;
    const int w = bottom_blob.w, h = bottom_blob.h;
    const int kernel_extent_w = dilation_w * (kernel_w - 1) + 1;
    const int kernel_extent_h = dilation_h * (kernel_h - 1) + 1;
    const int full_w = (w - 1) * stride_w + kernel_extent_w + output_pad_right;
    const int full_h = (h - 1) * stride_h + kernel_extent_h + output_pad_bottom;
    int cut_left = 0, cut_right = 0, cut_top = 0, cut_bottom = 0;
    if (pad_left > 0 || pad_right > 0 || pad_top > 0 || pad_bottom > 0)
    {
        cut_left = pad_left;
        cut_right = pad_right;
        cut_top = pad_top;
        cut_bottom = pad_bottom;
    }
    else if (output_w > 0 && output_h > 0)
    {
        const int wcut = full_w - output_w, hcut = full_h - output_h;
        if (pad_left == -233 || pad_right == -233 || pad_top == -233 || pad_bottom == -233)
        {
            cut_top = hcut / 2;
            cut_bottom = hcut - hcut / 2;
            cut_left = wcut / 2;
            cut_right = wcut - wcut / 2;
        }
        else if (pad_left == -234 || pad_right == -234 || pad_top == -234 || pad_bottom == -234)
        {
            cut_top = hcut - hcut / 2;
            cut_bottom = hcut / 2;
            cut_left = wcut - wcut / 2;
            cut_right = wcut / 2;
        }
        else
        {
            return -100; // the reference leaves top_blob empty here (neither SAME mode given) and reports -100
        }
    }
    if (cut_left < 0 || cut_right < 0 || cut_top < 0 || cut_bottom < 0) return -100;
    const int outw = full_w - cut_left - cut_right, outh = full_h - cut_top - cut_bottom;
    if (outw <= 0 || outh <= 0) return -100;
    top_blob.create(outw, outh, num_output, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_deconv2d_forward(handle, &b, &t, cut_left, cut_top, cmd.stream());
}

DeconvolutionDepthWise::DeconvolutionDepthWise()
{
}

} // namespace ncnn
