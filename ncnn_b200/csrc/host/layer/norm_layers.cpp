// norm_layers.cpp -- BatchNorm, Scale, ShuffleChannel: the operators the un-fused / ShuffleNet-style models of the reference's
// benchmark set add on top of the five named graphs.  Parameter ids, weight order and arithmetic follow
// src/layer/batchnorm.cpp, scale.cpp, shufflechannel.cpp (cited per method); kernels are behind include/ncnn_cuda.h.
#include <math.h>

#include "cuda_layers.h"

namespace ncnn {

// ------------------------------------------------------------------ BatchNorm (src/layer/batchnorm.cpp)
BatchNorm::BatchNorm()
{
    one_blob_only = true;
    support_inplace = true;
    channels = 0;
    eps = 0.f;
}

int BatchNorm::load_param(const ParamDict& pd) // :14-20
{
    channels = pd.get(0, 0);
    eps = pd.get(1, 0.f);
    return 0;
}

int BatchNorm::load_model(const ModelBin& mb) // :22-55: slope, mean, var, bias -> a = bias - slope*mean/sqrt(var+eps), b = slope/sqrt(var+eps)
{
    Mat slope = mb.load(channels, 1);
    Mat mean = mb.load(channels, 1);
    Mat var = mb.load(channels, 1);
    Mat bias = mb.load(channels, 1);
    if (slope.empty() || mean.empty() || var.empty() || bias.empty()) return -100;
    a_data.create(channels);
    b_data.create(channels);
    if (a_data.empty() || b_data.empty()) return -100;
    for (int i = 0; i < channels; i++)
    {
        float sqrt_var = sqrtf(var[i] + eps);
        if (sqrt_var == 0.f) sqrt_var = 0.0001f; // the reference's divide-by-zero guard
        a_data[i] = bias[i] - slope[i] * mean[i] / sqrt_var;
        b_data[i] = slope[i] / sqrt_var;
    }
    return 0;
}

int BatchNorm::create_pipeline(const Option&)
{
    int ret = upload_const(a_data, NCNN_CUDA_F32, a_dev);
    if (ret == 0) ret = upload_const(b_data, NCNN_CUDA_F32, b_dev);
    return ret;
}

int BatchNorm::destroy_pipeline(const Option&)
{
    a_dev.release();
    b_dev.release();
    return 0;
}

int BatchNorm::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const // :57-120
{
    const int count = bottom_top_blob.dims == 1 ? bottom_top_blob.w : (bottom_top_blob.dims == 2 ? bottom_top_blob.h : bottom_top_blob.c);
    if (count != channels)
    {
        NCNN_LOGE("BatchNorm: blob has %d channels, layer has %d", count, channels);
        return -1;
    }
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_channel_affine(&t, &t, (const float*)b_dev.data, (const float*)a_dev.data, cmd.stream());
}

// ------------------------------------------------------------------ Scale (src/layer/scale.cpp)
Scale::Scale()
{
    one_blob_only = true;
    support_inplace = true;
    scale_data_size = 0;
    bias_term = 0;
}

int Scale::load_param(const ParamDict& pd) // :14-23
{
    scale_data_size = pd.get(0, 0);
    bias_term = pd.get(1, 0);
    if (scale_data_size == -233)
    {
        NCNN_LOGE("Scale: the two-input form (scale from a second blob) is not supported on the CUDA path");
        return -1;
    }
    return 0;
}

int Scale::load_model(const ModelBin& mb) // :25-42
{
    scale_data = mb.load(scale_data_size, 1);
    if (scale_data.empty()) return -100;
    if (bias_term)
    {
        bias_data = mb.load(scale_data_size, 1);
        if (bias_data.empty()) return -100;
    }
    return 0;
}

int Scale::create_pipeline(const Option&)
{
    int ret = upload_const(scale_data, NCNN_CUDA_F32, scale_dev);
    if (ret == 0 && bias_term) ret = upload_const(bias_data, NCNN_CUDA_F32, bias_dev);
    return ret;
}

int Scale::destroy_pipeline(const Option&)
{
    scale_dev.release();
    bias_dev.release();
    return 0;
}

int Scale::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const // :44-168
{
    const int count = bottom_top_blob.dims == 1 ? bottom_top_blob.w : (bottom_top_blob.dims == 2 ? bottom_top_blob.h : bottom_top_blob.c);
    if (count != scale_data_size)
    {
        NCNN_LOGE("Scale: blob has %d channels, layer has %d", count, scale_data_size);
        return -1;
    }
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_channel_affine(&t, &t, (const float*)scale_dev.data, bias_term ? (const float*)bias_dev.data : 0, cmd.stream());
}

// ------------------------------------------------------------------ ShuffleChannel (src/layer/shufflechannel.cpp)
ShuffleChannel::ShuffleChannel()
{
    one_blob_only = true;
    support_inplace = false;
    group = 1;
    reverse = 0;
}

int ShuffleChannel::load_param(const ParamDict& pd) // :14-20
{
    group = pd.get(0, 1);
    reverse = pd.get(1, 0);
    return 0;
}

int ShuffleChannel::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const // :22-60
{
    const int channels = bottom_blob.c;
    if (bottom_blob.dims < 3 || group <= 0 || channels % group != 0) return -100; // "reject invalid group"
    const int g = reverse ? channels / group : group;
    top_blob.create_like(bottom_blob, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_shuffle_channel(&b, &t, g, cmd.stream());
}

} // namespace ncnn
