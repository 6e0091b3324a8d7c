#include "layer.h"

#include <string.h>

namespace ncnn {

Layer::Layer()
{
    one_blob_only = false;
    support_inplace = false;
    support_vulkan = false;
    support_packing = false;
    support_bf16_storage = true;
    support_fp16_storage = true;
    support_int8_storage = false;
    support_cuda = true;
    support_batch = true;
    featmask = 0;
    userdata = 0;
    typeindex = -1;
    top_count_hint = 1;
}

Layer::~Layer()
{
}

int Layer::load_param(const ParamDict&)
{
    return 0;
}
int Layer::load_model(const ModelBin&)
{
    return 0;
}
int Layer::create_pipeline(const Option&)
{
    return 0;
}
int Layer::destroy_pipeline(const Option&)
{
    return 0;
}

// ------------------------------------------------------------------ device overloads: defaults route between forms
int Layer::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    if (!support_inplace) return -1;
    top_blobs.resize(bottom_blobs.size());
    for (size_t i = 0; i < bottom_blobs.size(); i++)
    {
        int ret = cmd.record_clone(bottom_blobs[i], top_blobs[i], opt);
        if (ret != 0) return ret;
    }
    return forward_inplace(top_blobs, cmd, opt);
}

int Layer::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (!support_inplace) return -1;
    int ret = cmd.record_clone(bottom_blob, top_blob, opt);
    if (ret != 0) return ret;
    return forward_inplace(top_blob, cmd, opt);
}

int Layer::forward_inplace(std::vector<CudaMat>&, CudaCompute&, const Option&) const
{
    return -1;
}

int Layer::forward_inplace(CudaMat&, CudaCompute&, const Option&) const
{
    return -1;
}

// ------------------------------------------------------------------ host overloads
namespace {
struct ScopedContext
{
    CudaContext* ctx;
    explicit ScopedContext(const Option& opt)
        : ctx(acquire_cuda_context(opt.cuda_device_index))
    {
    }
    ~ScopedContext()
    {
        reclaim_cuda_context(ctx);
    }
};
} // namespace

int Layer::forward(const std::vector<Mat>& bottom_blobs, std::vector<Mat>& top_blobs, const Option& opt) const
{
    ScopedContext sc(opt);
    if (!sc.ctx)
    {
        NCNN_LOGE("layer %s: no CUDA device available (this runtime has no CPU compute path)", type.c_str());
        return -1;
    }
    CudaCompute cmd(sc.ctx);
    std::vector<CudaMat> b(bottom_blobs.size());
    for (size_t i = 0; i < bottom_blobs.size(); i++)
    {
        int ret = cmd.record_upload(bottom_blobs[i], b[i], opt);
        if (ret != 0) return ret;
    }
    std::vector<CudaMat> t(top_blobs.size() ? top_blobs.size() : (size_t)top_count_hint);
    int ret;
    if (one_blob_only && support_inplace)
    {
        ret = forward_inplace(b[0], cmd, opt);
        t.resize(1);
        t[0] = b[0];
    }
    else if (one_blob_only)
        ret = forward(b[0], t[0], cmd, opt);
    else if (support_inplace)
    {
        ret = forward_inplace(b, cmd, opt);
        t = b;
    }
    else
        ret = forward(b, t, cmd, opt);
    if (ret != 0) return ret;
    top_blobs.resize(t.size());
    for (size_t i = 0; i < t.size(); i++)
    {
        ret = cmd.record_download(t[i], top_blobs[i], opt);
        if (ret != 0) return ret;
    }
    return cmd.submit_and_wait();
}

int Layer::forward(const Mat& bottom_blob, Mat& top_blob, const Option& opt) const
{
    std::vector<Mat> b(1, bottom_blob), t(1);
    int ret = Layer::forward(b, t, opt);
    if (ret == 0) top_blob = t[0];
    return ret;
}

int Layer::forward_inplace(std::vector<Mat>& bottom_top_blobs, const Option& opt) const
{
    std::vector<Mat> t(bottom_top_blobs.size());
    int ret = Layer::forward(bottom_top_blobs, t, opt);
    if (ret == 0) bottom_top_blobs = t;
    return ret;
}

int Layer::forward_inplace(Mat& bottom_top_blob, const Option& opt) const
{
    Mat t;
    int ret = forward(bottom_top_blob, t, opt);
    if (ret == 0) bottom_top_blob = t;
    return ret;
}

ncnn_cuda_activation make_activation(int activation_type, const Mat& p)
{
    ncnn_cuda_activation a;
    a.type = activation_type;
    a.p0 = 0.f;
    a.p1 = 0.f;
    const float* v = (const float*)p.data;
    if (v && p.w >= 1) a.p0 = v[0];
    if (v && p.w >= 2) a.p1 = v[1];
    return a;
}

} // namespace ncnn
