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

// ------------------------------------------------------------------ LRN (src/layer/lrn.cpp)
LRN::LRN()
{
    one_blob_only = true;
    support_inplace = true; // as the reference declares it; the device kernel writes a fresh blob and swaps it in
    region_type = 0;
    local_size = 5;
    alpha = 1.f;
    beta = 0.75f;
    bias = 1.f;
}

int LRN::load_param(const ParamDict& pd) // :14-24
{
    region_type = pd.get(0, 0);
    local_size = pd.get(1, 5);
    alpha = pd.get(2, 1.f);
    beta = pd.get(3, 0.75f);
    bias = pd.get(4, 1.f);
    return 0;
}

int LRN::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const
{
    // the window reads neighbours, so the kernel is out of place: a new blob takes the old one's place
    if (bottom_top_blob.dims != 3) return -1;
    CudaMat top_blob;
    top_blob.create_like(bottom_top_blob, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_top_blob.view(), t = top_blob.view();
    int ret = ncnn_cuda_lrn(&b, &t, region_type, local_size, alpha, beta, bias, cmd.stream());
    if (ret == 0) bottom_top_blob = top_blob;
    return ret;
}

// ------------------------------------------------------------------ LayerNorm (src/layer/layernorm.cpp)
LayerNorm::LayerNorm()
{
    one_blob_only = true;
    support_inplace = true;
}

// src/layer/layernorm.cpp:14-21
int LayerNorm::load_param(const ParamDict& pd)
{
    affine_size = pd.get(0, 0);
    eps = pd.get(1, 0.001f);
    affine = pd.get(2, 1);
    return 0;
}

// src/layer/layernorm.cpp:23-36
int LayerNorm::load_model(const ModelBin& mb)
{
    if (affine == 0) return 0;
    gamma_data = mb.load(affine_size, 1);
    if (gamma_data.empty()) return -100;
    beta_data = mb.load(affine_size, 1);
    if (beta_data.empty()) return -100;
    return 0;
}

int LayerNorm::create_pipeline(const Option&)
{
    if (affine == 0) return 0;
    int ret = upload_const(gamma_data, NCNN_CUDA_F32, gamma_dev);
    if (ret == 0) ret = upload_const(beta_data, NCNN_CUDA_F32, beta_dev);
    return ret;
}

int LayerNorm::destroy_pipeline(const Option&)
{
    gamma_dev.release();
    beta_dev.release();
    return 0;
}

// group selection: src/layer/layernorm.cpp:78-181 (w for 1-D / 2-D blobs; w, w*h or w*h*d by affine_size otherwise)
int LayerNorm::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const
{
    const CudaMat& m = bottom_top_blob;
    int size;
    if (m.dims <= 2)
        size = m.w;
    else if (m.dims == 3)
        size = affine_size == m.w ? m.w : m.w * m.h;
    else
        size = affine_size == m.w ? m.w : (affine_size == m.w * m.h ? m.w * m.h : m.w * m.h * m.d);
    if (affine && size != affine_size)
    {
        NCNN_LOGE("LayerNorm: affine_size %d does not match the normalised extent %d", affine_size, size);
        return -1;
    }
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_layernorm(&t, &t, size, eps, affine ? (const float*)gamma_dev.data : 0, affine ? (const float*)beta_dev.data : 0, cmd.stream());
}

// ------------------------------------------------------------------ Reduction (src/layer/reduction.cpp)
Reduction::Reduction()
{
    one_blob_only = true;
    support_inplace = false;
}

// src/layer/reduction.cpp:17-33
int Reduction::load_param(const ParamDict& pd)
{
    operation = pd.get(0, 0);
    reduce_all = pd.get(1, 1);
    coeff = pd.get(2, 1.f);
    axes = pd.get(3, Mat());
    keepdims = pd.get(4, 0);
    int fixbug0 = pd.get(5, 0);
    if (fixbug0 == 0 && !axes.empty())
    {
        NCNN_LOGE("param is too old, please regenerate!");
        return -1;
    }
    if (operation < 0 || operation > 10) return -1;
    return 0;
}

// flags and output shape: src/layer/reduction.cpp:753-856
int Reduction::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    const int dims = bottom_blob.dims;
    if (dims < 1 || dims > 4) return -1;
    bool reduce_w = false, reduce_h = false, reduce_d = false, reduce_c = false;
    if (reduce_all)
    {
        reduce_w = reduce_h = reduce_d = reduce_c = true;
    }
    else
    {
        int axes_flag[4] = {0, 0, 0, 0};
        const int* axes_ptr = (const int*)axes.data;
        for (int i = 0; i < axes.w; i++)
        {
            int axis = axes_ptr[i];
            if (axis < 0) axis += dims;
            if (axis < 0 || axis >= dims) return -1;
            axes_flag[axis] = 1;
        }
        if (dims == 1) reduce_w = true;
        if (dims == 2)
        {
            reduce_h = axes_flag[0] == 1;
            reduce_w = axes_flag[1] == 1;
        }
        if (dims == 3)
        {
            reduce_c = axes_flag[0] == 1;
            reduce_h = axes_flag[1] == 1;
            reduce_w = axes_flag[2] == 1;
        }
        if (dims == 4)
        {
            reduce_c = axes_flag[0] == 1;
            reduce_d = axes_flag[1] == 1;
            reduce_h = axes_flag[2] == 1;
            reduce_w = axes_flag[3] == 1;
        }
    }
    int outdims, outw = 1, outh = 1, outd = 1, outc = 1;
    if (keepdims)
    {
        outdims = dims;
        outw = reduce_w ? 1 : bottom_blob.w;
        outh = (dims >= 2 && !reduce_h) ? bottom_blob.h : 1;
        outd = (dims == 4 && !reduce_d) ? bottom_blob.d : 1;
        outc = (dims >= 3 && !reduce_c) ? bottom_blob.c : 1;
    }
    else
    {
        int shape[4], ns = 0;
        if (!reduce_w) shape[ns++] = bottom_blob.w;
        if (dims >= 2 && !reduce_h) shape[ns++] = bottom_blob.h;
        if (dims == 4 && !reduce_d) shape[ns++] = bottom_blob.d;
        if (dims >= 3 && !reduce_c) shape[ns++] = bottom_blob.c;
        outdims = ns == 0 ? 1 : ns; // a full reduction is a 1-element 1-D blob (:866-869)
        if (ns >= 1) outw = shape[0];
        if (ns >= 2) outh = shape[1];
        if (ns == 3) outc = shape[2];
        if (ns == 4)
        {
            outd = shape[2];
            outc = shape[3];
        }
    }
    top_blob.create_dims(outdims, outw, outh, outd, outc, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor b = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_reduction(operation, reduce_w, reduce_h, reduce_d, reduce_c, keepdims, coeff, &b, &t, cmd.stream());
}

// ------------------------------------------------------------------ MemoryData (src/layer/memorydata.cpp)
MemoryData::MemoryData()
{
    one_blob_only = false;
    support_inplace = false;
    w = h = d = c = 0;
    load_type = 1;
}

// src/layer/memorydata.cpp:15-24
int MemoryData::load_param(const ParamDict& pd)
{
    w = pd.get(0, 0);
    h = pd.get(1, 0);
    d = pd.get(11, 0);
    c = pd.get(2, 0);
    load_type = pd.get(21, 1);
    return 0;
}

// src/layer/memorydata.cpp:26-53
int MemoryData::load_model(const ModelBin& mb)
{
    if (d != 0)
        data = mb.load(w, h, d, c, load_type);
    else if (c != 0)
        data = mb.load(w, h, c, load_type);
    else if (h != 0)
        data = mb.load(w, h, load_type);
    else if (w != 0)
        data = mb.load(w, load_type);
    else
    {
        data.create(1);
        if (!data.empty()) ((float*)data.data)[0] = 0.f;
    }
    if (data.empty()) return -100;
    return 0;
}

int MemoryData::create_pipeline(const Option& opt)
{
    if (data.elemsize != 4u) return -1; // integer constants (load_type 4) are outside the float-only device blobs
    int ret = upload_const(data, opt.cuda_elemtype(), data_dev);
    if (ret != 0) return ret;
    if (opt.lightmode) data.release();
    return 0;
}

int MemoryData::destroy_pipeline(const Option&)
{
    data_dev.release();
    return 0;
}

// The reference clones the constant per forward (memorydata.cpp:55-64).  Device blobs are shared by reference count and
// the executor already clones a shared blob before handing it to an in-place layer (NetPrivate::do_forward_layer), so the
// constant is handed out as a reference: no copy, and it stays immutable.
int MemoryData::forward(const std::vector<CudaMat>&, std::vector<CudaMat>& top_blobs, CudaCompute&, const Option&) const
{
    if (data_dev.empty()) return -100;
    top_blobs[0] = data_dev;
    return 0;
}

// ------------------------------------------------------------------ Noop (src/layer/noop.cpp)
Noop::Noop()
{
    support_inplace = true;
}

int Noop::forward_inplace(std::vector<CudaMat>&, CudaCompute&, const Option&) const
{
    return 0;
}

int Noop::forward_inplace(CudaMat&, CudaCompute&, const Option&) const
{
    return 0;
}

// ------------------------------------------------------------------ Crop (src/layer/crop.cpp)
Crop::Crop()
{
    one_blob_only = true;
    support_inplace = false;
    woffset = hoffset = doffset = coffset = 0;
    outw = outh = outd = outc = 0;
    woffset2 = hoffset2 = doffset2 = coffset2 = 0;
}

int Crop::load_param(const ParamDict& pd) // :18-60
{
    woffset = pd.get(0, 0);
    hoffset = pd.get(1, 0);
    doffset = pd.get(13, 0);
    coffset = pd.get(2, 0);
    outw = pd.get(3, 0);
    outh = pd.get(4, 0);
    outd = pd.get(14, 0);
    outc = pd.get(5, 0);
    woffset2 = pd.get(6, 0);
    hoffset2 = pd.get(7, 0);
    doffset2 = pd.get(15, 0);
    coffset2 = pd.get(8, 0);
    starts = pd.get(9, Mat());
    ends = pd.get(10, Mat());
    axes = pd.get(11, Mat());
    return 0;
}

static inline void crop_numpy_axis(int extent, int start, int end, int& offset, int& out)
{
    if (start == -233) start = 0;
    if (end == -233) end = extent;
    offset = start >= 0 ? start : extent + start;
    int e = end > 0 ? end : extent + end;
    if (e > extent) e = extent;
    out = e - offset;
}

int Crop::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    // resolve_crop_roi, crop.cpp:390-590
    const int dims = bottom_blob.dims;
    const int w = bottom_blob.w, h = bottom_blob.h, d = bottom_blob.d, channels = bottom_blob.c;
    int _wo = 0, _ho = 0, _do = 0, _co = 0, _ow = w, _oh = h, _od = d, _oc = channels;
    if (!starts.empty() && !ends.empty())
    {
        const int* sp = (const int*)starts.data;
        const int* ep = (const int*)ends.data;
        const int* ap = (const int*)axes.data;
        int num_axis = axes.w;
        int ax[4] = {0, 1, 2, 3};
        if (num_axis == 0)
            num_axis = dims;
        else
            for (int i = 0; i < num_axis && i < 4; i++) ax[i] = ap[i] < 0 ? dims + ap[i] : ap[i];
        if (num_axis > 4 || num_axis > starts.w || num_axis > ends.w) return -1;
        for (int i = 0; i < num_axis; i++)
        {
            const int a = ax[i];
            // logical axis -> (extent, offset, out): dims 1: w; 2: h,w; 3: c,h,w; 4: c,d,h,w
            const int k = a + (4 - dims); // position in (c, d, h, w)
            if (a < 0 || a >= dims) return -1;
            if (dims == 2 && a == 0)
                crop_numpy_axis(h, sp[i], ep[i], _ho, _oh);
            else if (k == 3)
                crop_numpy_axis(w, sp[i], ep[i], _wo, _ow);
            else if (k == 2)
                crop_numpy_axis(h, sp[i], ep[i], _ho, _oh);
            else if (k == 1 && dims == 4)
                crop_numpy_axis(d, sp[i], ep[i], _do, _od);
            else
                crop_numpy_axis(channels, sp[i], ep[i], _co, _oc);
        }
    }
    else
    {
        _wo = woffset;
        _ho = hoffset;
        _do = doffset;
        _co = coffset;
        _ow = w - woffset - woffset2;
        if (outw != -233) _ow = outw < _ow ? outw : _ow;
        if (dims >= 2)
        {
            _oh = h - hoffset - hoffset2;
            if (outh != -233) _oh = outh < _oh ? outh : _oh;
        }
        if (dims >= 3)
        {
            _oc = channels - coffset - coffset2;
            if (outc != -233) _oc = outc < _oc ? outc : _oc;
        }
        if (dims == 4)
        {
            _od = d - doffset - doffset2;
            if (outd != -233) _od = outd < _od ? outd : _od;
        }
    }
    if (_ow <= 0 || _oh <= 0 || _od <= 0 || _oc <= 0 || _wo < 0 || _ho < 0 || _do < 0 || _co < 0) return -1;
    if (_wo + _ow > w || (dims >= 2 && _ho + _oh > h) || (dims == 4 && _do + _od > d) || (dims >= 3 && _co + _oc > channels)) return -1;
    if (_ow == w && _oh == h && _od == d && _oc == channels)
    {
        top_blob = bottom_blob; // crop.cpp:80-85: nothing to cut
        return 0;
    }
    // one single-axis copy per axis that is actually cut (ncnn_cuda_copy_from_axis); axis numbers in the blob's own rank
    struct Cut
    {
        int axis, offset, extent, full;
    };
    Cut cuts[4];
    int nc = 0;
    if (dims == 1)
        cuts[nc++] = Cut{0, _wo, _ow, w};
    else if (dims == 2)
    {
        cuts[nc++] = Cut{0, _ho, _oh, h};
        cuts[nc++] = Cut{1, _wo, _ow, w};
    }
    else if (dims == 3)
    {
        cuts[nc++] = Cut{0, _co, _oc, channels};
        cuts[nc++] = Cut{1, _ho, _oh, h};
        cuts[nc++] = Cut{2, _wo, _ow, w};
    }
    else
    {
        cuts[nc++] = Cut{0, _co, _oc, channels};
        cuts[nc++] = Cut{1, _do, _od, d};
        cuts[nc++] = Cut{2, _ho, _oh, h};
        cuts[nc++] = Cut{3, _wo, _ow, w};
    }
    CudaMat cur = bottom_blob;
    for (int i = 0; i < nc; i++)
    {
        if (cuts[i].extent == cuts[i].full) continue;
        int cw = cur.w, ch = cur.h, cd = cur.d, cc = cur.c;
        const int k = cuts[i].axis + (4 - dims); // position in (c, d, h, w)
        if (dims == 2 && cuts[i].axis == 0)
            ch = cuts[i].extent;
        else if (k == 3)
            cw = cuts[i].extent;
        else if (k == 2)
            ch = cuts[i].extent;
        else if (k == 1 && dims == 4)
            cd = cuts[i].extent;
        else
            cc = cuts[i].extent;
        CudaMat next;
        next.create_dims(dims, cw, ch, cd, cc, cur.elemtype, cur.n, cmd.blob_allocator(opt));
        if (next.empty()) return -100;
        ncnn_cuda_tensor b = cur.view(), t = next.view();
        int ret = ncnn_cuda_copy_from_axis(&b, &t, cuts[i].axis, cuts[i].offset, cmd.stream());
        if (ret != 0) return ret;
        cur = next;
    }
    top_blob = cur;
    return 0;
}

} // namespace ncnn
