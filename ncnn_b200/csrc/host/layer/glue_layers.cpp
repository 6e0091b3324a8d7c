// glue_layers.cpp -- the remaining operators the named models use (ReLU, Eltwise, BinaryOp, Split, Concat, Slice,
// Interp, Softmax, Reshape, Flatten, Permute, Padding, Dropout, Input, ...).  Parameter ids and shape rules follow the
// reference's src/layer/<op>.cpp (cited per class); kernels are behind include/ncnn_cuda.h.
#include "cuda_layers.h"

namespace ncnn {

// ------------------------------------------------------------------ Input (src/layer/input.cpp)
Input::Input()
{
    one_blob_only = true;
    support_inplace = true;
    w = h = d = c = 0;
}

int Input::load_param(const ParamDict& pd)
{
    w = pd.get(0, 0);
    h = pd.get(1, 0);
    d = pd.get(11, 0);
    c = pd.get(2, 0);
    return 0;
}

int Input::forward_inplace(CudaMat&, CudaCompute&, const Option&) const
{
    return 0;
}

// ------------------------------------------------------------------ unary activations
UnaryActivation::UnaryActivation()
{
    one_blob_only = true;
    support_inplace = true;
    op = NCNN_CUDA_UNARY_RELU;
    p0 = p1 = 0.f;
    identity = false;
}

int UnaryActivation::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const
{
    if (identity) return 0;
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_unary(op, p0, p1, &t, &t, cmd.stream());
}

// src/layer/relu.cpp:16-20
int ReLU::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_RELU;
    p0 = pd.get(0, 0.f);
    return 0;
}

Sigmoid::Sigmoid()
{
    op = NCNN_CUDA_UNARY_SIGMOID;
}
Swish::Swish()
{
    op = NCNN_CUDA_UNARY_SWISH;
}
TanH::TanH()
{
    op = NCNN_CUDA_UNARY_TANH;
}
Mish::Mish()
{
    op = NCNN_CUDA_UNARY_MISH;
}

// src/layer/clip.cpp:16-22
int Clip::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_CLIP;
    p0 = pd.get(0, -3.402823466e+38f);
    p1 = pd.get(1, 3.402823466e+38f);
    return 0;
}

// src/layer/hardswish.cpp:16-24
int HardSwish::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_HARDSWISH;
    p0 = pd.get(0, 0.2f);
    p1 = pd.get(1, 0.5f);
    return 0;
}

// src/layer/hardsigmoid.cpp:16-24
int HardSigmoid::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_HARDSIGMOID;
    p0 = pd.get(0, 0.2f);
    p1 = pd.get(1, 0.5f);
    return 0;
}

// src/layer/gelu.cpp:14-19
int GELU::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_GELU;
    p0 = pd.get(0, 0) ? 1.f : 0.f; // fast_gelu
    return 0;
}

// src/layer/dropout.cpp:14-40: identity unless scale != 1
int Dropout::load_param(const ParamDict& pd)
{
    op = NCNN_CUDA_UNARY_SCALE;
    p0 = pd.get(0, 1.f);
    identity = p0 == 1.f;
    return 0;
}

// ------------------------------------------------------------------ Eltwise (src/layer/eltwise.cpp:14-178)
Eltwise::Eltwise()
{
    one_blob_only = false;
    support_inplace = false;
    op_type = 0;
    fused_relu = false;
}

int Eltwise::load_param(const ParamDict& pd)
{
    op_type = pd.get(0, 0);
    coeffs = pd.get(1, Mat());
    return 0;
}

int Eltwise::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blobs.size() < 2) return -1;
    CudaMat& top = top_blobs[0];
    top.create_like(bottom_blobs[0], cmd.blob_allocator(opt));
    if (top.empty()) return -100;
    std::vector<ncnn_cuda_tensor> b(bottom_blobs.size());
    for (size_t i = 0; i < b.size(); i++) b[i] = bottom_blobs[i].view();
    ncnn_cuda_tensor t = top.view();
    const float* cf = (coeffs.w == (int)b.size()) ? (const float*)coeffs.data : 0;
    return ncnn_cuda_eltwise(op_type, &b[0], (int)b.size(), cf, fused_relu ? 1 : 0, &t, cmd.stream());
}

// ------------------------------------------------------------------ BinaryOp (src/layer/binaryop.cpp)
BinaryOp::BinaryOp()
{
    one_blob_only = false;
    support_inplace = false;
    op_type = 0;
    with_scalar = 0;
    b = 0.f;
}

int BinaryOp::load_param(const ParamDict& pd)
{
    op_type = pd.get(0, 0);
    with_scalar = pd.get(1, 0);
    b = pd.get(2, 0.f);
    if (with_scalar != 0)
    {
        one_blob_only = true;
        support_inplace = true;
    }
    return 0;
}

// output shape of numpy-style broadcasting over ncnn blobs (docs/developer-guide/binaryop-broadcasting.md): the blob
// with the higher rank (or, at equal rank, the larger extents) gives the shape; lower-rank blobs align to the OUTER
// axes in the reference (binaryop.cpp:60-140), which the kernel reproduces from the two logical shapes.
int BinaryOp::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blobs.size() != 2) return -1;
    const CudaMat& A = bottom_blobs[0];
    const CudaMat& B = bottom_blobs[1];
    const int outdims = A.dims > B.dims ? A.dims : B.dims;
    CudaMat A2 = A, B2 = B;
    // rank promotion exactly as binaryop.cpp:345-406: the lower-rank operand is reshaped to the output rank
    for (int which = 0; which < 2; which++)
    {
        const CudaMat& src = which == 0 ? A : B;
        const CudaMat& other = which == 0 ? B : A;
        CudaMat& lo = which == 0 ? A2 : B2;
        if (src.dims >= outdims) continue;
        int w = 1, h = 1, d = 1, c = 1;
        if (outdims == 2)
        {
            if (src.w == other.h)
                h = src.w;
            else
                w = src.w;
        }
        else if (outdims == 3 && src.dims == 1)
        {
            if (src.w == other.c)
                c = src.w;
            else
                w = src.w;
        }
        else if (outdims == 3 && src.dims == 2)
        {
            h = src.w;
            c = src.h;
        }
        else if (outdims == 4 && src.dims == 1)
        {
            if (src.w == other.c)
                c = src.w;
            else
                w = src.w;
        }
        else if (outdims == 4 && src.dims == 2)
        {
            d = src.w;
            c = src.h;
        }
        else if (outdims == 4 && src.dims == 3)
        {
            h = src.w;
            d = src.h;
            c = src.c;
        }
        CudaMat tmp;
        tmp.create_dims(outdims, w, h, d, c, src.elemtype, src.n, cmd.blob_allocator(opt));
        if (tmp.empty()) return -100;
        ncnn_cuda_tensor s = src.view(), t = tmp.view();
        int ret = ncnn_cuda_reshape(&s, &t, cmd.stream());
        if (ret != 0) return ret;
        lo = tmp;
    }
    auto bmax = [](int x, int y) { return x > y ? x : y; };
    const int ow = bmax(A2.w, B2.w), oh = bmax(A2.h, B2.h), od = bmax(A2.d, B2.d), oc = bmax(A2.c, B2.c);
    CudaMat& top = top_blobs[0];
    top.create_dims(outdims, ow, oh, od, oc, A.elemtype, A.n > B.n ? A.n : B.n, cmd.blob_allocator(opt));
    if (top.empty()) return -100;
    ncnn_cuda_tensor a = A2.view(), bb = B2.view(), t = top.view();
    return ncnn_cuda_binaryop(op_type, &a, &bb, 0.f, &t, cmd.stream());
}

int BinaryOp::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const
{
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_binaryop(op_type, &t, 0, b, &t, cmd.stream());
}

// ------------------------------------------------------------------ Split (src/layer/split.cpp:18-27): refcount share
Split::Split()
{
    one_blob_only = false;
    support_inplace = false;
}

int Split::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute&, const Option&) const
{
    for (size_t i = 0; i < top_blobs.size(); i++) top_blobs[i] = bottom_blobs[0];
    return 0;
}

// ------------------------------------------------------------------ Concat (src/layer/concat.cpp:14-292)
Concat::Concat()
{
    one_blob_only = false;
    support_inplace = false;
    axis = 0;
}

int Concat::load_param(const ParamDict& pd)
{
    axis = pd.get(0, 0);
    return 0;
}

static int axis_extent(const CudaMat& m, int positive_axis)
{
    // axis order of the reference for each rank: 1-D (w) 2-D (h,w) 3-D (c,h,w) 4-D (c,d,h,w)
    if (m.dims == 1) return m.w;
    if (m.dims == 2) return positive_axis == 0 ? m.h : m.w;
    if (m.dims == 3) return positive_axis == 0 ? m.c : (positive_axis == 1 ? m.h : m.w);
    return positive_axis == 0 ? m.c : (positive_axis == 1 ? m.d : (positive_axis == 2 ? m.h : m.w));
}

static void set_axis_extent(int dims, int positive_axis, int v, int& w, int& h, int& d, int& c)
{
    if (dims == 1) w = v;
    else if (dims == 2) (positive_axis == 0 ? h : w) = v;
    else if (dims == 3) (positive_axis == 0 ? c : (positive_axis == 1 ? h : w)) = v;
    else (positive_axis == 0 ? c : (positive_axis == 1 ? d : (positive_axis == 2 ? h : w))) = v;
}

int Concat::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    const CudaMat& b0 = bottom_blobs[0];
    const int dims = b0.dims;
    const int positive_axis = axis < 0 ? dims + axis : axis;
    if (positive_axis < 0 || positive_axis >= dims) return -1;
    int total = 0;
    for (size_t i = 0; i < bottom_blobs.size(); i++)
    {
        if (bottom_blobs[i].dims != dims) return -1;
        total += axis_extent(bottom_blobs[i], positive_axis);
    }
    int w = b0.w, h = b0.h, d = b0.d, c = b0.c;
    set_axis_extent(dims, positive_axis, total, w, h, d, c);
    CudaMat& top = top_blobs[0];
    // an unbatched bottom (a MemoryData constant such as a transformer's class token) is repeated for every sample of the batch
    int batch = b0.n;
    for (size_t i = 1; i < bottom_blobs.size(); i++)
        if (bottom_blobs[i].n > batch) batch = bottom_blobs[i].n;
    for (size_t i = 0; i < bottom_blobs.size(); i++)
        if (bottom_blobs[i].n != batch && bottom_blobs[i].n > 1) return -1;
    top.create_dims(dims, w, h, d, c, b0.elemtype, batch, cmd.blob_allocator(opt));
    if (top.empty()) return -100;
    ncnn_cuda_tensor t = top.view();
    int offset = 0;
    for (size_t i = 0; i < bottom_blobs.size(); i++)
    {
        ncnn_cuda_tensor b = bottom_blobs[i].view();
        int ret = ncnn_cuda_copy_into_axis(&b, &t, positive_axis, offset, cmd.stream());
        if (ret != 0) return ret;
        offset += axis_extent(bottom_blobs[i], positive_axis);
    }
    return 0;
}

// ------------------------------------------------------------------ Slice (src/layer/slice.cpp:14-405)
Slice::Slice()
{
    one_blob_only = false;
    support_inplace = false;
    axis = 0;
}

int Slice::load_param(const ParamDict& pd)
{
    slices = pd.get(0, Mat());
    axis = pd.get(1, 0);
    indices = pd.get(2, Mat());
    return 0;
}

int Slice::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    const CudaMat& bottom = bottom_blobs[0];
    const int dims = bottom.dims;
    const int positive_axis = axis < 0 ? dims + axis : axis;
    if (positive_axis < 0 || positive_axis >= dims) return -1;
    const int* slices_ptr = (const int*)slices.data;
    const int* indices_ptr = (const int*)indices.data;
    const int extent = axis_extent(bottom, positive_axis);
    ncnn_cuda_tensor b = bottom.view();
    int q = 0;
    for (size_t i = 0; i < top_blobs.size(); i++)
    {
        int slice;
        if (indices_ptr)
        {
            if (i == top_blobs.size() - 1)
                slice = extent - q;
            else
            {
                int indice = indices_ptr[i];
                int positive_indice = indice < 0 ? extent + indice : indice;
                slice = positive_indice - q;
            }
        }
        else
        {
            if (!slices_ptr || (int)i >= slices.w) return -1;
            slice = slices_ptr[i];
            if (slice == -233) slice = (int)((extent - q) / (top_blobs.size() - i));
        }
        if (slice <= 0 || q + slice > extent) return -1;
        CudaMat& top = top_blobs[i];
        if (opt.use_cuda_graph_fusion && dims == 3 && positive_axis == 0)
        {
            // channel ranges of a channel-innermost blob are VIEWS (same pixels, same pitch, data moved by q channels): no copy;
            // needs a 16-byte aligned start (every consumer addresses blobs through cpitch / nstep)
            top = bottom.channel_range(q, slice);
            if (!top.empty())
            {
                q += slice;
                continue;
            }
        }
        int w = bottom.w, h = bottom.h, d = bottom.d, c = bottom.c;
        set_axis_extent(dims, positive_axis, slice, w, h, d, c);
        top.create_dims(dims, w, h, d, c, bottom.elemtype, bottom.n, cmd.blob_allocator(opt));
        if (top.empty()) return -100;
        ncnn_cuda_tensor t = top.view();
        int ret = ncnn_cuda_copy_from_axis(&b, &t, positive_axis, q, cmd.stream());
        if (ret != 0) return ret;
        q += slice;
    }
    return 0;
}

// ------------------------------------------------------------------ Interp (src/layer/interp.cpp)
Interp::Interp()
{
    one_blob_only = true;
    support_inplace = false;
}

int Interp::load_param(const ParamDict& pd)
{
    resize_type = pd.get(0, 0);
    height_scale = pd.get(1, 1.f);
    width_scale = pd.get(2, 1.f);
    output_height = pd.get(3, 0);
    output_width = pd.get(4, 0);
    dynamic_target_size = pd.get(5, 0);
    align_corner = pd.get(6, 0);
    if (resize_type != 1 && resize_type != 2)
    {
        NCNN_LOGE("Interp: resize_type %d is not supported by the CUDA backend (nearest = 1, bilinear = 2)", resize_type);
        return -1;
    }
    if (dynamic_target_size == 1) one_blob_only = false;
    if (pd.type(9) == 7)
    {
        NCNN_LOGE("Interp: size_expr is not supported by the CUDA backend");
        return -1;
    }
    return 0;
}

static int interp_run(const Interp* self, const CudaMat& bottom, int outw, int outh, bool explicit_size, CudaMat& top, CudaCompute& cmd, const Option& opt)
{
    if (bottom.dims != 3)
    {
        NCNN_LOGE("Interp: a 3-D bottom blob is required");
        return -1;
    }
    const int w = bottom.w, h = bottom.h;
    if (outw == w && outh == h)
    {
        top = bottom; // interp.cpp:594-598
        return 0;
    }
    top.create(outw, outh, bottom.c, bottom.elemtype, bottom.n, cmd.blob_allocator(opt));
    if (top.empty()) return -100;
    // interp.cpp:606-607: the nearest-neighbour source step comes from the layer's OWN output_height / output_width members when they
    // are set and from its scale factors otherwise -- also when the target size was lent by a second bottom (dynamic_target_size),
    // where the reference therefore steps by 1 / scale (1 by default) and clamps; reproduced as is
    (void)explicit_size;
    const float hs = self->output_height ? h / (float)outh : 1.f / self->height_scale;
    const float ws = self->output_width ? w / (float)outw : 1.f / self->width_scale;
    ncnn_cuda_tensor b = bottom.view(), t = top.view();
    return ncnn_cuda_interp(self->resize_type, self->align_corner, hs, ws, &b, &t, cmd.stream());
}

int Interp::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    int outw = output_width, outh = output_height;
    if (outw == 0 || outh == 0)
    {
        // interp.cpp:439-443
        outw = (int)(bottom_blob.w * width_scale);
        outh = (int)(bottom_blob.h * height_scale);
    }
    return interp_run(this, bottom_blob, outw, outh, output_width != 0 && output_height != 0, top_blob, cmd, opt);
}

int Interp::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blobs.size() < 2) return forward(bottom_blobs[0], top_blobs[0], cmd, opt);
    // dynamic target size: the second blob only lends its w/h
    return interp_run(this, bottom_blobs[0], bottom_blobs[1].w, bottom_blobs[1].h, true, top_blobs[0], cmd, opt);
}

// ------------------------------------------------------------------ Softmax (src/layer/softmax.cpp:17-250)
Softmax::Softmax()
{
    one_blob_only = true;
    support_inplace = true;
    axis = 0;
}

int Softmax::load_param(const ParamDict& pd)
{
    axis = pd.get(0, 0);
    int fixbug0 = pd.get(1, 0);
    if (fixbug0 == 0 && axis != 0)
    {
        NCNN_LOGE("param is too old, please regenerate!");
        return -1;
    }
    return 0;
}

int Softmax::forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option&) const
{
    const int dims = bottom_top_blob.dims;
    const int positive_axis = axis < 0 ? dims + axis : axis;
    if (positive_axis < 0 || positive_axis >= dims) return -1;
    ncnn_cuda_tensor t = bottom_top_blob.view();
    return ncnn_cuda_softmax(&t, &t, positive_axis, cmd.stream());
}

// ------------------------------------------------------------------ Reshape (src/layer/reshape.cpp:22-217)
Reshape::Reshape()
{
    one_blob_only = true;
    support_inplace = false;
}

int Reshape::load_param(const ParamDict& pd)
{
    w = pd.get(0, -233);
    h = pd.get(1, -233);
    d = pd.get(11, -233);
    c = pd.get(2, -233);
    ndim = 4;
    if (d == -233) ndim = 3;
    if (c == -233) ndim = 2;
    if (h == -233) ndim = 1;
    if (w == -233) ndim = 0;
    if (pd.get(12, 233) != 233 || pd.get(13, 233) != 233)
    {
        NCNN_LOGE("Reshape: batch-axis reshape (ids 12/13) is not supported by the CUDA backend");
        return -1;
    }
    if (pd.type(6) == 7)
    {
        NCNN_LOGE("Reshape: shape_expr is not supported by the CUDA backend");
        return -1;
    }
    return 0;
}

int Reshape::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    int outw = w, outh = h, outd = d, outc = c;
    const int total = bottom_blob.w * bottom_blob.h * bottom_blob.d * bottom_blob.c;
    const int dims = bottom_blob.dims;
    if (ndim == 1)
    {
        if (outw == 0) outw = bottom_blob.w;
        if (outw == -1) outw = total;
        outh = outd = outc = 1;
    }
    else if (ndim == 2)
    {
        if (outw == 0) outw = bottom_blob.w;
        if (outh == 0) outh = bottom_blob.h;
        if (outw == -1) outw = total / outh;
        if (outh == -1) outh = total / outw;
        outd = outc = 1;
    }
    else if (ndim == 3)
    {
        if (outw == 0) outw = bottom_blob.w;
        if (outh == 0) outh = bottom_blob.h;
        if (outc == 0) outc = bottom_blob.c;
        if (outw == -1) outw = total / outc / outh;
        if (outh == -1) outh = total / outc / outw;
        if (outc == -1) outc = total / outh / outw;
        outd = 1;
    }
    else if (ndim == 4)
    {
        if (outw == 0) outw = bottom_blob.w;
        if (outh == 0) outh = bottom_blob.h;
        if (outc == 0) outc = bottom_blob.c;
        if (outd == 0) outd = bottom_blob.d;
        if (outw == -1) outw = total / outc / outd / outh;
        if (outh == -1) outh = total / outc / outd / outw;
        if (outd == -1) outd = total / outc / outh / outw;
        if (outc == -1) outc = total / outd / outh / outw;
    }
    else
        return -1;
    if ((long long)outw * outh * outd * outc != (long long)total) return -1;
    if (ndim == dims && outw == bottom_blob.w && outh == bottom_blob.h && outd == bottom_blob.d && outc == bottom_blob.c)
    {
        top_blob = bottom_blob;
        return 0;
    }
    top_blob.create_dims(ndim, outw, outh, outd, outc, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor s = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_reshape(&s, &t, cmd.stream());
}

// ------------------------------------------------------------------ Flatten (src/layer/flatten.cpp)
Flatten::Flatten()
{
    one_blob_only = true;
    support_inplace = false;
}

int Flatten::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (bottom_blob.dims == 1)
    {
        top_blob = bottom_blob;
        return 0;
    }
    const int total = bottom_blob.w * bottom_blob.h * bottom_blob.d * bottom_blob.c;
    top_blob.create(total, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor s = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_reshape(&s, &t, cmd.stream());
}

// ------------------------------------------------------------------ Permute (src/layer/permute.cpp:16-164)
Permute::Permute()
{
    one_blob_only = true;
    support_inplace = false;
    order_type = 0;
}

int Permute::load_param(const ParamDict& pd)
{
    order_type = pd.get(0, 0);
    return 0;
}

int Permute::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    const int dims = bottom_blob.dims;
    const int w = bottom_blob.w, h = bottom_blob.h, d = bottom_blob.d, c = bottom_blob.c;
    if (dims == 1 || order_type == 0)
    {
        top_blob = bottom_blob;
        return 0;
    }
    int ow = w, oh = h, od = d, oc = c;
    if (dims == 2)
    {
        // order_type 1: h w -> w h
        if (order_type != 1) return -1;
        ow = h;
        oh = w;
    }
    else if (dims == 3)
    {
        // permute.cpp:39-47: 0 = w h c, 1 = h w c, 2 = w c h, 3 = c w h, 4 = h c w, 5 = c h w  (fastest axis first)
        static const int table[6][3] = {{0, 1, 2}, {1, 0, 2}, {0, 2, 1}, {2, 0, 1}, {1, 2, 0}, {2, 1, 0}};
        if (order_type < 0 || order_type > 5) return -1;
        const int in[3] = {w, h, c};
        ow = in[table[order_type][0]];
        oh = in[table[order_type][1]];
        oc = in[table[order_type][2]];
    }
    else if (dims == 4)
    {
        if (order_type < 0 || order_type > 23) return -1;
        // permute.cpp:188-212: new (w h d c) named by the old axes, 0 = w, 1 = h, 2 = d, 3 = c
        static const int t4[24][4] = {
            {0, 1, 2, 3}, {1, 0, 2, 3}, {0, 2, 1, 3}, {2, 0, 1, 3}, {1, 2, 0, 3}, {2, 1, 0, 3}, {0, 1, 3, 2}, {1, 0, 3, 2},
            {0, 3, 1, 2}, {3, 0, 1, 2}, {1, 3, 0, 2}, {3, 1, 0, 2}, {0, 2, 3, 1}, {2, 0, 3, 1}, {0, 3, 2, 1}, {3, 0, 2, 1},
            {2, 3, 0, 1}, {3, 2, 0, 1}, {1, 2, 3, 0}, {2, 1, 3, 0}, {1, 3, 2, 0}, {3, 1, 2, 0}, {2, 3, 1, 0}, {3, 2, 1, 0}};
        const int* perm = t4[order_type];
        const int in[4] = {w, h, d, c};
        ow = in[perm[0]];
        oh = in[perm[1]];
        od = in[perm[2]];
        oc = in[perm[3]];
    }
    top_blob.create_dims(dims, ow, oh, od, oc, bottom_blob.elemtype, bottom_blob.n, cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor s = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_permute(&s, &t, order_type, cmd.stream());
}

// ------------------------------------------------------------------ Padding (src/layer/padding.cpp)
Padding::Padding()
{
    one_blob_only = true;
    support_inplace = false;
}

int Padding::load_param(const ParamDict& pd)
{
    top = pd.get(0, 0);
    bottom = pd.get(1, 0);
    left = pd.get(2, 0);
    right = pd.get(3, 0);
    type = pd.get(4, 0);
    value = pd.get(5, 0.f);
    per_channel_pad_data_size = pd.get(6, 0);
    front = pd.get(7, 0);
    behind = pd.get(8, 0);
    if (per_channel_pad_data_size)
    {
        NCNN_LOGE("Padding: per-channel pad values are not supported by the CUDA backend");
        return -1;
    }
    return 0;
}

int Padding::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    if (top == 0 && bottom == 0 && left == 0 && right == 0 && front == 0 && behind == 0)
    {
        top_blob = bottom_blob;
        return 0;
    }
    if (bottom_blob.dims != 3)
    {
        NCNN_LOGE("Padding: a 3-D bottom blob is required");
        return -1;
    }
    top_blob.create(bottom_blob.w + left + right, bottom_blob.h + top + bottom, bottom_blob.c + front + behind, bottom_blob.elemtype, bottom_blob.n,
                    cmd.blob_allocator(opt));
    if (top_blob.empty()) return -100;
    ncnn_cuda_tensor s = bottom_blob.view(), t = top_blob.view();
    return ncnn_cuda_padding(&s, &t, top, left, front, type, value, cmd.stream());
}

} // namespace ncnn
