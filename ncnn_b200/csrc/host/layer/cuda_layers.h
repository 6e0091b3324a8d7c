// cuda_layers.h -- the CUDA layer classes of the hot path.  One class per reference operator, same class name as
// the reference (src/layer/<op>.h) and the same parameter ids; each forward() only resolves shapes on the host and
// enqueues kernels through the C ABI (include/ncnn_cuda.h) on the recorder's stream.
#ifndef NCNN_B200_CUDA_LAYERS_H
#define NCNN_B200_CUDA_LAYERS_H

#include "../layer.h"

namespace ncnn {

class Input : public Layer
{
public:
    Input();
    virtual int load_param(const ParamDict& pd);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int w, h, d, c;
};

// src/layer/convolution.cpp
class Convolution : public Layer
{
public:
    Convolution();
    virtual ~Convolution();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    // bottom_blobs[1] is the fused residual when `fused_residual` (graph-level Conv+Eltwise(SUM)[+ReLU] fold)
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int forward_impl(const CudaMat& bottom_blob, const CudaMat* residual, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;

public:
    int num_output, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    int pad_left, pad_right, pad_top, pad_bottom;
    float pad_value;
    int bias_term, weight_data_size, int8_scale_term, activation_type;
    Mat activation_params;
    int dynamic_weight;
    Mat weight_data, bias_data;
    // graph-level fusion state (set by Net before create_pipeline)
    bool fused_residual;
    int fused_post_activation; // activation applied after the residual add (-1: none)
    // projection shortcut folded into this layer (graph-level Conv1x1(a) + Conv1x1(b) -> Eltwise(SUM)[+ReLU]): the other
    // Convolution, owned by this one; bottom_blobs[1] is ITS input.  `shortcut_fused` says the kernel-level fold succeeded
    // (one two-operand GEMM); otherwise forward runs the shortcut layer and adds its output as a residual.
    Convolution* shortcut;
    bool shortcut_fused;
    // max Pooling 3x3 s2 folded behind this (stem) convolution: owned here; forward produces the POOLED blob in one kernel when
    // the geometry allows, else runs the two layers one after the other
    class Pooling* fused_pool;
    ncnn_cuda_conv2d_t handle;
    int handle_elemtype;
};

// src/layer/convolutiondepthwise.cpp
class ConvolutionDepthWise : public Layer
{
public:
    ConvolutionDepthWise();
    virtual ~ConvolutionDepthWise();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;

public:
    int num_output, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    int pad_left, pad_right, pad_top, pad_bottom;
    float pad_value;
    int bias_term, weight_data_size, group, int8_scale_term, activation_type;
    Mat activation_params;
    int dynamic_weight;
    Mat weight_data, bias_data;
    ncnn_cuda_dwconv2d_t handle;
    ncnn_cuda_conv2d_t dense_handle; // group == 1 degenerates to Convolution
    int handle_elemtype;
};

// src/layer/deconvolution.cpp; `group` stays 1 here and is read from id 7 by DeconvolutionDepthWise
class Deconvolution : public Layer
{
public:
    Deconvolution();
    virtual ~Deconvolution();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    virtual bool reads_group() const { return false; }

public:
    int num_output, kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    int pad_left, pad_right, pad_top, pad_bottom;
    int output_pad_right, output_pad_bottom, output_w, output_h;
    int bias_term, weight_data_size, group, activation_type;
    Mat activation_params;
    int dynamic_weight;
    Mat weight_data, bias_data;
    ncnn_cuda_deconv2d_t handle;
};

// src/layer/deconvolutiondepthwise.cpp: same parameters plus id 7 = group; depthwise and grouped branches share one kernel
class DeconvolutionDepthWise : public Deconvolution
{
public:
    DeconvolutionDepthWise();
    virtual bool reads_group() const { return true; }
};

// src/layer/innerproduct.cpp
class InnerProduct : public Layer
{
public:
    InnerProduct();
    virtual ~InnerProduct();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;

public:
    int num_output, bias_term, weight_data_size, int8_scale_term, activation_type;
    Mat activation_params;
    Mat weight_data, bias_data;
    int elemtype;
    // the packed weights depend on the bottom's (w,h,c) factorisation; built on first use per shape
    struct Pipe
    {
        int in_w, in_h, in_c;
        ncnn_cuda_linear_t handle;
    };
    mutable std::vector<Pipe> pipes;
    mutable std::mutex* pipes_lock;
};

// src/layer/pooling.cpp
class Pooling : public Layer
{
public:
    Pooling();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    void resolve_pads(int w, int h, int& al, int& ar, int& at, int& ab, int& wtail, int& htail) const;
    bool window_geometry(int w, int h, int& al, int& at, int& outw, int& outh) const;

public:
    int pooling_type, kernel_w, kernel_h, stride_w, stride_h, pad_left, pad_right, pad_top, pad_bottom;
    int global_pooling, pad_mode, avgpool_count_include_pad, adaptive_pooling, out_w, out_h;
};

// src/layer/gemm.cpp
class Gemm : public Layer
{
public:
    Gemm();
    virtual ~Gemm();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;

public:
    float alpha, beta;
    int transA, transB, constantA, constantB, constantC, constantM, constantN, constantK, constant_broadcast_type_C;
    int output_N1M, output_elempack, output_elemtype, output_transpose;
    Mat A_data, B_data, C_data;
    CudaMat A_dev, B_dev, C_dev; // constants as device matrices (fp32 for C, storage type for A/B)
    ncnn_cuda_linear_t linear;   // tcgen05 path: rows(X) * W^T + b with the constant operand as W (see gemm_layer.cpp)
    int linear_mode;             // 0 none; 1 constant B: Y = A W^T (rows of A are the activations); 2 constant A: Y^T = B^T W^T
    int elemtype;
};

// ---- elementwise
class UnaryActivation : public Layer // base of ReLU / Sigmoid / Swish / TanH / Clip / HardSwish / HardSigmoid / Mish / Dropout
{
public:
    UnaryActivation();
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int op;
    float p0, p1;
    bool identity;
};

class ReLU : public UnaryActivation
{
public:
    virtual int load_param(const ParamDict& pd);
};
class Sigmoid : public UnaryActivation
{
public:
    Sigmoid();
};
class Swish : public UnaryActivation
{
public:
    Swish();
};
class TanH : public UnaryActivation
{
public:
    TanH();
};
class Mish : public UnaryActivation
{
public:
    Mish();
};
class Clip : public UnaryActivation
{
public:
    virtual int load_param(const ParamDict& pd);
};
class HardSwish : public UnaryActivation
{
public:
    virtual int load_param(const ParamDict& pd);
};
class HardSigmoid : public UnaryActivation
{
public:
    virtual int load_param(const ParamDict& pd);
};
class Dropout : public UnaryActivation
{
public:
    virtual int load_param(const ParamDict& pd);
};
class GELU : public UnaryActivation // src/layer/gelu.cpp
{
public:
    virtual int load_param(const ParamDict& pd);
};

// a constant host Mat (1-D / 2-D) as a device blob of the same logical shape, allocated from the weight allocator (gemm_layer.cpp)
int upload_const(const Mat& src, int elemtype, CudaMat& dst);

// src/layer/batchnorm.cpp -- inference form: value = b * value + a per channel (per row for 2-D blobs)
class BatchNorm : public Layer
{
public:
    BatchNorm();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int channels;
    float eps;
    Mat a_data, b_data;
    CudaMat a_dev, b_dev;
};

// src/layer/scale.cpp -- value * scale (+ bias) per channel; the two-input form (scale_data_size = -233) is not on the device path
class Scale : public Layer
{
public:
    Scale();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int scale_data_size, bias_term;
    Mat scale_data, bias_data;
    CudaMat scale_dev, bias_dev;
};

// src/layer/shufflechannel.cpp
class ShuffleChannel : public Layer
{
public:
    ShuffleChannel();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int group, reverse;
};

// src/layer/lrn.cpp
class LRN : public Layer
{
public:
    LRN();
    virtual int load_param(const ParamDict& pd);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int region_type, local_size;
    float alpha, beta, bias;
};

// src/layer/multiheadattention.cpp -- self / cross attention without mask, kv cache or quantised weights
class MultiHeadAttention : public Layer
{
public:
    MultiHeadAttention();
    virtual ~MultiHeadAttention();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int embed_dim, num_heads, weight_data_size, kdim, vdim, attn_mask, kv_cache;
    float scale;
    Mat q_weight_data, q_bias_data, k_weight_data, k_bias_data, v_weight_data, v_bias_data, out_weight_data, out_bias_data;
    ncnn_cuda_linear_t q_fc, k_fc, v_fc, o_fc;
};

// src/layer/layernorm.cpp
class LayerNorm : public Layer
{
public:
    LayerNorm();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int affine_size, affine;
    float eps;
    Mat gamma_data, beta_data;
    CudaMat gamma_dev, beta_dev;
};

// src/layer/reduction.cpp
class Reduction : public Layer
{
public:
    Reduction();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int operation, reduce_all, keepdims;
    float coeff;
    Mat axes;
};

// src/layer/memorydata.cpp -- a constant blob stored in the model file; uploaded once, handed out by reference
class MemoryData : public Layer
{
public:
    MemoryData();
    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int w, h, d, c, load_type;
    Mat data;
    CudaMat data_dev;
};

// src/layer/noop.cpp -- passes its blobs through
class Noop : public Layer
{
public:
    Noop();
    virtual int forward_inplace(std::vector<CudaMat>& bottom_top_blobs, CudaCompute& cmd, const Option& opt) const;
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
};

// src/layer/crop.cpp -- single-input forms: offsets (+ sizes / trailing offsets) and numpy-style starts / ends / axes
class Crop : public Layer
{
public:
    Crop();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int woffset, hoffset, doffset, coffset, outw, outh, outd, outc, woffset2, hoffset2, doffset2, coffset2;
    Mat starts, ends, axes;
};

class Eltwise : public Layer
{
public:
    Eltwise();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int op_type;
    Mat coeffs;
    bool fused_relu; // graph-level Eltwise+ReLU fold
};

class BinaryOp : public Layer
{
public:
    BinaryOp();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int op_type, with_scalar;
    float b;
};

// ---- data movement
class Split : public Layer
{
public:
    Split();
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
};

class Concat : public Layer
{
public:
    Concat();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int axis;
};

class Slice : public Layer
{
public:
    Slice();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    Mat slices, indices;
    int axis;
};

class Interp : public Layer
{
public:
    Interp();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    int resize_type;
    float height_scale, width_scale;
    int output_height, output_width, dynamic_target_size, align_corner;
};

class Softmax : public Layer
{
public:
    Softmax();
    virtual int load_param(const ParamDict& pd);
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;
    int axis;
};

class Reshape : public Layer
{
public:
    Reshape();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int w, h, d, c, ndim;
};

class Flatten : public Layer
{
public:
    Flatten();
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
};

class Permute : public Layer
{
public:
    Permute();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int order_type;
};

class Padding : public Layer
{
public:
    Padding();
    virtual int load_param(const ParamDict& pd);
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    int top, bottom, left, right, type, front, behind, per_channel_pad_data_size;
    float value;
};

} // namespace ncnn

#endif // NCNN_B200_CUDA_LAYERS_H
