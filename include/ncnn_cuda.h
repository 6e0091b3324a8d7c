/* ncnn_cuda.h -- the thin C ABI between the host C++ runtime (ncnn_b200/csrc/host) and the
 * hand-written sm_100a kernels (ncnn_b200/csrc/cuda).
 *
 * Everything here is `extern "C"`, POD arguments only, `cudaStream_t` travels as `void*`,
 * no C++ types and no exceptions cross the boundary.  Return codes follow the reference's
 * convention (SURVEY.md 8b "errors"): 0 ok, -1 invalid / unsupported argument, -100 out of
 * memory or CUDA failure (reference: src/layer/convolution.cpp:64-71, src/net.cpp:641-642).
 *
 * What each group replaces in the reference (Tencent/ncnn @ a4d2ea1d):
 *   device/memory/stream ... VulkanDevice + VkAllocator::fastMalloc/fastFree (src/allocator.h:267-296)
 *                            and VkCompute::record_upload/record_download/submit_and_wait
 *                            (src/command.h:22-88)
 *   ncnn_cuda_tensor ........ VkMat (src/mat.h:387-555): a non-owning POD view of a device blob
 *   conv2d/dwconv2d/...  .... the bodies of Layer::forward for the hot-path operators
 *                            (src/layer/<op>.cpp, file:line cited per entry point)
 *
 * DEVICE LAYOUT (private to this backend; the host Mat layout of src/mat.h is converted at
 * upload/download): a blob is [n][P][cpitch] "pixels x channels", channels innermost:
 *     dims 1 (w)        : P = 1,     C = w
 *     dims 2 (w,h)      : P = h,     C = w
 *     dims 3 (w,h,c)    : P = h*w,   C = c      (NHWC)
 *     dims 4 (w,h,d,c)  : P = d*h*w, C = c
 * element (b, pixel p, channel q) lives at data[b*nstep + p*cpitch + q]; cpitch >= C
 * (16-byte multiple for 16-bit types so that TMA can address rows), nstep >= P*cpitch.
 * Lanes q in [C, cpitch) are padding: kernels never rely on their value.
 */
#ifndef NCNN_CUDA_H
#define NCNN_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NCNN_CUDA_API __attribute__((visibility("default")))
#else
#define NCNN_CUDA_API
#endif

/* element types of device blobs */
#define NCNN_CUDA_F32  0
#define NCNN_CUDA_BF16 1
#define NCNN_CUDA_F16  2

typedef struct ncnn_cuda_tensor
{
    void* data;       /* device pointer */
    int dims;         /* 1..4, as ncnn::Mat::dims */
    int w, h, d, c;   /* logical ncnn shape */
    int n;            /* batch (ncnn::Mat::n, src/mat.h:373-381) */
    int elemtype;     /* NCNN_CUDA_F32 / BF16 / F16 */
    int cpitch;       /* elements between consecutive pixels */
    long long nstep;  /* elements between consecutive samples */
} ncnn_cuda_tensor;

/* host Mat view used by upload/download (reference layout, src/mat.cpp:299-861):
 * planar, element (b,q,z,y,x) at data[b*nstep + q*cstep + (z*h+y)*w + x], fp32 */
typedef struct ncnn_cuda_hostmat
{
    void* data;       /* host OR device pointer to fp32 planar data */
    int dims;
    int w, h, d, c, n;
    long long cstep;  /* elements */
    long long nstep;  /* elements */
} ncnn_cuda_hostmat;

/* ------------------------------------------------------------------ device / memory / stream */
NCNN_CUDA_API int ncnn_cuda_device_count(void);
NCNN_CUDA_API int ncnn_cuda_set_device(int index);
NCNN_CUDA_API int ncnn_cuda_get_device(void);
/* name: >= 256 bytes. Any out pointer may be NULL. */
NCNN_CUDA_API int ncnn_cuda_device_info(int index, char* name, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
NCNN_CUDA_API const char* ncnn_cuda_last_error(void);

NCNN_CUDA_API int ncnn_cuda_malloc(void** ptr, size_t size);
NCNN_CUDA_API int ncnn_cuda_free(void* ptr);
NCNN_CUDA_API int ncnn_cuda_malloc_host(void** ptr, size_t size); /* pinned */
NCNN_CUDA_API int ncnn_cuda_free_host(void* ptr);
/* 1 when `ptr` is page-locked host memory (cudaMallocHost / cudaHostRegister): async copies need no staging */
NCNN_CUDA_API int ncnn_cuda_host_is_pinned(const void* ptr);
NCNN_CUDA_API int ncnn_cuda_memcpy_h2d_async(void* dst, const void* src, size_t size, void* stream);
NCNN_CUDA_API int ncnn_cuda_memcpy_d2h_async(void* dst, const void* src, size_t size, void* stream);
NCNN_CUDA_API int ncnn_cuda_memcpy_d2d_async(void* dst, const void* src, size_t size, void* stream);
NCNN_CUDA_API int ncnn_cuda_memset_async(void* dst, int value, size_t size, void* stream);

NCNN_CUDA_API int ncnn_cuda_stream_create(void** stream);
NCNN_CUDA_API int ncnn_cuda_stream_destroy(void* stream);
NCNN_CUDA_API int ncnn_cuda_stream_sync(void* stream);
NCNN_CUDA_API int ncnn_cuda_device_sync(void);

NCNN_CUDA_API int ncnn_cuda_event_create(void** event);
NCNN_CUDA_API int ncnn_cuda_event_destroy(void* event);
NCNN_CUDA_API int ncnn_cuda_event_record(void* event, void* stream);
NCNN_CUDA_API int ncnn_cuda_event_sync(void* event);
NCNN_CUDA_API int ncnn_cuda_event_elapsed_ms(void* start, void* stop, float* ms);

/* CUDA graph capture of one recorded forward walk (the analogue of re-submitting a recorded
 * VkCompute command buffer, src/command.cpp:1834) */
NCNN_CUDA_API int ncnn_cuda_graph_begin_capture(void* stream);
NCNN_CUDA_API int ncnn_cuda_graph_end_capture(void* stream, void** graph_exec);
NCNN_CUDA_API int ncnn_cuda_graph_launch(void* graph_exec, void* stream);
NCNN_CUDA_API int ncnn_cuda_graph_destroy(void* graph_exec);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
NCNN_CUDA_API unsigned long long ncnn_cuda_launch_count(void);
/* of those, launches of the tcgen05 implicit-GEMM kernel (tests: which Convolution / InnerProduct / Gemm forms reach the tensor cores) */
NCNN_CUDA_API unsigned long long ncnn_cuda_tc_launch_count(void);

/* ------------------------------------------------------------------ layout conversion
 * pack: planar fp32 (host layout, but `src->data` must be DEVICE memory: the caller has
 * already copied the raw Mat bytes) -> device blob layout/dtype.  unpack: the inverse.
 * Replaces VkCompute::record_upload/record_download + Packing/Cast (src/command.cpp:358,439). */
NCNN_CUDA_API int ncnn_cuda_pack_from_planar(const ncnn_cuda_hostmat* src, const ncnn_cuda_tensor* dst, void* stream);
NCNN_CUDA_API int ncnn_cuda_unpack_to_planar(const ncnn_cuda_tensor* src, const ncnn_cuda_hostmat* dst, void* stream);
/* general re-layout: dst's logical element order (ncnn dense order n, c, d, h, w) is taken from
 * src's logical order: covers Reshape / Flatten (src/layer/reshape.cpp, flatten.cpp), dtype casts
 * and clone (VkCompute::record_clone). Total logical element counts per sample must match. */
/* Device pre-processing: Mat::from_pixels (src/mat_pixel.cpp) + Mat::substract_mean_normalize (src/mat.cpp)
 * in one kernel.  `pixels_dev`: n images of h rows x `stride` bytes of interleaved 8-bit pixels with `channels` (1, 3, 4)
 * bytes each, `nstride` bytes apart, ALREADY on the device; swap_rb reverses the first three channels (PIXEL_RGB2BGR /
 * PIXEL_BGR2RGB).  top(w, h, channels)[c] = (pixel[c] - mean[c]) * norm[c]; mean_vals / norm_vals are HOST arrays of
 * `channels` floats or NULL (0 / 1). */
NCNN_CUDA_API int ncnn_cuda_pixels_to_blob(const unsigned char* pixels_dev, int channels, int w, int h, int stride, long long nstride, int swap_rb, const float* mean_vals,
                                           const float* norm_vals, const ncnn_cuda_tensor* top, void* stream);

/* The same with the reference's bilinear resize in front: Mat::from_pixels_resize (src/mat_pixel.cpp:2546-2578), i.e.
 * resize_bilinear_c1/c3/c4 (src/mat_pixel_resize.cpp:210-1039, 11-bit integer coefficients) on the 8-bit image, then from_pixels
 * and substract_mean_normalize, one thread per output pixel, bit-exact with the reference's integer arithmetic.
 * ncnn_cuda_resize_tables (host only, no device work) fills 3 * (w + h) ints -- source column / row offsets and the short
 * coefficients exactly as :599-669 compute them -- which the caller copies to the device as `tables_dev`.
 * top: (w, h, channels) blob of the target size; the source must be at least 2 x 2. */
NCNN_CUDA_API int ncnn_cuda_resize_tables_count(int w, int h);
NCNN_CUDA_API int ncnn_cuda_resize_tables(int src_w, int src_h, int w, int h, int* tables_host);
NCNN_CUDA_API int ncnn_cuda_pixels_resize_to_blob(const unsigned char* pixels_dev, int channels, int src_w, int src_h, int stride, long long nstride, int swap_rb,
                                                  const float* mean_vals, const float* norm_vals, const int* tables_dev, const ncnn_cuda_tensor* top, void* stream);

NCNN_CUDA_API int ncnn_cuda_reshape(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, void* stream);
/* Permute (src/layer/permute.cpp:16-164): order_type as in the reference */
NCNN_CUDA_API int ncnn_cuda_permute(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, int order_type, void* stream);

/* ------------------------------------------------------------------ fused activation
 * activation_type / params exactly as src/layer/fused_activation.h:10-64:
 * 0 none, 1 relu, 2 leakyrelu(slope), 3 clip(min,max), 4 sigmoid, 5 mish, 6 hardswish(alpha,beta) */
typedef struct ncnn_cuda_activation
{
    int type;
    float p0, p1;
} ncnn_cuda_activation;

/* ------------------------------------------------------------------ Convolution
 * src/layer/convolution.cpp:113-184 (+ make_padding :328-372).  Weights arrive in the
 * reference order [outch][inch][kh][kw] fp32 (convolution.cpp:159) and are re-packed ONCE
 * (the create_pipeline step, as convolution_x86.cpp:279-500 / convolution_vulkan.cpp do). */
typedef struct ncnn_cuda_conv2d_desc
{
    int inch, outch;
    int kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    int pad_left, pad_right, pad_top, pad_bottom; /* the layer's fixed padding, or -1 when it is resolved per call (SAME_UPPER / SAME_LOWER
                                                    * depend on the input size: the host passes the resolved pads to forward) */
    float pad_value;
    int bias_term;
    ncnn_cuda_activation act;
    int elemtype; /* compute/storage type of the blobs this pipeline will see */
} ncnn_cuda_conv2d_desc;

typedef struct ncnn_cuda_conv2d* ncnn_cuda_conv2d_t;

NCNN_CUDA_API int ncnn_cuda_conv2d_create(ncnn_cuda_conv2d_t* conv, const ncnn_cuda_conv2d_desc* desc, const float* weight_host, const float* bias_host, void* stream);
NCNN_CUDA_API int ncnn_cuda_conv2d_destroy(ncnn_cuda_conv2d_t conv);
/* pads may differ per call (SAME modes depend on the input size); pass desc pads for fixed padding.
 * `residual` (may be NULL): fused `top = act(conv + bias + residual)`, the Eltwise-SUM(+ReLU) fold
 * (tools/ncnnoptimize.cpp-style fusion done at load time). */
NCNN_CUDA_API int ncnn_cuda_conv2d_forward(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top,
                                           int pad_left, int pad_top, const ncnn_cuda_tensor* residual, const ncnn_cuda_activation* act_override,
                                           void* workspace, size_t workspace_size, void* stream);
/* Projection-shortcut fold (load-time fusion; the offline tool fuses only activations, tools/ncnnoptimize.cpp:1268-1419):
 * `top = act(conv(bottom) + shortcut(bottom2))` where both are 1x1 / pad 0 / no activation and the shortcut may be strided
 * (src/layer/convolution.cpp:113-184 twice + src/layer/eltwise.cpp SUM): ONE GEMM whose K axis runs over both inputs.
 * fuse_shortcut takes the reference-order fp32 weights of both layers; returns -1 when the pair cannot be folded (the caller
 * keeps two layers), as does forward_shortcut for blobs it cannot address (the caller falls back to conv + residual). */
NCNN_CUDA_API int ncnn_cuda_conv2d_fuse_shortcut(ncnn_cuda_conv2d_t conv, const float* weight_host, const float* bias_host, const ncnn_cuda_conv2d_desc* shortcut_desc,
                                                 const float* shortcut_weight_host, const float* shortcut_bias_host, void* stream);
NCNN_CUDA_API int ncnn_cuda_conv2d_forward_shortcut(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* bottom2, const ncnn_cuda_tensor* top,
                                                    const ncnn_cuda_activation* act, void* stream);
/* Stem fold (load-time fusion): Convolution (+bias, ReLU) followed by max Pooling 3x3 stride 2 (src/layer/convolution.cpp:113-184 +
 * src/layer/pooling.cpp:188-253) in ONE kernel for small-channel stride-2 stems (ResNet conv1 7x7, SqueezeNet conv1 3x3): the
 * full-resolution conv map never reaches HBM.  `top` is the pooled blob, pool_pad_* the window's leading pads (trailing windows
 * are clipped to the map, which is what the reference's -FLT_MAX border does); workspace as for `forward`
 * (ncnn_cuda_conv2d_workspace_size with the conv map's shape).  Returns -1 when the pair cannot be fused (caller runs two layers). */
NCNN_CUDA_API int ncnn_cuda_conv2d_maxpool3x3s2_supported(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, int conv_outw, int conv_outh, int pad_left, int pad_top,
                                                          int pw, int ph, int pool_pad_left, int pool_pad_top);
NCNN_CUDA_API int ncnn_cuda_conv2d_forward_maxpool3x3s2(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, int conv_outw, int conv_outh, int pad_left, int pad_top,
                                                        const ncnn_cuda_tensor* top, int pool_pad_left, int pool_pad_top, void* workspace, size_t workspace_size,
                                                        void* stream);
/* bytes of scratch `forward` needs for this input shape (0 for the implicit-GEMM paths) */
NCNN_CUDA_API size_t ncnn_cuda_conv2d_workspace_size(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top);
/* which kernel family a call would use: 0 SIMT fp32, 1 tcgen05 GEMM (1x1), 2 tcgen05 implicit GEMM (TMA im2col), 3 tcgen05 + explicit im2col */
NCNN_CUDA_API int ncnn_cuda_conv2d_algo(ncnn_cuda_conv2d_t conv, const ncnn_cuda_tensor* bottom);

/* ------------------------------------------------------------------ ConvolutionDepthWise
 * src/layer/convolutiondepthwise.cpp:146-270.  group == channels == num_output is the
 * depthwise branch (:181-214); other groupings run the grouped branch (:216-267).
 * weights [group][outch_g][inch_g][kh][kw]. */
typedef struct ncnn_cuda_dwconv2d_desc
{
    int inch, outch, group;
    int kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    float pad_value;
    int bias_term;
    ncnn_cuda_activation act;
    int elemtype;
} ncnn_cuda_dwconv2d_desc;

typedef struct ncnn_cuda_dwconv2d* ncnn_cuda_dwconv2d_t;

NCNN_CUDA_API int ncnn_cuda_dwconv2d_create(ncnn_cuda_dwconv2d_t* conv, const ncnn_cuda_dwconv2d_desc* desc, const float* weight_host, const float* bias_host, void* stream);
NCNN_CUDA_API int ncnn_cuda_dwconv2d_destroy(ncnn_cuda_dwconv2d_t conv);
NCNN_CUDA_API int ncnn_cuda_dwconv2d_forward(ncnn_cuda_dwconv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top,
                                             int pad_left, int pad_top, void* stream);

/* ------------------------------------------------------------------ Deconvolution / DeconvolutionDepthWise
 * src/layer/deconvolution.cpp:68-146 and src/layer/deconvolutiondepthwise.cpp:68-208 (group == 1 is Deconvolution).
 * weights [group][outch_g][inch_g][kh][kw] fp32 as the reference's load_model keeps them, re-packed once.
 * The reference computes a bordered output of (w-1)*stride + dilation*(kernel-1) + 1 + output_pad per axis and then
 * cuts the pads (cut_padding, deconvolution.cpp:364-392); here the host resolves the cut (pad_left/pad_top, or the
 * SAME_UPPER / SAME_LOWER split of the difference to output_w/output_h) and `top` is the already-cut blob:
 * top[y][x] = bordered[y + cut_top][x + cut_left]. */
typedef struct ncnn_cuda_deconv2d_desc
{
    int inch, outch, group;
    int kernel_w, kernel_h, dilation_w, dilation_h, stride_w, stride_h;
    int output_pad_right, output_pad_bottom;
    int bias_term;
    ncnn_cuda_activation act;
    int elemtype;
} ncnn_cuda_deconv2d_desc;

typedef struct ncnn_cuda_deconv2d* ncnn_cuda_deconv2d_t;

NCNN_CUDA_API int ncnn_cuda_deconv2d_create(ncnn_cuda_deconv2d_t* conv, const ncnn_cuda_deconv2d_desc* desc, const float* weight_host, const float* bias_host, void* stream);
NCNN_CUDA_API int ncnn_cuda_deconv2d_destroy(ncnn_cuda_deconv2d_t conv);
NCNN_CUDA_API int ncnn_cuda_deconv2d_forward(ncnn_cuda_deconv2d_t conv, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top,
                                             int cut_left, int cut_top, void* stream);

/* ------------------------------------------------------------------ Pooling
 * src/layer/pooling.cpp:39-348 (+ make_padding :350-412).  The host resolves pad_mode into
 * pad_left/pad_top and the output size; the kernel treats everything outside the real input as
 * padding (max: ignored; avg: excluded from the divisor unless count_include_pad, in which case
 * the divisor is kernel_w*kernel_h exactly as :317-343). */
typedef struct ncnn_cuda_pool2d_desc
{
    int pooling_type; /* 0 max, 1 avg */
    int kernel_w, kernel_h, stride_w, stride_h;
    int pad_left, pad_top;
    int global_pooling;
    int avgpool_count_include_pad;
    int adaptive_pooling; /* output size taken from `top` */
    /* avg without count_include_pad divides by the number of window taps inside [area_x0, area_x1) x [area_y0, area_y1)
     * (input coordinates; taps there but outside the image add 0).  For pad_mode 0/1 that is the image itself
     * (0, w, 0, h); for the SAME modes the reference counts its SAME padding because it tests the explicit
     * pad members only (pooling.cpp:283-300), so the host passes (pad_left_member - pad_left_applied, ...). */
    int area_x0, area_x1, area_y0, area_y1;
} ncnn_cuda_pool2d_desc;

NCNN_CUDA_API int ncnn_cuda_pool2d_forward(const ncnn_cuda_pool2d_desc* desc, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream);

/* ------------------------------------------------------------------ InnerProduct / Gemm
 * A dense layer over the "pixels x channels" view:  top[m][p] = act(bias[p] + sum_k bottom[m][k] * W[p][k]).
 * InnerProduct (src/layer/innerproduct.cpp:84-165): weights [num_output][num_input] with num_input in the
 * reference's flattened c-major order; `in_w,in_h,in_c` describe a 3-D bottom so create() can permute the
 * columns into this backend's channel-innermost order (VGG16 fc6 sees a 7x7x512 blob).
 * Implemented by the same kernels as a 1x1 convolution. */
typedef struct ncnn_cuda_linear_desc
{
    int num_input, num_output;
    int bias_term;
    ncnn_cuda_activation act;
    int elemtype;
    int in_w, in_h, in_c; /* 0,0,0 = bottom rows are already dense K vectors */
} ncnn_cuda_linear_desc;

typedef struct ncnn_cuda_conv2d* ncnn_cuda_linear_t; /* same object as a 1x1 conv */

NCNN_CUDA_API int ncnn_cuda_linear_create(ncnn_cuda_linear_t* fc, const ncnn_cuda_linear_desc* desc, const float* weight_host, const float* bias_host, void* stream);
NCNN_CUDA_API int ncnn_cuda_linear_destroy(ncnn_cuda_linear_t fc);
/* bottom: M x K view (rows = n * P), top: M x num_output */
NCNN_CUDA_API int ncnn_cuda_linear_forward(ncnn_cuda_linear_t fc, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream);

/* General matrix product for Gemm (src/layer/gemm.cpp:250-315) with runtime A and B:
 * C[i][j] = alpha * (sum_k A(i,k) * B(k,j) + beta * Cin(broadcast))  (gemm.cpp:269-303), strided fp32-accumulate SIMT kernel.
 * A(i,k) = a[i*a_rs + k*a_cs] etc.; c_* strides may be 0 for broadcasting (gemm.cpp:274-293). */
typedef struct ncnn_cuda_gemm_args
{
    int M, N, K, batch;
    const void* a; long long a_rs, a_cs, a_bs;
    const void* b; long long b_rs, b_cs, b_bs;
    const void* c; long long c_rs, c_cs, c_bs; /* c may be NULL */
    void* out; long long o_rs, o_cs, o_bs;
    float alpha, beta;
    int elemtype;   /* of a, b, out */
    int c_elemtype; /* of c (bias constants are kept fp32) */
} ncnn_cuda_gemm_args;

NCNN_CUDA_API int ncnn_cuda_gemm_strided(const ncnn_cuda_gemm_args* args, void* stream);

/* ------------------------------------------------------------------ glue operators */
/* unary, in place allowed (bottom == top).  op: see NCNN_CUDA_UNARY_*.
 * ReLU src/layer/relu.cpp:16-50 (p0 = slope), Sigmoid sigmoid.cpp, Swish swish.cpp:14-35,
 * plus the fused_activation codes so a standalone activation layer can reuse them. */
#define NCNN_CUDA_UNARY_RELU      1 /* p0 = negative slope */
#define NCNN_CUDA_UNARY_CLIP      3 /* p0 = min, p1 = max */
#define NCNN_CUDA_UNARY_SIGMOID   4
#define NCNN_CUDA_UNARY_MISH      5
#define NCNN_CUDA_UNARY_HARDSWISH 6 /* p0 = alpha, p1 = beta */
#define NCNN_CUDA_UNARY_SWISH     7
#define NCNN_CUDA_UNARY_SCALE     8 /* x * p0 (Dropout with scale != 1, dropout.cpp:14-40) */
#define NCNN_CUDA_UNARY_TANH      9
#define NCNN_CUDA_UNARY_HARDSIGMOID 10 /* p0 = alpha, p1 = beta */
#define NCNN_CUDA_UNARY_GELU      11 /* p0 != 0: the tanh form (fast_gelu), else 0.5 x erfc(-x / sqrt 2); src/layer/gelu.cpp:21-58 */
NCNN_CUDA_API int ncnn_cuda_unary(int op, float p0, float p1, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream);

/* BatchNorm (src/layer/batchnorm.cpp:57-120: value = b * value + a) and Scale (src/layer/scale.cpp:44-168: value * s + bias):
 * top = bottom * scale[i] + shift[i], i = channel for 1-D/3-D/4-D blobs, i = row (h) for 2-D blobs.  scale/shift are DEVICE
 * fp32 arrays (shift may be NULL = 0); bottom == top (in place) is allowed. */
NCNN_CUDA_API int ncnn_cuda_channel_affine(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const float* scale_dev, const float* shift_dev, void* stream);

/* LRN (src/layer/lrn.cpp:26-170): top = bottom * (bias + alpha/size * sum of squares over the window)^-beta, window =
 * local_size channels (region_type 0) or local_size x local_size pixels (region_type 1) around each element, zeros outside.
 * bottom and top must be distinct blobs (the window reads neighbours). */
NCNN_CUDA_API int ncnn_cuda_lrn(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int region_type, int local_size, float alpha, float beta, float bias, void* stream);

/* Reduction (src/layer/reduction.cpp:216-752): operation 0 sum, 1 asum, 2 sumsq, 3 mean, 4 max, 5 min, 6 prod, 7 L1, 8 L2,
 * 9 logsum, 10 logsumexp over the flagged axes of `bottom` (flags for axes the rank does not have are ignored; a 1-D blob
 * always reduces w, :786-789).  `top` has the shape resolve_reduce_flags_and_output_shape (:753-856) gives: the same rank
 * with 1s when keepdims, else the surviving extents in (w, h, d, c) order (a full reduction is a 1-D blob of w = 1).
 * `coeff` multiplies the result (divided by the reduced element count for the mean, :709-749). */
NCNN_CUDA_API int ncnn_cuda_reduction(int operation, int reduce_w, int reduce_h, int reduce_d, int reduce_c, int keepdims, float coeff,
                                      const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream);

/* LayerNorm (src/layer/layernorm.cpp:38-181): every run of `group_size` consecutive elements along the reference's planar
 * order inside one channel (1-D / 2-D blobs: inside one row) is normalised to zero mean / unit variance (biased, + eps) and,
 * when gamma/beta (fp32 device arrays of group_size values, both or neither) are given, scaled and shifted element-wise.
 * group_size is w, w*h or w*h*d as the layer's affine_size selects; in place (bottom == top) allowed. */
NCNN_CUDA_API int ncnn_cuda_layernorm(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int group_size, float eps, const float* gamma_dev,
                                      const float* beta_dev, void* stream);

/* YOLOv8 head decode on the device: the generate_proposals step of the reference's examples/yolov8.cpp:160-273.
 * `pred`: 2-D blob per image, one row per anchor point = 4 x 16 box-distribution logits + num_class class logits
 * (w = 64 + num_class, h = sum over strides of (in_w / stride) * (in_h / stride), rows ordered stride by stride, y-major).
 * `proposals`: 2-D fp32 blob, w = 6, h = pred->h, same batch: per anchor {x, y, width, height, prob, label} in input pixels,
 * prob = sigmoid(max class logit); anchors with prob < prob_threshold get prob = 0, label = -1.  One record per anchor in
 * anchor order (deterministic); sorting and NMS (examples/yolov8.cpp:73-153) stay with the caller. */
NCNN_CUDA_API int ncnn_cuda_yolov8_decode(const ncnn_cuda_tensor* pred, const int* strides, int num_strides, int in_w, int in_h, float prob_threshold,
                                          const ncnn_cuda_tensor* proposals, void* stream);

/* ShuffleChannel (src/layer/shufflechannel.cpp:22-60): top channel group*j + i = bottom channel (c/group)*i + j.
 * `group` is the effective group count (the caller resolves the layer's `reverse` flag: group = c / group). */
NCNN_CUDA_API int ncnn_cuda_shuffle_channel(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int group, void* stream);

/* Eltwise src/layer/eltwise.cpp:22-178: op 0 PROD, 1 SUM (optional coeffs), 2 MAX over `count` same-shape
 * bottoms; `relu` fuses a following ReLU layer (graph-level fold). */
NCNN_CUDA_API int ncnn_cuda_eltwise(int op, const ncnn_cuda_tensor* bottoms, int count, const float* coeffs, int relu, const ncnn_cuda_tensor* top, void* stream);

/* BinaryOp src/layer/binaryop.cpp (op codes binaryop.h:24-45) with numpy-style broadcasting over the
 * logical (w,h,d,c) dims (docs/developer-guide/binaryop-broadcasting.md); b may be NULL for the
 * with_scalar form (then `scalar` is used). */
NCNN_CUDA_API int ncnn_cuda_binaryop(int op, const ncnn_cuda_tensor* a, const ncnn_cuda_tensor* b, float scalar, const ncnn_cuda_tensor* top, void* stream);

/* Concat src/layer/concat.cpp:14-292 / Slice src/layer/slice.cpp: `axis` is the reference's positive axis
 * index for the blob rank (3-D: 0 = c, 1 = h, 2 = w).  One call copies one bottom into `top` at element
 * offset `offset` along that axis (concat), or the range [offset, offset + extent(top)) of `bottom` (slice). */
NCNN_CUDA_API int ncnn_cuda_copy_into_axis(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, int offset, void* stream);
NCNN_CUDA_API int ncnn_cuda_copy_from_axis(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, int offset, void* stream);

/* Interp src/layer/interp.cpp: resize_type 1 nearest (:606-625; hs/ws are the reference's float source steps,
 * in_y = min((int)(y*hs), h-1)), 2 bilinear (linear_coeffs :56-90, align_corner 0/1; hs/ws unused) */
NCNN_CUDA_API int ncnn_cuda_interp(int resize_type, int align_corner, float hs, float ws, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream);

/* Softmax src/layer/softmax.cpp:32-250 along the reference's positive axis for the blob rank; in place allowed */
NCNN_CUDA_API int ncnn_cuda_softmax(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, void* stream);

/* Padding src/layer/padding.cpp: constant (type 0), replicate (1), reflect (2) on w/h (and c via front/behind) */
NCNN_CUDA_API int ncnn_cuda_padding(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int top_pad, int left_pad, int front_pad, int type, float value, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* NCNN_CUDA_H */
