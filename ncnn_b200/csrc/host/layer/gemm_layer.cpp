// gemm_layer.cpp -- Gemm (src/layer/gemm.cpp:67-248 params/weights, :579-740 forward).
// Y = alpha * (op(A) * op(B) + beta * C), operands 2-D (w, h) or the 3-D (w, 1, c) form the reference accepts (:608-650).
//
// tcgen05 routes (the constant operand becomes the re-packed weight matrix of the implicit-GEMM kernel, tc_gemm.cuh; alpha is
// folded into the weights, alpha * beta * C into the epilogue bias when C is constant and broadcasts along the weight rows):
//   constant B, any transB   : Y[m][n]   = sum_k A[m][k] W[n][k],  W = alpha * op(B)^T      (the nn.Linear form and its transposes)
//   constant A, runtime B^T  : Y^T[n][m] = sum_k B^T[n][k] W[m][k], W = alpha * op(A)       (transB = 1: B arrives as [N][K])
// followed by one tiled transpose when the requested output orientation is the other one.  Everything else -- two runtime
// operands, transposed runtime activations, C broadcasts across the weight rows, 3-D operands -- goes through the strided
// CUDA-core kernel with transposes and broadcasts expressed as strides.
#include "cuda_layers.h"

#include <vector>

namespace ncnn {

Gemm::Gemm()
{
    one_blob_only = false;
    support_inplace = false;
    linear = 0;
    linear_mode = 0;
    elemtype = NCNN_CUDA_F32;
}

Gemm::~Gemm()
{
    if (linear) ncnn_cuda_linear_destroy(linear);
}

int Gemm::load_param(const ParamDict& pd)
{
    alpha = pd.get(0, 1.f);
    beta = pd.get(1, 1.f);
    transA = pd.get(2, 0);
    transB = pd.get(3, 0);
    constantA = pd.get(4, 0);
    constantB = pd.get(5, 0);
    constantC = pd.get(6, 0);
    constantM = pd.get(7, 0);
    constantN = pd.get(8, 0);
    constantK = pd.get(9, 0);
    constant_broadcast_type_C = pd.get(10, 0);
    output_N1M = pd.get(11, 0);
    output_elempack = pd.get(12, 0);
    output_elemtype = pd.get(13, 0);
    output_transpose = pd.get(14, 0);
    if (pd.get(18, 0) != 0)
    {
        NCNN_LOGE("Gemm: quantized forms (quantize_term) are outside the CUDA backend's scope");
        return -1;
    }
    // gemm.cpp:153-160
    if (constantA == 0 && constantB == 1 && constantC == 1) one_blob_only = true;
    if (constantA == 1 && constantB == 0 && constantC == 1) one_blob_only = true;
    if (constantA == 1 && constantB == 1 && constantC == 0) one_blob_only = true;
    return 0;
}

// gemm.cpp:164-215
int Gemm::load_model(const ModelBin& mb)
{
    if (constantA == 1)
    {
        A_data = transA == 0 ? mb.load(constantK, constantM, 0) : mb.load(constantM, constantK, 0);
        if (A_data.empty()) return -100;
    }
    if (constantB == 1)
    {
        B_data = transB == 0 ? mb.load(constantN, constantK, 0) : mb.load(constantK, constantN, 0);
        if (B_data.empty()) return -100;
    }
    if (constantC == 1 && constant_broadcast_type_C != -1)
    {
        if (constant_broadcast_type_C == 0) C_data = mb.load(1, 0);
        if (constant_broadcast_type_C == 1) C_data = mb.load(constantM, 0);
        if (constant_broadcast_type_C == 2) C_data = mb.load(1, constantM, 0);
        if (constant_broadcast_type_C == 3) C_data = mb.load(constantN, constantM, 0);
        if (constant_broadcast_type_C == 4) C_data = mb.load(constantN, 1, 0);
        if (C_data.empty()) return -100;
    }
    return 0;
}

int upload_const(const Mat& src, int elemtype, CudaMat& dst)
{
    // constants are 1-D/2-D host Mats; keep them as device blobs of the same logical shape
    CudaContext* ctx = acquire_cuda_context(-1);
    if (!ctx) return -1;
    int ret;
    {
        CudaCompute cmd(ctx);
        Option o;
        o.use_fp16_storage = elemtype == NCNN_CUDA_F16;
        o.use_bf16_storage = elemtype == NCNN_CUDA_BF16;
        o.blob_cuda_allocator = get_cuda_weight_allocator(ctx->device_index);
        ret = cmd.record_upload(src, dst, o);
        int s = cmd.submit_and_wait();
        if (ret == 0) ret = s;
    }
    reclaim_cuda_context(ctx);
    return ret;
}

// logical matrix view of a constant host Mat: 2-D (w, h) rows = h, or 3-D (w, 1, c) rows = c (src/layer/gemm.cpp:617-650)
static void host_matrix(const Mat& m, int& rows, int& cols, size_t& rstep)
{
    cols = m.w;
    if (m.dims == 3)
    {
        rows = m.c;
        rstep = m.cstep;
    }
    else
    {
        rows = m.h;
        rstep = (size_t)m.w;
    }
}

int Gemm::create_pipeline(const Option& opt)
{
    elemtype = opt.cuda_elemtype();
    const bool has_c = constantC == 1 && constant_broadcast_type_C != -1;
    const int bt = constant_broadcast_type_C;
    linear_mode = 0;
    // ---- constant-operand routes (tcgen05 for 16-bit blobs, the strict-fp32 implicit-GEMM kernel for fp32 blobs): 2-D results,
    // constant C (or none) that broadcasts along the weight rows
    if (output_N1M == 0 && constantC == 1)
    {
        const bool c_cols_ok = !has_c || bt == 0 || bt == 4;              // scalar / per column n
        const bool c_rows_ok = !has_c || bt == 0 || bt == 1 || bt == 2;   // scalar / per row m
        if (constantB == 1 && constantA == 0 && transA == 0 && c_cols_ok && B_data.dims == 2)
            linear_mode = 1;
        else if (constantA == 1 && constantB == 0 && transB == 1 && c_rows_ok && A_data.dims == 2)
            linear_mode = 2;
    }
    if (linear_mode)
    {
        // W[out][k]: mode 1 out = n, W = alpha * op(B)^T; mode 2 out = m, W = alpha * op(A)
        const Mat& Wsrc = linear_mode == 1 ? B_data : A_data;
        const int out = linear_mode == 1 ? constantN : constantM;
        // stored [out][K] when (mode 1, transB = 1) or (mode 2, transA = 0); otherwise stored [K][out]
        const bool stored_out_major = linear_mode == 1 ? transB == 1 : transA == 0;
        std::vector<float> W((size_t)out * constantK), bias;
        const float* src = (const float*)Wsrc.data;
        for (int o = 0; o < out; o++)
            for (int k = 0; k < constantK; k++)
                W[(size_t)o * constantK + k] = alpha * (stored_out_major ? src[(size_t)o * Wsrc.w + k] : src[(size_t)k * Wsrc.w + o]);
        if (has_c)
        {
            bias.resize(out);
            const float* c = (const float*)C_data.data;
            for (int o = 0; o < out; o++) bias[o] = alpha * beta * (bt == 0 ? c[0] : c[o]);
        }
        ncnn_cuda_linear_desc d;
        memset(&d, 0, sizeof(d));
        d.num_input = constantK;
        d.num_output = out;
        d.bias_term = has_c ? 1 : 0;
        d.elemtype = elemtype;
        int ret = ncnn_cuda_linear_create(&linear, &d, W.data(), has_c ? bias.data() : 0, 0);
        if (ret != 0) return ret;
        return 0;
    }
    if (constantA)
    {
        int ret = upload_const(A_data, elemtype, A_dev);
        if (ret != 0) return ret;
    }
    if (constantB)
    {
        int ret = upload_const(B_data, elemtype, B_dev);
        if (ret != 0) return ret;
    }
    if (has_c)
    {
        int ret = upload_const(C_data, NCNN_CUDA_F32, C_dev);
        if (ret != 0) return ret;
    }
    return 0;
}

int Gemm::destroy_pipeline(const Option&)
{
    if (linear) ncnn_cuda_linear_destroy(linear);
    linear = 0;
    A_dev.release();
    B_dev.release();
    C_dev.release();
    return 0;
}

int Gemm::forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
{
    std::vector<CudaMat> b(1, bottom_blob), t(1);
    int ret = forward(b, t, cmd, opt);
    top_blob = t[0];
    return ret;
}

// rows x cols view of a device operand: a 2-D blob (w, h) is h pixels of w channels, a 3-D blob (w, 1, c) is w pixels of c
// channels -- the same logical matrix (rows = c) with the strides swapped
struct DevMatrix
{
    int rows, cols;
    long long rs, cs;
    bool ok;
};

static DevMatrix dev_matrix(const CudaMat& m)
{
    DevMatrix v;
    v.ok = true;
    if (m.dims == 2)
    {
        v.rows = m.h;
        v.cols = m.w;
        v.rs = m.cpitch;
        v.cs = 1;
    }
    else if (m.dims == 3 && m.h == 1)
    {
        v.rows = m.c;
        v.cols = m.w;
        v.rs = 1;
        v.cs = m.cpitch;
    }
    else if (m.dims == 1)
    {
        v.rows = 1;
        v.cols = m.w;
        v.rs = m.cpitch;
        v.cs = 1;
    }
    else
        v.ok = false;
    return v;
}

int Gemm::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    CudaMat& top = top_blobs[0];
    if (linear && bottom_blobs[0].dims == 2 && bottom_blobs[0].w == constantK && bottom_blobs[0].elemtype == elemtype)
    {
        // rows of the runtime operand x W^T: mode 1 gives Y (M x N), mode 2 gives Y^T (N x M)
        const CudaMat& X = bottom_blobs[0];
        const int out = linear_mode == 1 ? constantN : constantM;
        const bool want_transposed = linear_mode == 1 ? output_transpose != 0 : output_transpose == 0;
        CudaMat direct;
        CudaMat& first = want_transposed ? direct : top;
        first.create(out, X.h, X.elemtype, X.n, want_transposed ? cmd.workspace_allocator(opt) : cmd.blob_allocator(opt));
        if (first.empty()) return -100;
        ncnn_cuda_tensor b = X.view(), t = first.view();
        int ret = ncnn_cuda_linear_forward(linear, &b, &t, cmd.stream());
        if (ret != 0 || !want_transposed) return ret;
        top.create(X.h, out, X.elemtype, X.n, cmd.blob_allocator(opt));
        if (top.empty()) return -100;
        ncnn_cuda_tensor s = direct.view(), d = top.view();
        return ncnn_cuda_permute(&s, &d, 1, cmd.stream());
    }
    if (linear)
    {
        NCNN_LOGE("Gemm %s: runtime operand does not match the constant-operand pipeline (dims %d w %d, K %d)", name.c_str(), bottom_blobs[0].dims, bottom_blobs[0].w, constantK);
        return -1;
    }

    const CudaMat& A0 = constantA ? A_dev : bottom_blobs[0];
    const CudaMat& B0 = constantB ? B_dev : (constantA ? bottom_blobs[0] : bottom_blobs[1]);
    const DevMatrix Am = dev_matrix(A0), Bm = dev_matrix(B0);
    if (!Am.ok || !Bm.ok)
    {
        NCNN_LOGE("Gemm %s: operands must be 2-D or (w, 1, c) blobs", name.c_str());
        return -1;
    }
    const int M = transA ? Am.cols : Am.rows;
    const int K = transA ? Am.rows : Am.cols;
    const int N = transB ? Bm.rows : Bm.cols;
    if ((transB ? Bm.cols : Bm.rows) != K) return -1;

    CudaMat C;
    int bt = 0;
    if (constantC)
    {
        if (constant_broadcast_type_C != -1)
        {
            C = C_dev;
            bt = constant_broadcast_type_C;
        }
    }
    else
    {
        // gemm.cpp:668-715
        if (constantA && constantB && bottom_blobs.size() == 1)
            C = bottom_blobs[0];
        else if ((constantA || constantB) && bottom_blobs.size() == 2)
            C = bottom_blobs[1];
        else if (bottom_blobs.size() == 3)
            C = bottom_blobs[2];
        if (!C.empty())
        {
            if (C.dims == 1 && C.w == 1) bt = 0;
            if (C.dims == 1 && C.w == M) bt = 1;
            if (C.dims == 1 && C.w == N) bt = 4;
            if (C.dims == 2 && C.w == 1 && C.h == M) bt = 2;
            if (C.dims == 2 && C.w == N && C.h == M) bt = 3;
            if (C.dims == 2 && C.w == N && C.h == 1) bt = 4;
        }
    }

    int n = 1;
    for (size_t i = 0; i < bottom_blobs.size(); i++)
        if (bottom_blobs[i].n > n) n = bottom_blobs[i].n;
    const int et = A0.elemtype;
    if (output_transpose)
    {
        if (output_N1M)
            top.create(M, 1, N, et, n, cmd.blob_allocator(opt));
        else
            top.create(M, N, et, n, cmd.blob_allocator(opt));
    }
    else
    {
        if (output_N1M)
            top.create(N, 1, M, et, n, cmd.blob_allocator(opt));
        else
            top.create(N, M, et, n, cmd.blob_allocator(opt));
    }
    if (top.empty()) return -100;

    ncnn_cuda_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = M;
    g.N = N;
    g.K = K;
    g.batch = n;
    g.a = A0.data;
    g.a_rs = transA ? Am.cs : Am.rs; // A(i, k)
    g.a_cs = transA ? Am.rs : Am.cs;
    g.a_bs = (constantA || A0.n <= 1) ? 0 : (long long)A0.nstep;
    g.b = B0.data;
    g.b_rs = transB ? Bm.cs : Bm.rs; // B(k, j)
    g.b_cs = transB ? Bm.rs : Bm.cs;
    g.b_bs = (constantB || B0.n <= 1) ? 0 : (long long)B0.nstep;
    if (!C.empty())
    {
        g.c = C.data;
        g.c_elemtype = C.elemtype;
        g.c_bs = (constantC || C.n <= 1) ? 0 : (long long)C.nstep;
        // device layout of 1-D blobs: one pixel, w channels; 2-D blobs: h pixels of w channels
        if (bt == 0)
            g.c_rs = 0, g.c_cs = 0;
        else if (bt == 1)
            g.c_rs = 1, g.c_cs = 0;
        else if (bt == 2)
            g.c_rs = C.cpitch, g.c_cs = 0;
        else if (bt == 3)
            g.c_rs = C.cpitch, g.c_cs = 1;
        else
            g.c_rs = 0, g.c_cs = 1;
    }
    g.out = top.data;
    // element (i, j): row i of M, column j of N
    if (output_N1M)
    {
        // 3-D top: the channel axis is innermost on the device
        if (output_transpose)
            g.o_rs = top.cpitch, g.o_cs = 1; // (w = M, h = 1, c = N): pixel = i, channel = j
        else
            g.o_rs = 1, g.o_cs = top.cpitch; // (w = N, h = 1, c = M): pixel = j, channel = i
    }
    else if (output_transpose)
        g.o_rs = 1, g.o_cs = top.cpitch;
    else
        g.o_rs = top.cpitch, g.o_cs = 1;
    g.o_bs = (long long)top.nstep;
    g.alpha = alpha;
    g.beta = beta;
    g.elemtype = et;
    return ncnn_cuda_gemm_strided(&g, cmd.stream());
}

} // namespace ncnn
