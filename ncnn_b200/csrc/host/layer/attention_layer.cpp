// attention_layer.cpp -- MultiHeadAttention (src/layer/multiheadattention.cpp:240-560 of the reference, the fp32 forward;
// refers to torch.nn.MultiheadAttention) as a composition of the backend's kernels: the four affine maps run on the dense path
// (ncnn_cuda_linear_*, tensor cores for 16-bit blobs), Q.K^T and P.V per head on the strided batched product
// (ncnn_cuda_gemm_strided, one launch per sample with the heads as its batch), the softmax over the key axis on
// ncnn_cuda_softmax.  Supported: self- and cross-attention (1, 2 or 3 bottoms), no attention mask, no kv cache, no
// quantised weights -- anything else fails at load_param, loudly.
#include "cuda_layers.h"

#include <math.h>
#include <vector>

namespace ncnn {

MultiHeadAttention::MultiHeadAttention()
{
    one_blob_only = false;
    support_inplace = false;
    q_fc = k_fc = v_fc = o_fc = 0;
}

MultiHeadAttention::~MultiHeadAttention()
{
    destroy_pipeline(Option());
}

// src/layer/multiheadattention.cpp:64-118
int MultiHeadAttention::load_param(const ParamDict& pd)
{
    embed_dim = pd.get(0, 0);
    num_heads = pd.get(1, 1);
    weight_data_size = pd.get(2, 0);
    kdim = pd.get(3, embed_dim);
    vdim = pd.get(4, embed_dim);
    attn_mask = pd.get(5, 0);
    if (embed_dim <= 0 || num_heads <= 0 || embed_dim % num_heads != 0 || weight_data_size <= 0 || weight_data_size % embed_dim != 0) return -1;
    scale = pd.get(6, 1.f / sqrtf((float)(embed_dim / num_heads)));
    kv_cache = pd.get(7, 0);
    int quantize_term = pd.get(18, 0);
    if (attn_mask || kv_cache || quantize_term)
    {
        NCNN_LOGE("MultiHeadAttention: attn_mask / kv_cache / quantised weights are not supported by the CUDA backend");
        return -1;
    }
    return 0;
}

// src/layer/multiheadattention.cpp:191-238 (the unquantised branch)
int MultiHeadAttention::load_model(const ModelBin& mb)
{
    const int qdim = weight_data_size / embed_dim;
    q_weight_data = mb.load(embed_dim * qdim, 0);
    q_bias_data = mb.load(embed_dim, 1);
    k_weight_data = mb.load(embed_dim * kdim, 0);
    k_bias_data = mb.load(embed_dim, 1);
    v_weight_data = mb.load(embed_dim * vdim, 0);
    v_bias_data = mb.load(embed_dim, 1);
    out_weight_data = mb.load(qdim * embed_dim, 0);
    out_bias_data = mb.load(qdim, 1);
    if (q_weight_data.empty() || q_bias_data.empty() || k_weight_data.empty() || k_bias_data.empty() || v_weight_data.empty() || v_bias_data.empty()
        || out_weight_data.empty() || out_bias_data.empty())
        return -100;
    return 0;
}

static int make_fc(ncnn_cuda_linear_t* fc, int num_input, int num_output, const float* w, const float* b, int elemtype)
{
    ncnn_cuda_linear_desc d;
    memset(&d, 0, sizeof(d));
    d.num_input = num_input;
    d.num_output = num_output;
    d.bias_term = 1;
    d.elemtype = elemtype;
    return ncnn_cuda_linear_create(fc, &d, w, b, 0);
}

int MultiHeadAttention::create_pipeline(const Option& opt)
{
    const int qdim = weight_data_size / embed_dim;
    const int elemtype = opt.cuda_elemtype();
    // the reference scales the projected query, (x.Wq + bq) * scale (:299-302): folded into the weights and the bias here
    std::vector<float> wq((size_t)embed_dim * qdim), bq((size_t)embed_dim);
    for (size_t i = 0; i < wq.size(); i++) wq[i] = ((const float*)q_weight_data.data)[i] * scale;
    for (size_t i = 0; i < bq.size(); i++) bq[i] = ((const float*)q_bias_data.data)[i] * scale;
    int ret = make_fc(&q_fc, qdim, embed_dim, wq.data(), bq.data(), elemtype);
    if (ret == 0) ret = make_fc(&k_fc, kdim, embed_dim, (const float*)k_weight_data.data, (const float*)k_bias_data.data, elemtype);
    if (ret == 0) ret = make_fc(&v_fc, vdim, embed_dim, (const float*)v_weight_data.data, (const float*)v_bias_data.data, elemtype);
    if (ret == 0) ret = make_fc(&o_fc, embed_dim, qdim, (const float*)out_weight_data.data, (const float*)out_bias_data.data, elemtype);
    if (ret != 0) return ret;
    if (opt.lightmode)
    {
        q_weight_data.release();
        k_weight_data.release();
        v_weight_data.release();
        out_weight_data.release();
    }
    return 0;
}

int MultiHeadAttention::destroy_pipeline(const Option&)
{
    if (q_fc) ncnn_cuda_linear_destroy(q_fc);
    if (k_fc) ncnn_cuda_linear_destroy(k_fc);
    if (v_fc) ncnn_cuda_linear_destroy(v_fc);
    if (o_fc) ncnn_cuda_linear_destroy(o_fc);
    q_fc = k_fc = v_fc = o_fc = 0;
    return 0;
}

int MultiHeadAttention::forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
{
    // src/layer/multiheadattention.cpp:914-1010 without mask / cache: 1 bottom = self-attention, 2 = (q, k=v), 3 = (q, k, v)
    const size_t nb = bottom_blobs.size();
    if (nb < 1 || nb > 3 || !q_fc) return -1;
    const CudaMat& q_blob = bottom_blobs[0];
    const CudaMat& k_blob = bottom_blobs[nb == 1 ? 0 : 1];
    const CudaMat& v_blob = bottom_blobs[nb == 1 ? 0 : (nb == 2 ? 1 : 2)];
    const int qdim = weight_data_size / embed_dim;
    if (q_blob.dims != 2 || k_blob.dims != 2 || v_blob.dims != 2 || q_blob.w != qdim || k_blob.w != kdim || v_blob.w != vdim || k_blob.h != v_blob.h) return -1;
    if (k_blob.n != q_blob.n || v_blob.n != q_blob.n) return -1;
    const int src = q_blob.h, dst = k_blob.h, heads = num_heads, D = embed_dim / num_heads, n = q_blob.n < 1 ? 1 : q_blob.n;
    const int et = q_blob.elemtype;
    CudaAllocator* wa = cmd.workspace_allocator(opt);
    void* st = cmd.stream();

    CudaMat Q, K, V, S, O;
    Q.create(embed_dim, src, et, q_blob.n, wa);
    K.create(embed_dim, dst, et, q_blob.n, wa);
    V.create(embed_dim, dst, et, q_blob.n, wa);
    S.create(dst, heads * src, et, q_blob.n, wa); // scores / probabilities: row (h * src + i), column j
    O.create(embed_dim, src, et, q_blob.n, wa);
    if (Q.empty() || K.empty() || V.empty() || S.empty() || O.empty()) return -100;
    ncnn_cuda_tensor qb = q_blob.view(), kb = k_blob.view(), vb = v_blob.view();
    ncnn_cuda_tensor tq = Q.view(), tk = K.view(), tv = V.view(), ts = S.view(), to = O.view();
    int ret = ncnn_cuda_linear_forward(q_fc, &qb, &tq, st); // :280-305
    if (ret == 0) ret = ncnn_cuda_linear_forward(k_fc, &kb, &tk, st); // :307-352
    if (ret == 0) ret = ncnn_cuda_linear_forward(v_fc, &vb, &tv, st); // :354-400
    if (ret != 0) return ret;
    const size_t es = et == NCNN_CUDA_F32 ? 4 : 2;
    for (int b = 0; b < n; b++)
    {
        // xqk[h][i][j] = sum_d Q[i][h*D + d] * K[j][h*D + d]   (:402-430)
        ncnn_cuda_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.M = src;
        g.N = dst;
        g.K = D;
        g.batch = heads;
        g.a = (const char*)tq.data + (size_t)b * tq.nstep * es;
        g.a_rs = tq.cpitch;
        g.a_cs = 1;
        g.a_bs = D;
        g.b = (const char*)tk.data + (size_t)b * tk.nstep * es;
        g.b_rs = 1;
        g.b_cs = tk.cpitch;
        g.b_bs = D;
        g.out = (char*)ts.data + (size_t)b * ts.nstep * es;
        g.o_rs = ts.cpitch;
        g.o_cs = 1;
        g.o_bs = (long long)src * ts.cpitch;
        g.alpha = 1.f;
        g.beta = 0.f;
        g.elemtype = et;
        g.c_elemtype = NCNN_CUDA_F32;
        ret = ncnn_cuda_gemm_strided(&g, st);
        if (ret != 0) return ret;
    }
    ret = ncnn_cuda_softmax(&ts, &ts, 1, st); // over the key axis j (:460-464: Softmax axis -1 of the (dst, src, heads) cube)
    if (ret != 0) return ret;
    for (int b = 0; b < n; b++)
    {
        // xqkv[i][h*D + d] = sum_j P[h][i][j] * V[j][h*D + d]   (:466-497)
        ncnn_cuda_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.M = src;
        g.N = D;
        g.K = dst;
        g.batch = heads;
        g.a = (const char*)ts.data + (size_t)b * ts.nstep * es;
        g.a_rs = ts.cpitch;
        g.a_cs = 1;
        g.a_bs = (long long)src * ts.cpitch;
        g.b = (const char*)tv.data + (size_t)b * tv.nstep * es;
        g.b_rs = tv.cpitch;
        g.b_cs = 1;
        g.b_bs = D;
        g.out = (char*)to.data + (size_t)b * to.nstep * es;
        g.o_rs = to.cpitch;
        g.o_cs = 1;
        g.o_bs = D;
        g.alpha = 1.f;
        g.beta = 0.f;
        g.elemtype = et;
        g.c_elemtype = NCNN_CUDA_F32;
        ret = ncnn_cuda_gemm_strided(&g, st);
        if (ret != 0) return ret;
    }
    CudaMat& top = top_blobs[0];
    top.create(qdim, src, et, q_blob.n, cmd.blob_allocator(opt));
    if (top.empty()) return -100;
    ncnn_cuda_tensor tt = top.view();
    ret = ncnn_cuda_linear_forward(o_fc, &to, &tt, st); // :499-525
    // the workspace blobs are consumed by kernels already enqueued on this stream; the allocators recycle in stream order
    return ret;
}

} // namespace ncnn
