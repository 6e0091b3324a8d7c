// gemm_strided.cu -- general strided matrix product for the Gemm layer's runtime-operand forms
// (src/layer/gemm.cpp:250-315 of the reference: any transA/transB via strides, alpha/beta, 5 C-broadcast kinds
// via zero strides).  CUDA-core, fp32 accumulate, 32x32 shared-memory tiles.  Gemm with a constant B (the form
// pnnx emits for nn.Linear) does not come here: it runs on the tensor-core path through ncnn_cuda_linear_*.
#include "common.cuh"

using namespace ncnn_cuda;

namespace {

template<typename T, typename TC>
__global__ void __launch_bounds__(256) gemm_strided_kernel(ncnn_cuda_gemm_args g)
{
    __shared__ float As[32][33];
    __shared__ float Bs[32][33];
    const int batch = blockIdx.z;
    const T* a = (const T*)g.a + (long long)batch * g.a_bs;
    const T* b = (const T*)g.b + (long long)batch * g.b_bs;
    T* out = (T*)g.out + (long long)batch * g.o_bs;
    const TC* c = g.c ? (const TC*)g.c + (long long)batch * g.c_bs : 0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < g.K; k0 += 32)
    {
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            int ii = ty + 8 * r;
            int i = i0 + ii, k = k0 + tx;
            As[ii][tx] = (i < g.M && k < g.K) ? to_f32(a[(long long)i * g.a_rs + (long long)k * g.a_cs]) : 0.f;
            int kk = ty + 8 * r;
            int kb = k0 + kk, j = j0 + tx;
            Bs[kk][tx] = (kb < g.K && j < g.N) ? to_f32(b[(long long)kb * g.b_rs + (long long)j * g.b_cs]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; k++)
        {
            float bv = Bs[k][tx];
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] = fmaf(As[ty + 8 * r][k], bv, acc[r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int i = i0 + ty + 8 * r, j = j0 + tx;
        if (i < g.M && j < g.N)
        {
            // the reference's order (gemm.cpp:269-303): sum = beta * C; sum += A.B; sum *= alpha
            float v = acc[r];
            if (c) v += g.beta * to_f32(c[(long long)i * g.c_rs + (long long)j * g.c_cs]);
            v *= g.alpha;
            out[(long long)i * g.o_rs + (long long)j * g.o_cs] = from_f32<T>(v);
        }
    }
}

template<typename T>
static int run(const ncnn_cuda_gemm_args* g, cudaStream_t stream)
{
    if (g->M == 0 || g->N == 0) return 0;
    dim3 grid(ceil_div(g->N, 32), ceil_div(g->M, 32), g->batch < 1 ? 1 : g->batch);
    if (g->c_elemtype == NCNN_CUDA_F32)
        gemm_strided_kernel<T, float><<<grid, 256, 0, stream>>>(*g);
    else
        gemm_strided_kernel<T, T><<<grid, 256, 0, stream>>>(*g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" int ncnn_cuda_gemm_strided(const ncnn_cuda_gemm_args* g, void* stream)
{
    NC_REQUIRE(g && g->a && g->b && g->out, "gemm_strided: null operand");
    NC_REQUIRE(!g->c || g->c_elemtype == NCNN_CUDA_F32 || g->c_elemtype == g->elemtype, "gemm_strided: unsupported C element type");
    switch (g->elemtype)
    {
    case NCNN_CUDA_F32: return run<float>(g, as_stream(stream));
    case NCNN_CUDA_BF16: return run<__nv_bfloat16>(g, as_stream(stream));
    case NCNN_CUDA_F16: return run<__half>(g, as_stream(stream));
    }
    return -1;
}
