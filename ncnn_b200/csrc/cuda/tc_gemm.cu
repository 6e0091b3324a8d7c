// tc_gemm.cu -- host side of the tcgen05 implicit-GEMM path: weight re-packing (the create_pipeline step of
// Convolution / InnerProduct, cf. src/layer/x86/convolution_x86.cpp:279-500 in the reference), TMA descriptor
// encoding (tiled for weights and 1x1 convs, im2col mode for everything else) and the persistent-kernel launch.
#include "tc_gemm.cuh"

#include <mutex>
#include <string.h>
#include <vector>

namespace ncnn_cuda {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

static PFN_encodeTiled g_encodeTiled = 0;
static PFN_encodeIm2col g_encodeIm2col = 0;
static int g_tc_state = -1; // -1 unknown, 0 unavailable, 1 ok
static int g_driver_version = 0;
static std::mutex g_tc_mutex;

int tc_available()
{
    std::lock_guard<std::mutex> lk(g_tc_mutex);
    if (g_tc_state >= 0) return g_tc_state;
    g_tc_state = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return 0;
    cudaDriverEntryPointQueryResult qres;
    void* fn = 0;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return 0;
    g_encodeTiled = (PFN_encodeTiled)fn;
    fn = 0;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return 0;
    g_encodeIm2col = (PFN_encodeIm2col)fn;
    cudaDriverGetVersion(&g_driver_version);
    g_tc_state = 1;
    return 1;
}

int tc_pick_block_k(int inch)
{
    if (inch > 32) return 64;
    if (inch > 16) return 32;
    return 16;
}

int tc_pick_block_n(int outch)
{
    if (outch > 128) return 256;
    if (outch > 64) return 128;
    if (outch > 32) return 64;
    return 32;
}

static inline uint16_t f32_to_16(float f, int elemtype)
{
    if (elemtype == NCNN_CUDA_BF16)
    {
        __nv_bfloat16 b = __float2bfloat16_rn(f);
        uint16_t u;
        memcpy(&u, &b, 2);
        return u;
    }
    __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}

static CUtensorMapSwizzle swizzle_for(int block_k)
{
    return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static CUtensorMapDataType dtype_for(int elemtype)
{
    return elemtype == NCNN_CUDA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

int tc_plan_create(TcPlan* plan, int elemtype, int inch, int outch, int taps, const float* w, const float* bias, cudaStream_t stream)
{
    memset(plan, 0, sizeof(*plan));
    if (!tc_available()) return -1;
    plan->elemtype = elemtype;
    plan->block_k = tc_pick_block_k(inch);
    plan->block_n = tc_pick_block_n(outch);
    plan->cblocks = (inch + plan->block_k - 1) / plan->block_k;
    plan->taps = taps;
    plan->num_k_blocks = taps * plan->cblocks;
    plan->Kp = plan->num_k_blocks * plan->block_k;
    plan->outch = outch;
    plan->outch_pad = ((outch + plan->block_n - 1) / plan->block_n) * plan->block_n;

    const int cpad = plan->cblocks * plan->block_k;
    std::vector<uint16_t> packed((size_t)outch * plan->Kp, 0);
    for (int oc = 0; oc < outch; oc++)
        for (int t = 0; t < taps; t++)
        {
            const float* src = w + ((size_t)oc * taps + t) * inch;
            uint16_t* dst = packed.data() + (size_t)oc * plan->Kp + (size_t)t * cpad;
            for (int ci = 0; ci < inch; ci++) dst[ci] = f32_to_16(src[ci], elemtype);
        }
    NC_CHECK(cudaMalloc(&plan->w_packed, packed.size() * 2));
    NC_CHECK(cudaMemcpyAsync(plan->w_packed, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, stream));
    std::vector<float> bpad((size_t)plan->outch_pad + 256, 0.f);
    if (bias) memcpy(bpad.data(), bias, sizeof(float) * outch);
    NC_CHECK(cudaMalloc((void**)&plan->bias_pad, bpad.size() * sizeof(float)));
    NC_CHECK(cudaMemcpyAsync(plan->bias_pad, bpad.data(), bpad.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    NC_CHECK(cudaStreamSynchronize(stream)); // the staging vectors die at scope exit

    cuuint64_t gdim[2] = {(cuuint64_t)plan->Kp, (cuuint64_t)outch};
    cuuint64_t gstride[1] = {(cuuint64_t)plan->Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)plan->block_k, (cuuint32_t)plan->block_n};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encodeTiled(&plan->tmap_b, dtype_for(elemtype), 2, plan->w_packed, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_for(plan->block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_last_error_msg("cuTensorMapEncodeTiled(weights) failed");
        tc_plan_destroy(plan);
        return -1;
    }
    return 0;
}

void tc_plan_destroy(TcPlan* plan)
{
    if (plan->w_packed) cudaFree(plan->w_packed);
    if (plan->bias_pad) cudaFree(plan->bias_pad);
    plan->w_packed = 0;
    plan->bias_pad = 0;
}

int tc_conv_supported(const TcPlan* plan, const TcConvCall* c)
{
    if (!plan->w_packed) return 0;
    if ((c->in_cpitch & 7) || (c->out_cpitch & 7)) return 0;
    if (((uintptr_t)c->in & 15) || ((uintptr_t)c->out & 15)) return 0;
    if (c->residual && ((c->res_cpitch & 7) || ((uintptr_t)c->residual & 15))) return 0;
    if (c->tiled) return 1;
    // TMA im2col limits for 2 spatial dims: corners in [-128, 127], filter offsets <= 255
    // (cutlass/conv/collective/sm100_implicit_gemm_umma_warpspecialized.hpp can_implement)
    int lower_w = -c->pad_left, lower_h = -c->pad_top;
    int upper_w = c->pad_right - (c->kernel_w - 1) * c->dil_w;
    int upper_h = c->pad_bottom - (c->kernel_h - 1) * c->dil_h;
    if (lower_w < -128 || lower_h < -128 || upper_w < -128 || upper_h < -128 || upper_w > 127 || upper_h > 127) return 0;
    if ((c->kernel_w - 1) * c->dil_w > 255 || (c->kernel_h - 1) * c->dil_h > 255) return 0;
    if (c->stride_w > 8 || c->stride_h > 8) return 0;
    // the traversal must produce exactly outw x outh pixels per image
    int q = (c->inw + upper_w - lower_w - 1) / c->stride_w + 1;
    int pch = (c->inh + upper_h - lower_h - 1) / c->stride_h + 1;
    if (q != c->outw || pch != c->outh) return 0;
    return 1;
}

template<typename T, int BLOCK_N, int BLOCK_K, bool IM2COL>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const tc::Params& p, cudaStream_t stream)
{
    using Plan = tc::SmemPlan<BLOCK_N, BLOCK_K>;
    auto kern = tc::tc_gemm_kernel<T, BLOCK_N, BLOCK_K, IM2COL>;
    static bool attr_set = false;
    if (!attr_set)
    {
        NC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::total));
        attr_set = true;
    }
    long long tiles = ((p.M + tc::BLOCK_M - 1) / tc::BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N);
    int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    kern<<<grid, tc::kNumThreads, Plan::total, stream>>>(ta, tb, p);
    NC_LAUNCH_CHECK();
    return 0;
}

template<typename T, bool IM2COL>
static int dispatch_tc(int block_n, int block_k, const CUtensorMap& ta, const CUtensorMap& tb, const tc::Params& p, cudaStream_t stream)
{
#define NC_TC(BN, BK) \
    if (block_n == BN && block_k == BK) return launch_tc<T, BN, BK, IM2COL>(ta, tb, p, stream)
    NC_TC(256, 64);
    NC_TC(128, 64);
    NC_TC(64, 64);
    NC_TC(32, 64);
    NC_TC(256, 32);
    NC_TC(128, 32);
    NC_TC(64, 32);
    NC_TC(32, 32);
    NC_TC(256, 16);
    NC_TC(128, 16);
    NC_TC(64, 16);
    NC_TC(32, 16);
#undef NC_TC
    set_last_error_msg("tc_gemm: no kernel instance for this tile shape");
    return -1;
}

int tc_conv_forward(const TcPlan* plan, const TcConvCall* c, cudaStream_t stream)
{
    if (!tc_conv_supported(plan, c)) return -1;
    CUtensorMap ta;
    const long long M = (long long)c->n * c->outh * c->outw;
    if (M == 0) return 0;
    if (c->tiled)
    {
        cuuint64_t gdim[2] = {(cuuint64_t)c->inch, (cuuint64_t)M};
        cuuint64_t gstride[1] = {(cuuint64_t)c->in_cpitch * 2};
        cuuint32_t box[2] = {(cuuint32_t)plan->block_k, (cuuint32_t)tc::BLOCK_M};
        cuuint32_t estride[2] = {1, 1};
        CUresult r = g_encodeTiled(&ta, dtype_for(plan->elemtype), 2, (void*)c->in, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle_for(plan->block_k), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            set_last_error_msg("cuTensorMapEncodeTiled(activations) failed");
            return -1;
        }
    }
    else
    {
        cuuint64_t gdim[4] = {(cuuint64_t)c->inch, (cuuint64_t)c->inw, (cuuint64_t)c->inh, (cuuint64_t)c->n};
        cuuint64_t gstride[3] = {(cuuint64_t)c->in_cpitch * 2, (cuuint64_t)c->in_cpitch * 2 * c->inw, (cuuint64_t)c->in_cpitch * 2 * c->inw * c->inh};
        int lower[2] = {-c->pad_left, -c->pad_top};
        int upper[2] = {c->pad_right - (c->kernel_w - 1) * c->dil_w, c->pad_bottom - (c->kernel_h - 1) * c->dil_h};
        cuuint32_t estride[4] = {1, (cuuint32_t)c->stride_w, (cuuint32_t)c->stride_h, 1};
        CUresult r = g_encodeIm2col(&ta, dtype_for(plan->elemtype), 4, (void*)c->in, gdim, gstride, lower, upper, (cuuint32_t)plan->block_k,
                                    (cuuint32_t)tc::BLOCK_M, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(plan->block_k),
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            set_last_error_msg("cuTensorMapEncodeIm2col(activations) failed");
            return -1;
        }
        // Drivers up to 13.1 mis-handle im2col descriptors of tensors smaller than 128 KiB unless bit 21 of the
        // second descriptor word is cleared (same workaround as cute/atom/copy_traits_sm90_im2col.hpp).
        if (g_driver_version <= 13010)
        {
            size_t bytes = (size_t)c->n * c->inh * c->inw * c->in_cpitch * 2;
            if (bytes < 131072) reinterpret_cast<uint64_t*>(&ta)[1] &= ~(1ull << 21);
        }
    }

    tc::Params p;
    p.M = M;
    p.N = plan->outch;
    p.num_k_blocks = plan->num_k_blocks;
    p.cblocks = plan->cblocks;
    p.taps_w = c->kernel_w;
    p.outw = c->outw;
    p.outh = c->outh;
    p.stride_w = c->stride_w;
    p.stride_h = c->stride_h;
    p.dil_w = c->dil_w;
    p.dil_h = c->dil_h;
    p.pad_left = c->pad_left;
    p.pad_top = c->pad_top;
    p.bias = plan->bias_pad;
    p.out = c->out;
    p.out_cpitch = c->out_cpitch;
    p.residual = c->residual;
    p.res_cpitch = c->res_cpitch;
    p.act_type = c->act_type;
    p.act_p0 = c->act_p0;
    p.act_p1 = c->act_p1;

    if (plan->elemtype == NCNN_CUDA_BF16)
        return c->tiled ? dispatch_tc<__nv_bfloat16, false>(plan->block_n, plan->block_k, ta, plan->tmap_b, p, stream)
                        : dispatch_tc<__nv_bfloat16, true>(plan->block_n, plan->block_k, ta, plan->tmap_b, p, stream);
    return c->tiled ? dispatch_tc<__half, false>(plan->block_n, plan->block_k, ta, plan->tmap_b, p, stream)
                    : dispatch_tc<__half, true>(plan->block_n, plan->block_k, ta, plan->tmap_b, p, stream);
}

} // namespace ncnn_cuda
