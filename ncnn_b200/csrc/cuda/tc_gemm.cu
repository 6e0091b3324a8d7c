// tc_gemm.cu -- host side of the tcgen05 implicit-GEMM path: weight re-packing (the create_pipeline step of
// Convolution / InnerProduct, cf. src/layer/x86/convolution_x86.cpp:279-500 in the reference), TMA descriptor
// encoding (tiled for weights, outputs, residuals and 1x1 convs; im2col mode for general convs; overlapping-stride
// tiled maps for small-channel stems) and the persistent-kernel launch.
#include "tc_gemm.cuh"
#include "stem_pool.cuh"

#include <mutex>
#include <stdlib.h>
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
    // NCNN_B200_TC_BN caps the tile width (experiments on wave quantisation)
    static int cap = -1;
    if (cap < 0)
    {
        const char* e = getenv("NCNN_B200_TC_BN");
        cap = e ? atoi(e) : 256;
    }
    if (outch > 128 && cap >= 256) return 256;
    if (outch > 64 && cap >= 128) return 128;
    if (outch > 32 && cap >= 64) return 64;
    if (cap < 64) return 32;
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

static CUtensorMapSwizzle swizzle_for_bytes(int row_bytes)
{
    return row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static CUtensorMapSwizzle swizzle_for(int block_k)
{
    return swizzle_for_bytes(block_k * 2);
}

static CUtensorMapDataType dtype_for(int elemtype)
{
    return elemtype == NCNN_CUDA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

int tma_encode_tiled_plain(CUtensorMap* map, int elemtype, int rank, const void* ptr, const unsigned long long* gdim, const unsigned long long* gstride_bytes,
                           const unsigned int* box, int oob_nan)
{
    if (!tc_available() || rank < 1 || rank > 5) return -1;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++)
    {
        gd[i] = gdim[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i < rank - 1) gs[i] = gstride_bytes[i];
    }
    CUtensorMapDataType dt = elemtype == NCNN_CUDA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dtype_for(elemtype);
    // 128-byte promotion: a tile row can be as narrow as 32 bytes of a pixel's channel run; wider promotion only helps when
    // the neighbouring channel blocks are consumed soon after (they are: channel block is the fastest tile index)
    CUresult r = g_encodeTiled(map, dt, (cuuint32_t)rank, (void*)ptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, oob_nan ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

static int encode_weights(CUtensorMap* map, int elemtype, void* w, int Kp, int outch, int block_k, int block_n)
{
    cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)outch};
    cuuint64_t gstride[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)block_n};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encodeTiled(map, dtype_for(elemtype), 2, w, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

int tc_plan_create(TcPlan* plan, int elemtype, int inch, int outch, int kernel_w, int kernel_h, int stride_w, int dil_w, int pad_left, const float* w,
                   const float* bias, cudaStream_t stream)
{
    memset(plan, 0, sizeof(*plan));
    if (!tc_available()) return -1;
    const int taps = kernel_w * kernel_h;
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

    // ---- A_ROWS variant for small-channel stems (see tc_gemm.cuh): needs the layer's explicit left padding
    std::vector<uint16_t> rows;
    if (inch <= 8 && taps > 1 && dil_w == 1 && pad_left >= 0 && kernel_w <= 16)
    {
        int cp, shift, wp;
        if (inch <= 4 && (stride_w % 2) == 0)
        {
            cp = 4;                     // 8-byte pixels: the window start (stride_w * ox - pad_left - shift pixels) must be 16-byte aligned
            shift = pad_left & 1;
            wp = ((kernel_w + shift + 3) / 4) * 4;
        }
        else
        {
            cp = 8;                     // 16-byte pixels: any start is aligned
            shift = 0;
            wp = ((kernel_w + 1) / 2) * 2;
        }
        const int krow = wp * cp;
        if (krow == 16 || krow == 32 || krow == 64 || krow == 128)
        {
            plan->rows_cp = cp;
            plan->rows_wp = wp;
            plan->rows_shift = shift;
            plan->rows_block_k = krow > 64 ? 64 : krow;
            plan->rows_cblocks = krow / plan->rows_block_k;
            plan->rows_Kp = kernel_h * krow;
            rows.assign((size_t)outch * plan->rows_Kp, 0);
            for (int oc = 0; oc < outch; oc++)
                for (int ky = 0; ky < kernel_h; ky++)
                    for (int kx = 0; kx < kernel_w; kx++)
                    {
                        const float* src = w + ((size_t)oc * taps + ky * kernel_w + kx) * inch;
                        uint16_t* dst = rows.data() + (size_t)oc * plan->rows_Kp + (size_t)ky * krow + (size_t)(kx + shift) * cp;
                        for (int ci = 0; ci < inch; ci++) dst[ci] = f32_to_16(src[ci], elemtype);
                    }
            NC_CHECK(cudaMalloc(&plan->w_rows, rows.size() * 2));
            NC_CHECK(cudaMemcpyAsync(plan->w_rows, rows.data(), rows.size() * 2, cudaMemcpyHostToDevice, stream));
            plan->rows_ok = 1;
        }
    }
    NC_CHECK(cudaStreamSynchronize(stream)); // the staging vectors die at scope exit

    if (encode_weights(&plan->tmap_b, elemtype, plan->w_packed, plan->Kp, outch, plan->block_k, plan->block_n) != 0)
    {
        set_last_error_msg("cuTensorMapEncodeTiled(weights) failed");
        tc_plan_destroy(plan);
        return -1;
    }
    // CTA-pair variant (cta_group::2): each CTA of the pair stages block_n / 2 weight rows per k-block
    plan->pair_ok = 0;
    if (plan->block_k == 64 && plan->block_n >= 128 &&
            encode_weights(&plan->tmap_b_half, elemtype, plan->w_packed, plan->Kp, outch, plan->block_k, plan->block_n / 2) == 0)
        plan->pair_ok = 1;
    // 64-wide tiles for calls with few output pixels (InnerProduct on a batch of a few hundred rows): more CTAs share the weight stream
    plan->narrow_ok = 0;
    if (plan->block_k == 64 && plan->block_n > 64 && encode_weights(&plan->tmap_b_64, elemtype, plan->w_packed, plan->Kp, outch, plan->block_k, 64) == 0) plan->narrow_ok = 1;
    if (plan->rows_ok && encode_weights(&plan->tmap_b_rows, elemtype, plan->w_rows, plan->rows_Kp, outch, plan->rows_block_k, plan->block_n) != 0)
    {
        cudaFree(plan->w_rows);
        plan->w_rows = 0;
        plan->rows_ok = 0;
    }
    return 0;
}

void tc_plan_destroy(TcPlan* plan)
{
    if (plan->w_packed) cudaFree(plan->w_packed);
    if (plan->bias_pad) cudaFree(plan->bias_pad);
    if (plan->w_rows) cudaFree(plan->w_rows);
    plan->w_packed = 0;
    plan->bias_pad = 0;
    plan->w_rows = 0;
}

// ---------------------------------------------------------------- A_ROWS geometry
// The small-channel stem variant (tc_gemm.cuh, A_ROWS) reads a zero-padded copy of the input, [n][Hp][Wpitch][Cp] with
// Cp = 4 (stride 2) or 8 (stride 1) 16-bit channels per pixel so that the windows of consecutive output columns start 16
// bytes apart; rows_pack_kernel writes it once per call.
struct RowsGeom
{
    int Lp, Hp, Wpitch; // physical left pad, padded rows, padded row pitch (pixels)
    int seg_bytes;      // one tile's row segment: 127 windows' starts + one window
    size_t bytes;
};

static bool rows_applicable(const TcPlan* plan, const TcConvCall* c, RowsGeom* g)
{
    if (!plan->rows_ok || c->tiled || c->residual) return false;
    if (c->dil_w != 1 || plan->rows_cblocks != 1) return false;
    const int cp = plan->rows_cp, wp = plan->rows_wp;
    if (c->stride_w * cp * 2 != 16) return false;       // consecutive windows exactly one 16-byte core-matrix row apart
    if (plan->outch > plan->block_n) return false;      // one n-block: the weights stay resident
    if (cp == 4 && ((c->pad_left + plan->rows_shift) & 1)) return false;
    g->Lp = c->pad_left + plan->rows_shift;
    int need_w = c->stride_w * (c->outw - 1) + wp;
    int wpitch = g->Lp + c->inw > need_w ? g->Lp + c->inw : need_w;
    // rows start on 128-byte lines: a cp.async.bulk row segment whose source is only 16-byte aligned moves at ~14 B/cycle/SM
    // (measured: the ResNet stem was paced by its seven 2 KB copies per tile, not by the MMAs or the epilogue)
    wpitch = (wpitch + 64 / cp - 1) / (64 / cp) * (64 / cp);
    int need_h = c->stride_h * (c->outh - 1) + (c->kernel_h - 1) * c->dil_h + 1;
    int hp = c->pad_top + c->inh > need_h ? c->pad_top + c->inh : need_h;
    g->Wpitch = wpitch;
    g->Hp = hp;
    g->seg_bytes = 16 * (tc::BLOCK_M - 1) + wp * cp * 2;
    // the last chunk of a row reads up to one segment past its first column: slack behind the last row
    g->bytes = (size_t)c->n * hp * wpitch * cp * 2 + (size_t)g->seg_bytes + 16 * tc::BLOCK_M;
    if ((long long)c->n * c->outh > 0x7fffffffLL) return false;
    if ((long long)hp * wpitch * cp * 2 > 0x7fffffffLL) return false;
    return true;
}

size_t tc_conv_workspace(const TcPlan* plan, const TcConvCall* c)
{
    RowsGeom g;
    if (!plan->w_packed || !rows_applicable(plan, c, &g)) return 0;
    return g.bytes;
}

// NHWC blob (cpitch channels per pixel) -> zero-padded [n][Hp][Wpitch][CP] copy.  One CTA per padded row at a time (one division
// per row, none per pixel); a thread produces one 16-byte unit of the row -- two 4-channel pixels or one 8-channel pixel -- from
// 8 / 16-byte loads of the source pixels' leading channels; lanes of the blob beyond `inch` are masked to zero.
template<int CP>
__global__ void __launch_bounds__(128) rows_pack_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int n, int inh, int inw, int inch, int in_cpitch,
                                                        int Hp, int Wpitch, int Lp, int pad_top)
{
    NC_PDL_PROLOGUE();
    const int rows = n * Hp;
    const int units = Wpitch * CP / 8; // 16-byte units per padded row
    const unsigned long long keep = inch >= 4 ? ~0ull : ((1ull << (16 * inch)) - 1ull);
    for (int row = blockIdx.x; row < rows; row += gridDim.x)
    {
        const int b = row / Hp;
        const int sy = row - b * Hp - pad_top;
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)row * Wpitch * CP);
        if (sy < 0 || sy >= inh)
        {
            for (int u = threadIdx.x; u < units; u += blockDim.x) dst[u] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const uint16_t* srow = in + ((long long)b * inh + sy) * (long long)inw * in_cpitch;
        for (int u = threadIdx.x; u < units; u += blockDim.x)
        {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (CP == 4)
            {
                const int x0 = 2 * u - Lp, x1 = x0 + 1;
                unsigned long long p0 = 0ull, p1 = 0ull;
                if (x0 >= 0 && x0 < inw) p0 = __ldg(reinterpret_cast<const unsigned long long*>(srow + (long long)x0 * in_cpitch)) & keep;
                if (x1 >= 0 && x1 < inw) p1 = __ldg(reinterpret_cast<const unsigned long long*>(srow + (long long)x1 * in_cpitch)) & keep;
                v.x = (uint32_t)p0;
                v.y = (uint32_t)(p0 >> 32);
                v.z = (uint32_t)p1;
                v.w = (uint32_t)(p1 >> 32);
            }
            else
            {
                const int x0 = u - Lp;
                if (x0 >= 0 && x0 < inw)
                {
                    v = __ldg(reinterpret_cast<const uint4*>(srow + (long long)x0 * in_cpitch));
                    if (inch < 8)
                    {
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            if (2 * k >= inch) w[k] = 0u;
                            else if (2 * k + 1 >= inch) w[k] &= 0xFFFFu;
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            dst[u] = v;
        }
    }
}

static int launch_rows_pack(int cp, const TcConvCall* c, int Hp, int Wpitch, int Lp, cudaStream_t stream)
{
    const int rows = c->n * Hp;
    const int grid = rows < sm_count() * 16 ? rows : sm_count() * 16;
    if (cp == 4)
        NC_PDL_LAUNCH(rows_pack_kernel<4>, grid, 128, 0, stream, (const uint16_t*)c->in, (uint16_t*)c->workspace, c->n, c->inh, c->inw, c->inch, c->in_cpitch, Hp, Wpitch, Lp,
                      c->pad_top);
    else
        NC_PDL_LAUNCH(rows_pack_kernel<8>, grid, 128, 0, stream, (const uint16_t*)c->in, (uint16_t*)c->workspace, c->n, c->inh, c->inw, c->inch, c->in_cpitch, Hp, Wpitch, Lp,
                      c->pad_top);
    return 0;
}

// ---------------------------------------------------------------- A_SHIFT geometry (tc_gemm.cuh)
struct ShiftGeom
{
    int bw, rows, colstep, chunks_x, tiles_y;
    int stage_bytes, box_bytes, aux_bytes;
};

static int shift_mode()
{
    // NCNN_B200_CONV_SHIFT=0 forces the im2col path (A/B comparisons)
    static int v = -1;
    if (v < 0)
    {
        const char* e = getenv("NCNN_B200_CONV_SHIFT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static bool shift_applicable(const TcPlan* plan, const TcConvCall* c, ShiftGeom* g)
{
    if (!shift_mode() || c->tiled || c->residual) return false;
    if (c->stride_w != 1 || c->stride_h != 1 || c->dil_w != 1 || c->dil_h != 1) return false;
    if (c->kernel_w * c->kernel_h < 2 || c->kernel_w > 8 || c->kernel_h > 8) return false;
    if (plan->block_k != 64 || plan->outch > plan->block_n) return false; // 128-byte pixel rows; one n-block (resident weights)
    if (c->pad_left < 0 || c->pad_top < 0 || c->pad_right < 0 || c->pad_bottom < 0) return false;
    const int pw = c->inw + c->pad_left + c->pad_right;
    if (c->outw != pw - (c->kernel_w - 1) || c->outh != c->inh + c->pad_top + c->pad_bottom - (c->kernel_h - 1)) return false;
    if (pw <= 128)
    {
        g->bw = pw;
        g->colstep = pw - (c->kernel_w - 1); // = outw
        g->chunks_x = 1;
        g->rows = 128 / pw;
    }
    else
    {
        g->bw = 128;
        g->colstep = 128 - (c->kernel_w - 1);
        g->chunks_x = (c->outw + g->colstep - 1) / g->colstep;
        g->rows = 1;
    }
    if (g->rows > c->outh) g->rows = c->outh;
    if (g->rows + c->kernel_h - 1 > 256) return false;
    g->tiles_y = (c->outh + g->rows - 1) / g->rows;
    g->box_bytes = g->bw * (g->rows + c->kernel_h - 1) * 128;
    // the MMA reads 128 rows from the last tap's start: (kh-1)*bw + (kw-1) + 128 pixels
    int px = (c->kernel_h - 1) * g->bw + (c->kernel_w - 1) + 128;
    int need = px * 128 > g->box_bytes ? px * 128 : g->box_bytes;
    g->stage_bytes = (need + 1023) / 1024 * 1024;
    g->aux_bytes = plan->num_k_blocks * plan->block_n * 64 * 2;
    // at least 3 stages next to the resident weights (and the layer's bias vector)
    const int bias_need = (plan->outch_pad * 4 + 1023) / 1024 * 1024;
    const int budget = 227 * 1024 - 512 - bias_need - 1024;
    if ((budget - g->aux_bytes) / g->stage_bytes < 3) return false;
    // utilisation of the 128 MMA rows: below ~60 % the im2col path wins
    const long long useful = (long long)g->rows * g->colstep;
    if (useful * 100 < 128 * 55) return false;
    if ((long long)c->n * g->tiles_y * g->chunks_x > 0x3fffffffLL) return false;
    return true;
}

int tc_conv_supported(const TcPlan* plan, const TcConvCall* c)
{
    if (!plan->w_packed) return 0;
    if ((c->in_cpitch & 7) || (c->out_cpitch & 7)) return 0;
    if (((uintptr_t)c->in & 15) || ((uintptr_t)c->out & 15)) return 0;
    if (c->residual && ((c->res_cpitch & 7) || ((uintptr_t)c->residual & 15))) return 0;
    if ((long long)c->n * c->outh * c->outw > 0x7fffff00LL) return 0;
    if (plan->dual_k1_blocks > 0)
    {
        // a dual plan only runs as the two-operand GEMM it was packed for
        if (!c->tiled || !c->in2 || c->residual || plan->block_k != 64) return 0;
        if ((c->in2_cpitch & 7) || ((uintptr_t)c->in2 & 15) || c->in2_ch != plan->dual_inch2) return 0;
        if (c->in2_stride_w < 1 || c->in2_stride_h < 1 || c->in2_stride_w > 8 || c->in2_stride_h > 8) return 0;
        if ((c->in2_w - 1) / c->in2_stride_w + 1 != c->outw || (c->in2_h - 1) / c->in2_stride_h + 1 != c->outh) return 0;
        return 1;
    }
    if (c->in2) return 0;
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

// launch with the programmatic-stream-serialization attribute AND a cluster of `cluster_x` CTAs (CTA pairs)
template<typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster_x, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned int)cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template<typename T, int BLOCK_N, int BLOCK_K, int AMODE, int CG = 1>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& to, tc::Params& p, long long tiles, cudaStream_t stream)
{
    using Plan = tc::SmemPlan<BLOCK_N, BLOCK_K, CG>;
    static_assert(Plan::stages_for(true) >= 2, "not enough shared memory for a 2-stage pipeline");
    auto kern = tc::tc_gemm_kernel<T, BLOCK_N, BLOCK_K, AMODE, CG>;
    static bool attr_set = false;
    if (!attr_set)
    {
        NC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const bool has_res = p.residual != 0;
    int smem_bytes;
    const int bias_count = ((p.N + BLOCK_N - 1) / BLOCK_N) * BLOCK_N;
    const int bias_need = (bias_count * 4 + 1023) / 1024 * 1024;
    p.bias_smem_bytes = Plan::bias_bytes; // the general modes reserve the worst case
    if (AMODE == tc::A_ROWS)
    {
        const int aux = p.taps_h * Plan::b_bytes + (p.tma_store ? Plan::kStageSlots * Plan::res_slot_bytes : 0); // resident weights (+ output staging slots)
        p.bias_smem_bytes = bias_need <= Plan::bias_bytes ? bias_need : 0;
        p.num_stages = Plan::stages_with_aux(aux, p.rows_stage_bytes, p.bias_smem_bytes);
        smem_bytes = Plan::total_with_aux(aux, p.rows_stage_bytes, p.bias_smem_bytes);
        if (p.num_stages < 2)
        {
            set_last_error_msg("tc_gemm: stem weights do not fit in shared memory");
            return -1;
        }
    }
    else if (AMODE == tc::A_SHIFT)
    {
        const int aux = p.num_k_blocks * Plan::b_bytes; // resident weights
        p.bias_smem_bytes = bias_need <= Plan::bias_bytes ? bias_need : 0;
        p.num_stages = Plan::stages_with_aux(aux, p.sh_stage_bytes, p.bias_smem_bytes);
        smem_bytes = Plan::total_with_aux(aux, p.sh_stage_bytes, p.bias_smem_bytes);
    }
    else
    {
        p.num_stages = Plan::stages_for(has_res, p.tma_store != 0);
        smem_bytes = Plan::total_for(has_res, p.tma_store != 0);
    }
    // experiment knobs: NCNN_B200_TC_STAGES caps the ring depth, NCNN_B200_TC_GRID the number of CTAs
    static int stage_cap = -1, grid_cap = -1;
    if (stage_cap < 0)
    {
        const char* e = getenv("NCNN_B200_TC_STAGES");
        stage_cap = e ? atoi(e) : 64;
        e = getenv("NCNN_B200_TC_GRID");
        grid_cap = e ? atoi(e) : 1 << 20;
    }
    if (p.num_stages > stage_cap && stage_cap >= 2) p.num_stages = stage_cap;
    const int sms = sm_count() < grid_cap ? sm_count() : grid_cap;
    if (CG == 2)
    {
        // `tiles` counts pair tiles: one cluster of two CTAs (the two SMs of a TPC) per tile, persistent
        const int clusters = (int)(tiles < sms / 2 ? tiles : sms / 2);
        NC_CHECK(launch_pdl_cluster(kern, dim3(2 * clusters), dim3(tc::kNumThreads), (size_t)smem_bytes, 2, stream, ta, tb, tr, to, p));
        NC_LAUNCH_CHECK();
        count_tc_launch();
        return 0;
    }
    int grid = (int)(tiles < sms ? tiles : sms);
    NC_CHECK(launch_pdl(kern, dim3(grid), dim3(tc::kNumThreads), (size_t)smem_bytes, stream, ta, tb, tr, to, p));
    NC_LAUNCH_CHECK();
    count_tc_launch();
    return 0;
}

// the CTA-pair instances: 64-element k-blocks, 128 / 256-wide tiles, tiled and im2col operands
template<typename T, int AMODE>
static int dispatch_tc_pair(int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& to, tc::Params& p, long long tiles,
                            cudaStream_t stream)
{
    if (block_n == 256) return launch_tc<T, 256, 64, AMODE, 2>(ta, tb, tr, to, p, tiles, stream);
    if (block_n == 128) return launch_tc<T, 128, 64, AMODE, 2>(ta, tb, tr, to, p, tiles, stream);
    set_last_error_msg("tc_gemm: no CTA-pair kernel instance for this tile shape");
    return -1;
}

template<typename T, int AMODE>
static int dispatch_tc(int block_n, int block_k, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& to, tc::Params& p,
                       long long tiles, cudaStream_t stream)
{
#define NC_TC(BN, BK) \
    if (block_n == BN && block_k == BK) return launch_tc<T, BN, BK, AMODE>(ta, tb, tr, to, p, tiles, stream)
    NC_TC(256, 64);
    NC_TC(128, 64);
    NC_TC(64, 64);
    NC_TC(32, 64);
    if constexpr (AMODE != tc::A_SHIFT)
    {
    NC_TC(256, 32);
    NC_TC(128, 32);
    NC_TC(64, 32);
    NC_TC(32, 32);
    NC_TC(256, 16);
    NC_TC(128, 16);
    NC_TC(64, 16);
    NC_TC(32, 16);
    }
#undef NC_TC
    set_last_error_msg("tc_gemm: no kernel instance for this tile shape");
    return -1;
}

// (channels, columns, rows) map of a channel-innermost blob for the epilogue's TMA store / residual load
static int encode_out(CUtensorMap* map, int elemtype, const void* ptr, int C, int cpitch, long long cols, long long rows, int epi_n, int box_rows = tc::BLOCK_M)
{
    cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[2] = {(cuuint64_t)cpitch * 2, (cuuint64_t)cpitch * 2 * (cuuint64_t)cols};
    cuuint32_t box[3] = {(cuuint32_t)epi_n, (cuuint32_t)box_rows, 1};
    cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = g_encodeTiled(map, dtype_for(elemtype), 3, (void*)ptr, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(epi_n * 2),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

int tc_conv_forward(const TcPlan* plan, const TcConvCall* c, cudaStream_t stream)
{
    if (!tc_conv_supported(plan, c)) return -1;
    CUtensorMap ta, tr;
    const long long M = (long long)c->n * c->outh * c->outw;
    if (M == 0) return 0;
    // few output pixels and wide tiles leave most SMs idle (ResNet-50 fc1000 on 256 rows: 8 tiles of 128 x 256 for 148 SMs, each
    // streaming 1/4 of the weights): 64-wide tiles give four times the CTAs for the same bytes
    int plan_block_n = plan->block_n;
    const CUtensorMap* tb_plain = &plan->tmap_b;
    bool allow_pair = true;
    if (plan->narrow_ok && c->tiled && !c->residual && plan->dual_k1_blocks == 0)
    {
        const long long wide_tiles = ((M + tc::BLOCK_M - 1) / tc::BLOCK_M) * ((plan->outch + plan->block_n - 1) / plan->block_n);
        if (wide_tiles * 2 <= sm_count())
        {
            plan_block_n = 64;
            tb_plain = &plan->tmap_b_64;
            allow_pair = false;
        }
    }
    const int epi_n = plan_block_n < 64 ? plan_block_n : 64;

    RowsGeom rg;
    const bool use_rows = rows_applicable(plan, c, &rg) && c->workspace && c->workspace_size >= rg.bytes;
    int amode = c->tiled ? tc::A_TILED : tc::A_IM2COL;
    int block_k = plan->block_k;
    const CUtensorMap* tb = tb_plain;

    tc::Params p;
    memset(&p, 0, sizeof(p));
    ShiftGeom sg;
    const bool use_shift = !use_rows && shift_applicable(plan, c, &sg);
    if (use_shift)
    {
        cuuint64_t gdim[4] = {(cuuint64_t)c->inch, (cuuint64_t)c->inw, (cuuint64_t)c->inh, (cuuint64_t)c->n};
        cuuint64_t gstride[3] = {(cuuint64_t)c->in_cpitch * 2, (cuuint64_t)c->in_cpitch * 2 * c->inw, (cuuint64_t)c->in_cpitch * 2 * c->inw * c->inh};
        cuuint32_t box[4] = {64u, (cuuint32_t)sg.bw, (cuuint32_t)(sg.rows + c->kernel_h - 1), 1u};
        cuuint32_t estride[4] = {1, 1, 1, 1};
        CUresult r = g_encodeTiled(&ta, dtype_for(plan->elemtype), 4, (void*)c->in, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
        {
            amode = tc::A_SHIFT;
            p.sh_bw = sg.bw;
            p.sh_rows = sg.rows;
            p.sh_colstep = sg.colstep;
            p.sh_chunks_x = sg.chunks_x;
            p.sh_tiles_y = sg.tiles_y;
            p.sh_stage_bytes = sg.stage_bytes;
            p.sh_box_bytes = sg.box_bytes;
            p.div_sh_chunks_x = make_fastdiv((unsigned int)sg.chunks_x);
            p.div_sh_tiles_y = make_fastdiv((unsigned int)sg.tiles_y);
            p.div_sh_bw = make_fastdiv((unsigned int)sg.bw);
        }
    }
    if (use_rows)
    {
        const int cp = plan->rows_cp;
        // zero-padded small-channel copy of the input
        if (launch_rows_pack(cp, c, rg.Hp, rg.Wpitch, rg.Lp, stream) != 0) return -100;
        amode = tc::A_ROWS;
        block_k = plan->rows_block_k;
        tb = &plan->tmap_b_rows;
        ta = plan->tmap_b_rows; // unused by this mode (the A operand is a plain bulk copy)
        p.cblocks = 1;
        p.num_k_blocks = c->kernel_h;
        p.chunks_per_row = (c->outw + tc::BLOCK_M - 1) / tc::BLOCK_M;
        p.rows_src = (const unsigned char*)c->workspace;
        p.rows_row_bytes = rg.Wpitch * cp * 2;
        p.rows_img_bytes = (long long)rg.Hp * rg.Wpitch * cp * 2;
        p.rows_seg_bytes = rg.seg_bytes;
        p.rows_seg_pitch = (rg.seg_bytes + 127) / 128 * 128;
        p.rows_stage_bytes = ((c->kernel_h * p.rows_seg_pitch + 1023) / 1024) * 1024;
    }
    if (amode == tc::A_SHIFT)
    {
        // (descriptor encoded above)
    }
    else if (amode == tc::A_TILED)
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
    else if (amode == tc::A_IM2COL)
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
    // (the second operand's map travels in the residual slot of the kernel's parameters: a dual plan never has a fused residual)
    CUtensorMap ta2 = ta;
    if (plan->dual_k1_blocks > 0)
    {
        p.nk2 = plan->num_k_blocks - plan->dual_k1_blocks;
        p.a2_stride_w = c->in2_stride_w;
        p.a2_stride_h = c->in2_stride_h;
        CUresult r;
        if (c->in2_stride_w == 1 && c->in2_stride_h == 1)
        {
            cuuint64_t gdim[2] = {(cuuint64_t)c->in2_ch, (cuuint64_t)M};
            cuuint64_t gstride[1] = {(cuuint64_t)c->in2_cpitch * 2};
            cuuint32_t box[2] = {64u, (cuuint32_t)tc::BLOCK_M};
            cuuint32_t estride[2] = {1, 1};
            r = g_encodeTiled(&ta2, dtype_for(plan->elemtype), 2, (void*)c->in2, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        else
        {
            // 1x1 window walked with the shortcut's stride: TMA im2col mode, upper corner = the (negative) far-side remainder
            p.a2_im2col = 1;
            cuuint64_t gdim[4] = {(cuuint64_t)c->in2_ch, (cuuint64_t)c->in2_w, (cuuint64_t)c->in2_h, (cuuint64_t)c->n};
            cuuint64_t gstride[3] = {(cuuint64_t)c->in2_cpitch * 2, (cuuint64_t)c->in2_cpitch * 2 * c->in2_w, (cuuint64_t)c->in2_cpitch * 2 * c->in2_w * c->in2_h};
            int lower[2] = {0, 0};
            int upper[2] = {(c->outw - 1) * c->in2_stride_w + 1 - c->in2_w, (c->outh - 1) * c->in2_stride_h + 1 - c->in2_h};
            cuuint32_t estride[4] = {1, (cuuint32_t)c->in2_stride_w, (cuuint32_t)c->in2_stride_h, 1};
            r = g_encodeIm2col(&ta2, dtype_for(plan->elemtype), 4, (void*)c->in2, gdim, gstride, lower, upper, 64u, (cuuint32_t)tc::BLOCK_M, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS && g_driver_version <= 13010)
            {
                size_t bytes = (size_t)c->n * c->in2_h * c->in2_w * c->in2_cpitch * 2;
                if (bytes < 131072) reinterpret_cast<uint64_t*>(&ta2)[1] &= ~(1ull << 21);
            }
        }
        if (r != CUDA_SUCCESS)
        {
            set_last_error_msg("cuTensorMapEncode(second operand) failed");
            return -1;
        }
    }
    if (amode != tc::A_ROWS)
    {
        p.cblocks = plan->cblocks;
        p.num_k_blocks = plan->num_k_blocks;
        p.chunks_per_row = 1;
    }

    // residual map (the output itself leaves through per-lane vector stores, no descriptor)
    const long long cols = amode == tc::A_ROWS ? c->outw : M;
    const long long rows = amode == tc::A_ROWS ? (long long)c->n * c->outh : 1;
    if (c->residual)
    {
        if (encode_out(&tr, plan->elemtype, c->residual, plan->outch, c->res_cpitch, cols, rows, epi_n) != 0)
        {
            set_last_error_msg("cuTensorMapEncodeTiled(residual) failed");
            return -1;
        }
    }
    else
        tr = plan->dual_k1_blocks > 0 ? ta2 : ta;
    // output map of the TMA-store epilogue: per-warp boxes of 32 rows x epi_n channels; the channel extent is the blob's padded
    // run (padding lanes up to the next 16-byte unit may be written, as by the per-lane stores)
    CUtensorMap to;
    static int tma_store_mode = -1;
    if (tma_store_mode < 0)
    {
        const char* e = getenv("NCNN_B200_TC_TMASTORE");
        tma_store_mode = e ? atoi(e) : 0;
    }
    p.tma_store = 0;
    to = ta;
    if (tma_store_mode && amode != tc::A_SHIFT)
    {
        int cn = (plan->outch + 7) & ~7;
        if (cn > c->out_cpitch) cn = c->out_cpitch;
        if (encode_out(&to, plan->elemtype, c->out, cn, c->out_cpitch, cols, rows, epi_n, 32) == 0)
            p.tma_store = 1;
        else
            to = ta;
    }

    p.M = M;
    p.N = plan->outch;
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
    p.v8_ok = (c->out_cpitch % 16 == 0) && (((uintptr_t)c->out & 31) == 0);
    p.taps_h = c->tiled ? 1 : c->kernel_h;
    p.div_n_blocks = make_fastdiv((unsigned int)((plan->outch + plan_block_n - 1) / plan_block_n));
    p.div_opix = make_fastdiv((unsigned int)(c->outw * c->outh > 0 ? c->outw * c->outh : 1));
    p.div_outw = make_fastdiv((unsigned int)(c->outw > 0 ? c->outw : 1));
    p.div_chunks = make_fastdiv((unsigned int)p.chunks_per_row);
    p.div_outh = make_fastdiv((unsigned int)(c->outh > 0 ? c->outh : 1));

    // CTA pairs (cta_group::2): halves the weight bytes every SM pulls through L2 -> SM, the limiter of the 128-row tiles
    // (NCNN_B200_TC_PAIR=0 turns them off, =1 (default) uses them wherever an instance exists and the layer has >= 2 m-blocks)
    static int pair_mode = -1;
    if (pair_mode < 0)
    {
        const char* e = getenv("NCNN_B200_TC_PAIR");
        pair_mode = e ? atoi(e) : 1;
    }
    const bool use_pair = allow_pair && pair_mode != 0 && plan->pair_ok && (amode == tc::A_TILED || amode == tc::A_IM2COL) && block_k == 64 && M > tc::BLOCK_M;
    if (use_pair)
    {
        const long long pair_tiles = ((M + tc::BLOCK_M - 1) / tc::BLOCK_M + 1) / 2 * ((plan->outch + plan->block_n - 1) / plan->block_n);
        if (plan->elemtype == NCNN_CUDA_BF16)
            return amode == tc::A_TILED ? dispatch_tc_pair<__nv_bfloat16, tc::A_TILED>(plan->block_n, ta, plan->tmap_b_half, tr, to, p, pair_tiles, stream)
                                        : dispatch_tc_pair<__nv_bfloat16, tc::A_IM2COL>(plan->block_n, ta, plan->tmap_b_half, tr, to, p, pair_tiles, stream);
        return amode == tc::A_TILED ? dispatch_tc_pair<__half, tc::A_TILED>(plan->block_n, ta, plan->tmap_b_half, tr, to, p, pair_tiles, stream)
                                    : dispatch_tc_pair<__half, tc::A_IM2COL>(plan->block_n, ta, plan->tmap_b_half, tr, to, p, pair_tiles, stream);
    }

    const long long m_blocks = amode == tc::A_ROWS ? rows * p.chunks_per_row
                               : (amode == tc::A_SHIFT ? (long long)c->n * p.sh_tiles_y * p.sh_chunks_x : (M + tc::BLOCK_M - 1) / tc::BLOCK_M);
    const long long tiles = m_blocks * ((plan->outch + plan_block_n - 1) / plan_block_n);

#define NC_MODE(T)                                                                                                            \
    if (amode == tc::A_SHIFT) return dispatch_tc<T, tc::A_SHIFT>(plan_block_n, block_k, ta, *tb, tr, to, p, tiles, stream);  \
    if (amode == tc::A_TILED) return dispatch_tc<T, tc::A_TILED>(plan_block_n, block_k, ta, *tb, tr, to, p, tiles, stream);  \
    if (amode == tc::A_IM2COL) return dispatch_tc<T, tc::A_IM2COL>(plan_block_n, block_k, ta, *tb, tr, to, p, tiles, stream); \
    return dispatch_tc<T, tc::A_ROWS>(plan_block_n, block_k, ta, *tb, tr, to, p, tiles, stream)
    if (plan->elemtype == NCNN_CUDA_BF16)
    {
        NC_MODE(__nv_bfloat16);
    }
    NC_MODE(__half);
#undef NC_MODE
}

// ---------------------------------------------------------------- stem convolution + 3x3 s2 max pooling in one kernel (stem_pool.cuh)
int tc_stem_pool_supported(const TcPlan* plan, const TcConvCall* c, const TcPoolCall* pc)
{
    RowsGeom rg;
    if (!tc_conv_supported(plan, c) || !rows_applicable(plan, c, &rg)) return 0;
    if (plan->rows_cp != 4 || plan->block_n != 64 || plan->outch > 64 || c->outw > tc::BLOCK_M || c->dil_h != 1) return 0;
    if (c->act_type != 0 && c->act_type != 1) return 0; // the activation must commute with max: none or ReLU
    if (pc->pad_left < 0 || pc->pad_left > 2 || pc->pad_top < 0 || pc->pad_top > 2 || pc->pw <= 0 || pc->ph <= 0) return 0;
    // every pooling window must hold at least one conv element
    if (2 * (pc->pw - 1) - pc->pad_left >= c->outw || 2 * (pc->ph - 1) - pc->pad_top >= c->outh) return 0;
    if ((pc->out_cpitch & 7) || ((uintptr_t)pc->out & 15)) return 0;
    if ((long long)c->n * pc->ph > 0x3fffffffLL) return 0;
    return 1;
}

template<typename T, int BLOCK_K>
static int launch_stem_pool(const CUtensorMap& tb, tc::StemPoolParams& p, cudaStream_t stream)
{
    auto kern = tc::stem_pool_kernel<T, BLOCK_K>;
    static bool attr_set = false;
    if (!attr_set)
    {
        NC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int b_bytes = tc::kStemN * BLOCK_K * 2;
    const int fixed = ((p.taps_h * b_bytes + 1023) & ~1023) + tc::kStemRowSlots * tc::kStemRowSlotBytes + tc::kStemN * 4 + 512 + 1024;
    int stages = (227 * 1024 - fixed) / p.rows_stage_bytes;
    if (stages > 16) stages = 16;
    if (stages < 2)
    {
        set_last_error_msg("stem_pool: the filter rows do not fit in shared memory");
        return -1;
    }
    p.num_stages = stages;
    const int smem_bytes = stages * p.rows_stage_bytes + fixed;
    const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
    NC_CHECK(launch_pdl(kern, dim3(grid), dim3(tc::kStemThreads), (size_t)smem_bytes, stream, tb, p));
    NC_LAUNCH_CHECK();
    count_tc_launch();
    return 0;
}

int tc_stem_pool_forward(const TcPlan* plan, const TcConvCall* c, const TcPoolCall* pc, cudaStream_t stream)
{
    RowsGeom rg;
    if (!tc_stem_pool_supported(plan, c, pc) || !rows_applicable(plan, c, &rg)) return -1;
    if (!c->workspace || c->workspace_size < rg.bytes) return -1;
    if (launch_rows_pack(plan->rows_cp, c, rg.Hp, rg.Wpitch, rg.Lp, stream) != 0) return -100;
    tc::StemPoolParams p;
    memset(&p, 0, sizeof(p));
    const int cp = plan->rows_cp;
    p.rows_src = (const unsigned char*)c->workspace;
    p.rows_row_bytes = rg.Wpitch * cp * 2;
    p.rows_img_bytes = (long long)rg.Hp * rg.Wpitch * cp * 2;
    p.rows_seg_bytes = rg.seg_bytes;
    // one contiguous copy per tile: filter row ky of a conv row is padded row r * stride_h + ky, the shared-memory pitch is the
    // global row pitch, and the overhang of the last row's 128-window segment comes along (the workspace has the slack)
    p.rows_seg_pitch = p.rows_row_bytes;
    {
        int over = rg.seg_bytes - p.rows_row_bytes;
        if (over < 0) over = 0;
        p.rows_copy_bytes = c->kernel_h * p.rows_row_bytes + (over + 15) / 16 * 16;
    }
    p.rows_stage_bytes = ((p.rows_copy_bytes + 1023) / 1024) * 1024;
    p.taps_h = c->kernel_h;
    p.stride_h = c->stride_h;
    p.outw = c->outw;
    p.outh = c->outh;
    p.N = plan->outch;
    p.bias = plan->bias_pad;
    p.relu = c->act_type == 1;
    p.pad_left = pc->pad_left;
    p.pad_top = pc->pad_top;
    p.pw = pc->pw;
    p.ph = pc->ph;
    p.out = pc->out;
    p.out_cpitch = pc->out_cpitch;
    // bands of pooled rows: one conv row per band is computed twice, fewer bands waste fewer rows but quantise worse over the
    // SMs -- pick the band height with the fewest conv-row tiles on the busiest CTA
    {
        const int sms = sm_count();
        long long best_cost = -1;
        int best_rows = pc->ph;
        for (int bands = 1; bands <= pc->ph && bands <= 64; bands++)
        {
            const int br = (pc->ph + bands - 1) / bands;
            const int nb = (pc->ph + br - 1) / br;
            const long long items = (long long)c->n * nb;
            const long long rounds = (items + sms - 1) / sms;
            const long long cost = rounds * (2 * br + 1);
            if (best_cost < 0 || cost < best_cost)
            {
                best_cost = cost;
                best_rows = br;
            }
        }
        p.band_rows = best_rows;
        p.bands = (pc->ph + best_rows - 1) / best_rows;
        p.num_items = c->n * p.bands;
        p.div_bands = make_fastdiv((unsigned int)p.bands);
    }
    {
        static int dbg = -1;
        if (dbg < 0)
        {
            const char* e = getenv("NCNN_B200_STEM_DBG");
            dbg = e ? atoi(e) : 0;
        }
        p.dbg = dbg;
    }
    const int bk = plan->rows_block_k;
#define NC_STEM(T)                                                                      \
    if (bk == 16) return launch_stem_pool<T, 16>(plan->tmap_b_rows, p, stream);         \
    if (bk == 32) return launch_stem_pool<T, 32>(plan->tmap_b_rows, p, stream);         \
    return launch_stem_pool<T, 64>(plan->tmap_b_rows, p, stream)
    if (plan->elemtype == NCNN_CUDA_BF16)
    {
        NC_STEM(__nv_bfloat16);
    }
    NC_STEM(__half);
#undef NC_STEM
}

} // namespace ncnn_cuda
