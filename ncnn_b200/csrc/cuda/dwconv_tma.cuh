// dwconv_tma.cuh -- the bandwidth path of ConvolutionDepthWise (src/layer/convolutiondepthwise.cpp:181-214 of the
// reference): 3x3 depthwise, stride 1 or 2, dilation 1, zero padding, on channel-innermost blobs.
//
// One persistent CTA per SM walks (channel block, image, tile) work items.  A producer warp keeps a kStages-deep ring of
// input tiles (tile + 1-pixel halo) in shared memory with 4-D tiled TMA loads -- the halo outside the image is the TMA
// unit's out-of-bounds zero fill, so there is no padded copy (reference: copy_make_border in
// ConvolutionDepthWise::make_padding) and no per-tap bounds test.  256 consumer threads each own one output column of
// the tile for a 16-byte channel vector and R consecutive output rows: every input pixel is read from shared memory once
// (LDS.128), converted once and feeds up to three output rows from registers.  The layer's whole filter bank ([9][C] fp32)
// and bias sit in shared memory for the lifetime of the CTA; the loop is filter-column major so that only three taps are
// live in registers at a time (<= 113 registers -> two CTAs per SM, 16 consumer warps hide the LDS/convert latency).
// fp32 accumulation with packed FFMA2, bias first.  Outputs leave as 16-byte vector stores (a warp writes whole
// pixels' channel runs).  Tile indices are decoded with multiply-shift division (no integer divide in the loop).
//
// HBM traffic is the algorithmic minimum (input once, output once); the halo re-reads (<= 1.4x of the input) are L2 hits.
#pragma once
#include "tc_gemm.cuh"

namespace ncnn_cuda {
namespace dwt {

constexpr int kConsumers = 256;
constexpr int kThreads = kConsumers + 32;

using ncnn_cuda::FastDiv;
using ncnn_cuda::make_fastdiv;
using ncnn_cuda::fast_div;

struct Params
{
    int C, outw, outh, n;
    int tiles_x, tiles_y;
    FastDiv div_cblocks, div_image, div_tiles_x;
    int n_spatial; // n * tiles_y * tiles_x
    int num_tiles; // cblocks * n_spatial
    int pad_left, pad_top;
    int out_cpitch;
    long long out_nstep;
    long long out_row_stride; // outw * out_cpitch
    const float* w;    // [9][cpad]
    const float* bias; // [C] or NULL
    int cpad;
    int act_type;
    float act_p0, act_p1;
    int num_stages;
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

// 16 bytes of T in shared memory -> VEC/2 float2
template<typename T>
struct Vec16;
template<>
struct Vec16<__half>
{
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void load(const void* p, float2 (&v)[4])
    {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = __half22float2(h[i]);
    }
    static __device__ __forceinline__ void store(void* p, const float2 (&v)[4])
    {
        uint4 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[i].x, v[i].y);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ void store_clamped(void* p, const float2 (&v)[4], float lo, float hi, bool relu_only)
    {
        uint4 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
        const __half2 l2 = __float2half2_rn(lo), h2 = __float2half2_rn(hi);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            __half2 t = __hmax2(__floats2half2_rn(v[i].x, v[i].y), l2);
            h[i] = relu_only ? t : __hmin2(t, h2);
        }
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template<>
struct Vec16<__nv_bfloat16>
{
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void load(const void* p, float2 (&v)[4])
    {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            // bf16 -> fp32 is a 16-bit shift: integer pipe, not the FMA pipe
            v[i].x = __uint_as_float(w[i] << 16);
            v[i].y = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(void* p, const float2 (&v)[4])
    {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
        *reinterpret_cast<uint4*>(p) = u;
    }
    static __device__ __forceinline__ void store_clamped(void* p, const float2 (&v)[4], float lo, float hi, bool relu_only)
    {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
        const __nv_bfloat162 l2 = __float2bfloat162_rn(lo), h2 = __float2bfloat162_rn(hi);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            __nv_bfloat162 t = __hmax2(__floats2bfloat162_rn(v[i].x, v[i].y), l2);
            h[i] = relu_only ? t : __hmin2(t, h2);
        }
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template<>
struct Vec16<float>
{
    static constexpr int VEC = 4;
    static __device__ __forceinline__ void load(const void* p, float2 (&v)[2])
    {
        const float4 u = *reinterpret_cast<const float4*>(p);
        v[0] = make_float2(u.x, u.y);
        v[1] = make_float2(u.z, u.w);
    }
    static __device__ __forceinline__ void store(void* p, const float2 (&v)[2])
    {
        *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    }
    static __device__ __forceinline__ void store_clamped(void* p, const float2 (&v)[2], float, float, bool)
    {
        store(p, v); // (fp32 blobs clamp in registers)
    }
};

// S stride; CV 16-byte channel vectors per tile; TW output columns per tile (one thread each); TY thread rows; R output rows per thread
template<typename T, int S, int CV, int TW, int TY, int R>
struct Cfg
{
    static constexpr int VEC = Vec16<T>::VEC;
    static constexpr int CB = CV * VEC; // channels per tile
    static constexpr int TH = TY * R;   // output rows per tile
    static constexpr int IW = (TW - 1) * S + 3;
    static constexpr int IH = (TH - 1) * S + 3;
    static constexpr int tile_bytes = IW * IH * CV * 16;
    static constexpr int stage_bytes = (tile_bytes + 127) / 128 * 128;
    // two CTAs per SM: each gets half of the 227 KB
    static constexpr int cta_budget = 111 * 1024;
    static constexpr int fixed_bytes = 128 /*alignment*/ + 128 /*barriers*/;
    // one consumer thread per (channel vector, column, thread row); channel counts that are not a power of two times 8 (144 = 2 x 72)
    // use CV = 9 and leave the last few threads of the CTA idle (they still take part in the barriers)
    static constexpr int kActive = CV * TW * TY;
    static_assert(kActive <= kConsumers && kActive > kConsumers - 32, "consumer mapping must fill all eight warps");
    // this CTA's slice of the filter bank: 9 taps + bias for its CB channels
    static constexpr int w_bytes = 10 * CB * 4;
    static constexpr int stages_fit = (cta_budget - fixed_bytes - w_bytes) / stage_bytes;
    static constexpr int kStages = stages_fit > 4 ? 4 : stages_fit;
    static constexpr int smem_bytes = kStages * stage_bytes + w_bytes + fixed_bytes;
    static_assert(kStages >= 2, "tile too large for a 2-stage ring");
};

template<typename T, int S, int CV, int TW, int TY, int R>
__global__ void __launch_bounds__(kThreads, 2) dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmap_in, T* __restrict__ out, const Params p)
{
    using C = Cfg<T, S, CV, TW, TY, R>;
    constexpr int VEC = C::VEC;
    constexpr int H2 = VEC / 2; // float2 per channel vector
    constexpr int NROWS = (R - 1) * S + 3; // input rows a thread walks
    constexpr int kStages = C::kStages;

    extern __shared__ uint8_t dw_smem_raw[];
    // (pointer arithmetic, not an integer round trip: keeps the shared address space visible to the compiler -> LDS, not generic LD)
    uint8_t* smem = dw_smem_raw + ((128u - (tc::smem_u32(dw_smem_raw) & 127u)) & 127u);
    float* smem_w = reinterpret_cast<float*>(smem + kStages * C::stage_bytes); // [9][CB] taps, then [CB] bias, of this CTA's channel block
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_w + 10 * C::CB);
    uint64_t* empty_bar = full_bar + 4;

    const int tid = threadIdx.x;
    tc::pdl_launch_dependents();
    if (tid == 0)
    {
        tc::prefetch_tmap(&tmap_in);
        for (int i = 0; i < kStages; i++)
        {
            tc::mbar_init(tc::smem_u32(&full_bar[i]), 1);
            tc::mbar_init(tc::smem_u32(&empty_bar[i]), kConsumers / 32);
        }
        tc::fence_barrier_init();
    }
    // A CTA works on ONE channel block for its whole life (blockIdx.x % cblocks) and strides over the spatial tiles, so the
    // CTAs that run side by side cover all channel blocks of the same pixels (shared DRAM pages / L2 lines) and each keeps
    // just its own 9 x CB filter taps + bias resident.
    const int cblocks = (int)p.div_cblocks.d;
    const int cb = blockIdx.x % cblocks;
    const int sp_first = blockIdx.x / cblocks;
    const int sp_step = gridDim.x / cblocks;
    {
        // this CTA's filter slice, 16 bytes per load, every load of a thread issued before its first store (the slice is a few KB:
        // a load -> store -> load chain here costs several microseconds of the 10-20 us a small layer takes)
        constexpr int V4 = 10 * C::CB / 4;                       // float4 units: [10][CB / 4]
        constexpr int PER = (V4 + kThreads - 1) / kThreads;
        float4 tmp[PER];
#pragma unroll
        for (int j = 0; j < PER; j++)
        {
            const int i = tid + j * kThreads;
            const int t = i / (C::CB / 4), c = cb * C::CB + (i - t * (C::CB / 4)) * 4;
            tmp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            // (the last channel block may be partial: its missing channels read as zeros from the TMA and are never stored;
            //  C and cpad are multiples of 4 floats on this path)
            if (i < V4 && c < p.C)
            {
                if (t < 9)
                    tmp[j] = __ldg(reinterpret_cast<const float4*>(p.w + (long long)t * p.cpad + c));
                else if (p.bias)
                    tmp[j] = __ldg(reinterpret_cast<const float4*>(p.bias + c));
            }
        }
#pragma unroll
        for (int j = 0; j < PER; j++)
        {
            const int i = tid + j * kThreads;
            if (i < V4) reinterpret_cast<float4*>(smem_w)[i] = tmp[j];
        }
    }
    __syncthreads();
    tc::pdl_wait(); // the filter slice above is a constant; the previous layer's blob is touched only from here on

    if (tid >= kConsumers)
    {
        // ===================== TMA producer (one lane) =====================
        if (tid == kConsumers)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int sp = sp_first; sp < p.n_spatial; sp += sp_step)
            {
                const int b = fast_div(sp, p.div_image);
                const int t2 = sp - b * (int)p.div_image.d;
                const int tyi = fast_div(t2, p.div_tiles_x);
                const int txi = t2 - tyi * p.tiles_x;
                tc::mbar_wait(tc::smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t fb = tc::smem_u32(&full_bar[stage]);
                tc::mbar_expect_tx(fb, C::tile_bytes);
                tc::tma_load_4d(tc::smem_u32(smem + stage * C::stage_bytes), &tmap_in, fb, cb * C::CB, txi * TW * S - p.pad_left, tyi * C::TH * S - p.pad_top, b);
                if (++stage == kStages)
                {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const bool active = tid < C::kActive;
    const int mt = active ? tid : 0; // idle threads shadow thread 0 (they never store)
    const int cv = mt % CV;
    const int tx = (mt / CV) % TW;
    const int ty = mt / (CV * TW);
    const int lane = tid & 31;

    int stage = 0;
    uint32_t phase = 0;
    // byte offset of this thread's first input pixel inside a staged tile
    const int thread_off = ((ty * R * S) * C::IW + tx * S) * (CV * 16) + cv * 16;
    const int act = p.act_type;

    const int c0 = cb * C::CB + cv * VEC;
    const float* wc = smem_w + cv * VEC;
    for (int sp = sp_first; sp < p.n_spatial; sp += sp_step)
    {
        const int b = fast_div(sp, p.div_image);
        const int t2 = sp - b * (int)p.div_image.d;
        const int tyi = fast_div(t2, p.div_tiles_x);
        const int txi = t2 - tyi * p.tiles_x;
        float2 acc[R][H2];
        {
            float2 bias2[H2];
#pragma unroll
            for (int i = 0; i < H2; i += 2)
            {
                // cpad and c0 are multiples of 4 floats: 16-byte broadcast loads
                const float4 t = *reinterpret_cast<const float4*>(wc + 9 * C::CB + 2 * i);
                bias2[i] = make_float2(t.x, t.y);
                bias2[i + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int i = 0; i < H2; i++) acc[r][i] = bias2[i];
        }

        tc::mbar_wait(tc::smem_u32(&full_bar[stage]), phase);
        const uint8_t* base = smem + stage * C::stage_bytes + thread_off;
#pragma unroll
        for (int kx = 0; kx < 3; kx++)
        {
            // the three taps of this filter column
            float2 w[3][H2];
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int i = 0; i < H2; i += 2)
                {
                    const float4 t = *reinterpret_cast<const float4*>(wc + (ky * 3 + kx) * C::CB + 2 * i);
                    w[ky][i] = make_float2(t.x, t.y);
                    w[ky][i + 1] = make_float2(t.z, t.w);
                }
#pragma unroll
            for (int ii = 0; ii < NROWS; ii++)
            {
                float2 x[H2];
                Vec16<T>::load(base + (ii * C::IW + kx) * (CV * 16), x);
#pragma unroll
                for (int r = 0; r < R; r++)
                {
                    const int ky = ii - r * S;
                    if (ky >= 0 && ky < 3)
                    {
#pragma unroll
                        for (int i = 0; i < H2; i++) acc[r][i] = ffma2(x[i], w[ky][i], acc[r][i]);
                    }
                }
            }
        }
        // this warp is done with the staged tile: hand the slot back before the global stores
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty_bar[stage]));
        if (++stage == kStages)
        {
            stage = 0;
            phase ^= 1;
        }

        // activation on the whole register tile (one uniform branch), then the stores.  For 16-bit blobs ReLU and clip run on the
        // PACKED pairs inside the store (round(clamp(x)) == clamp(round(x), round(lo), round(hi)): rounding is monotonic), one
        // packed min / max per two elements instead of two fp32 ones each
        constexpr bool kPackedAct = sizeof(T) == 2;
        if (kPackedAct && (act == 1 || act == 3))
        {
        }
        else if (act == 1)
        {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int i = 0; i < H2; i++) acc[r][i] = make_float2(fmaxf(acc[r][i].x, 0.f), fmaxf(acc[r][i].y, 0.f));
        }
        else if (act == 3)
        {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int i = 0; i < H2; i++)
                    acc[r][i] = make_float2(fminf(fmaxf(acc[r][i].x, p.act_p0), p.act_p1), fminf(fmaxf(acc[r][i].y, p.act_p0), p.act_p1));
        }
        else if (act != 0)
        {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int i = 0; i < H2; i++)
                    acc[r][i] = make_float2(tc::apply_activation_call(acc[r][i].x, act, p.act_p0, p.act_p1), tc::apply_activation_call(acc[r][i].y, act, p.act_p0, p.act_p1));
        }
        const int ox = txi * TW + tx;
        const int oy0 = tyi * C::TH + ty * R;
        if (active && ox < p.outw && c0 < p.C)
        {
            T* op = out + ((long long)b * p.out_nstep + ((long long)oy0 * p.outw + ox) * p.out_cpitch + c0);
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                if (oy0 + r < p.outh)
                {
                    if (kPackedAct && act == 1)
                        Vec16<T>::store_clamped(op, acc[r], 0.f, 3.0e38f, true);
                    else if (kPackedAct && act == 3)
                        Vec16<T>::store_clamped(op, acc[r], p.act_p0, p.act_p1, false);
                    else
                        Vec16<T>::store(op, acc[r]);
                }
                op += p.out_row_stride;
            }
        }
    }
}

template<typename T, int S, int CV, int TW, int TY, int R>
static int launch_dw_tma(const CUtensorMap& tm, T* out, Params& p, cudaStream_t stream)
{
    using C = Cfg<T, S, CV, TW, TY, R>;
    auto kern = dwconv3x3_tma_kernel<T, S, CV, TW, TY, R>;
    static bool attr_set = false;
    if (!attr_set)
    {
        NC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
        attr_set = true;
    }
    p.tiles_x = (p.outw + TW - 1) / TW;
    p.tiles_y = (p.outh + C::TH - 1) / C::TH;
    const long long n_spatial = (long long)p.n * p.tiles_x * p.tiles_y;
    const int cblocks = (p.C + C::CB - 1) / C::CB;
    if (n_spatial > 0x3fffffffLL) return 1; // caller falls back
    p.n_spatial = (int)n_spatial;
    p.num_tiles = 0;
    p.div_cblocks = make_fastdiv((unsigned int)cblocks);
    p.div_image = make_fastdiv((unsigned int)(p.tiles_x * p.tiles_y));
    p.div_tiles_x = make_fastdiv((unsigned int)p.tiles_x);
    // two resident CTAs per SM; the grid is a whole number of channel-block groups
    long long groups = (2LL * sm_count()) / cblocks;
    if (groups < 1) groups = 1;
    if (groups > n_spatial) groups = n_spatial;
    const int grid = (int)(groups * cblocks);
    NC_CHECK(launch_pdl(kern, dim3(grid), dim3(kThreads), (size_t)C::smem_bytes, stream, tm, out, p));
    NC_LAUNCH_CHECK();
    return 0;
}

// box dims of the staged input tile for a configuration
template<typename T, int S, int CV, int TW, int TY, int R>
static void box_of(unsigned int (&box)[4])
{
    using C = Cfg<T, S, CV, TW, TY, R>;
    box[0] = C::CB;
    box[1] = C::IW;
    box[2] = C::IH;
    box[3] = 1;
}

struct Call
{
    const void* in;
    void* out;
    int elemtype;
    int C, inw, inh, outw, outh, n;
    int stride;
    int pad_left, pad_top;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
    const float* w;
    const float* bias;
    int cpad;
    int act_type;
    float act_p0, act_p1;
};

// 0 launched, 1 not applicable (caller uses the generic kernel), < 0 error
template<typename T>
static int forward(const Call& c, cudaStream_t stream)
{
    constexpr int VEC = Vec16<T>::VEC;
    const int es = (int)sizeof(T);
    if (!tc_available()) return 1;
    if (c.stride != 1 && c.stride != 2) return 1;
    if (c.C % (2 * VEC) != 0) return 1;
    if (((size_t)c.in_cpitch * es) % 16 || ((size_t)c.in_nstep * es) % 16 || ((uintptr_t)c.in & 15)) return 1;
    if (((size_t)c.out_cpitch * es) % 16 || ((size_t)c.out_nstep * es) % 16 || ((uintptr_t)c.out & 15)) return 1;
    if (c.pad_left < 0 || c.pad_top < 0 || c.pad_left > 64 || c.pad_top > 64) return 1;
    const int cv = (c.C % (8 * VEC) == 0) ? 8 : ((c.C % (4 * VEC) == 0) ? 4 : 2);
    const bool small = c.outw <= 8;

    Params p;
    memset(&p, 0, sizeof(p));
    p.C = c.C;
    p.outw = c.outw;
    p.outh = c.outh;
    p.n = c.n;
    p.pad_left = c.pad_left;
    p.pad_top = c.pad_top;
    p.out_cpitch = c.out_cpitch;
    p.out_nstep = c.out_nstep;
    p.out_row_stride = (long long)c.outw * c.out_cpitch;
    p.w = c.w;
    p.bias = c.bias;
    p.cpad = c.cpad;
    p.act_type = c.act_type;
    p.act_p0 = c.act_p0;
    p.act_p1 = c.act_p1;

    unsigned int box[4];
    unsigned long long gdim[4] = {(unsigned long long)c.C, (unsigned long long)c.inw, (unsigned long long)c.inh, (unsigned long long)c.n};
    unsigned long long gstride[3] = {(unsigned long long)c.in_cpitch * es, (unsigned long long)c.in_cpitch * es * c.inw, (unsigned long long)c.in_nstep * es};
    CUtensorMap tm;

#define NC_DW(S_, CV_, TW_, TY_, R_)                                                            \
    do                                                                                          \
    {                                                                                           \
        box_of<T, S_, CV_, TW_, TY_, R_>(box);                                                  \
        if (tma_encode_tiled_plain(&tm, c.elemtype, 4, c.in, gdim, gstride, box) != 0) return 1; \
        return launch_dw_tma<T, S_, CV_, TW_, TY_, R_>(tm, (T*)c.out, p, stream);               \
    } while (0)

    // (tried for C = 144 = 2 x 72: CV = 9 channel vectors x 28 columns instead of 32-byte channel blocks -- 71.8 us against 69.3 us
    //  at 56 x 56 stride 1, 39.4 against 34.0 at stride 2: the stride-1 layers are bound by the consumers' dependent
    //  LDS -> convert -> FFMA2 chains at ~0.3 instructions per cycle per scheduler, not by the TMA box shape; profiles/r2)
    if (c.stride == 1)
    {
        // maps of 7k rows up to 32 columns wide (7x7, 14x14, 28x28): one column x seven rows per thread -- each staged pixel is
        // converted once for up to three outputs and the per-tile fixed cost is spread over 7 outputs per thread; wide channel
        // blocks keep 256 threads busy on narrow maps (the last block may be partial)
        if (c.outh % 7 == 0 && c.C >= 8 * VEC)
        {
            if (c.outw <= 8 && c.C >= 24 * VEC) NC_DW(1, 32, 8, 1, 7);
            // whole 14 x 14 images per tile in 64-channel blocks when those divide C and 128-channel blocks do not (576 = 9 x 64)
            if (c.outw <= 16 && c.outw > 8 && c.outh == 14 && c.C % (8 * VEC) == 0 && c.C % (16 * VEC) != 0) NC_DW(1, 8, 16, 2, 7);
            if (c.outw <= 16 && c.outw > 8 && c.C >= 12 * VEC) NC_DW(1, 16, 16, 1, 7);
            if (c.outw <= 32 && c.outw > 16) NC_DW(1, 8, 32, 1, 7);
        }
        if (small && cv == 8) NC_DW(1, 8, 8, 4, 2);
        if (cv == 8) NC_DW(1, 8, 16, 2, 4);
        if (cv == 4) NC_DW(1, 4, 16, 4, 4);
        NC_DW(1, 2, 32, 4, 4);
    }
    else
    {
        if (small && cv == 8) NC_DW(2, 8, 8, 4, 2);
        if (cv == 8) NC_DW(2, 8, 16, 2, 2);
        if (cv == 4) NC_DW(2, 4, 16, 4, 2);
        NC_DW(2, 2, 32, 4, 2);
    }
#undef NC_DW
    return 1;
}

} // namespace dwt
} // namespace ncnn_cuda
