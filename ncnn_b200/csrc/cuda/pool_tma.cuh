// pool_tma.cuh -- the bandwidth path of max Pooling (src/layer/pooling.cpp:188-253 of the reference): KxK windows
// (K = 2, 3, 5), stride 1 or 2, on channel-innermost blobs.  Same machinery as dwconv_tma.cuh: a persistent CTA owns one
// channel block, a producer lane keeps a ring of input tiles (+ window overhang) in shared memory with 4-D tiled TMA
// loads, 256 consumer threads each own one output column of the tile for a 16-byte channel vector and R output rows.
//
// The reference pads with -FLT_MAX (pooling.cpp:367-383, copy_make_border) so that out-of-image taps never win; here the
// tensor map's out-of-bounds fill is NaN and the maximum is taken with the NaN-ignoring packed min/max instructions
// (HMNMX2 for fp16/bf16 -- the maximum of stored 16-bit values is exact, no conversion at all; FMNMX for fp32), which
// has the same effect without a padded copy.  A window that lies entirely in padding yields NaN here and -FLT_MAX in the
// reference; both mark "no input" and the tests treat them alike.
#pragma once
#include "dwconv_tma.cuh"

namespace ncnn_cuda {
namespace plt {

constexpr int kConsumers = dwt::kConsumers;
constexpr int kThreads = dwt::kThreads;

struct Params
{
    int C, outw, outh, n;
    int tiles_x, tiles_y;
    dwt::FastDiv div_image, div_tiles_x;
    int n_spatial;
    int cblocks;
    int pad_left, pad_top;
    int out_cpitch;
    long long out_nstep;
    long long out_row_stride;
};

template<typename T>
struct MaxVec;
template<>
struct MaxVec<__half>
{
    static __device__ __forceinline__ uint4 init()
    {
        return make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u); // -inf
    }
    static __device__ __forceinline__ void acc(uint4& m, const uint4& x)
    {
        __half2* a = reinterpret_cast<__half2*>(&m);
        const __half2* b = reinterpret_cast<const __half2*>(&x);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = __hmax2(a[i], b[i]);
    }
};
template<>
struct MaxVec<__nv_bfloat16>
{
    static __device__ __forceinline__ uint4 init()
    {
        return make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u); // -inf
    }
    static __device__ __forceinline__ void acc(uint4& m, const uint4& x)
    {
        __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&m);
        const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = __hmax2(a[i], b[i]);
    }
};
template<>
struct MaxVec<float>
{
    static __device__ __forceinline__ uint4 init()
    {
        const uint32_t v = __float_as_uint(-FLT_MAX);
        return make_uint4(v, v, v, v);
    }
    static __device__ __forceinline__ void acc(uint4& m, const uint4& x)
    {
        m.x = __float_as_uint(fmaxf(__uint_as_float(m.x), __uint_as_float(x.x)));
        m.y = __float_as_uint(fmaxf(__uint_as_float(m.y), __uint_as_float(x.y)));
        m.z = __float_as_uint(fmaxf(__uint_as_float(m.z), __uint_as_float(x.z)));
        m.w = __float_as_uint(fmaxf(__uint_as_float(m.w), __uint_as_float(x.w)));
    }
};

template<typename T, int K, int S, int CV, int TW, int TY, int R>
struct Cfg
{
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int CB = CV * VEC;
    static constexpr int TH = TY * R;
    static constexpr int IW = (TW - 1) * S + K;
    static constexpr int IH = (TH - 1) * S + K;
    static constexpr int tile_bytes = IW * IH * CV * 16;
    static constexpr int stage_bytes = (tile_bytes + 127) / 128 * 128;
    static constexpr int cta_budget = 111 * 1024;
    static constexpr int fixed_bytes = 128 + 128;
    static constexpr int stages_fit = (cta_budget - fixed_bytes) / stage_bytes;
    static constexpr int kStages = stages_fit > 4 ? 4 : stages_fit;
    static constexpr int smem_bytes = kStages * stage_bytes + fixed_bytes;
    static_assert(CV * TW * TY == kConsumers, "one consumer thread per (channel vector, column, thread row)");
    static_assert(kStages >= 2, "tile too large for a 2-stage ring");
};

template<typename T, int K, int S, int CV, int TW, int TY, int R>
__global__ void __launch_bounds__(kThreads, 2) maxpool_tma_kernel(const __grid_constant__ CUtensorMap tmap_in, T* __restrict__ out, const Params p)
{
    using C = Cfg<T, K, S, CV, TW, TY, R>;
    constexpr int VEC = C::VEC;
    constexpr int NROWS = (R - 1) * S + K;
    constexpr int kStages = C::kStages;

    extern __shared__ uint8_t pool_smem_raw[];
    uint8_t* smem = pool_smem_raw + ((128u - (tc::smem_u32(pool_smem_raw) & 127u)) & 127u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::stage_bytes);
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
    __syncthreads();
    tc::pdl_wait();

    // one channel block per CTA, striding over the spatial tiles (see dwconv_tma.cuh)
    const int cb = blockIdx.x % p.cblocks;
    const int sp_first = blockIdx.x / p.cblocks;
    const int sp_step = gridDim.x / p.cblocks;

    if (tid >= kConsumers)
    {
        if (tid == kConsumers)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int sp = sp_first; sp < p.n_spatial; sp += sp_step)
            {
                const int b = dwt::fast_div(sp, p.div_image);
                const int t2 = sp - b * (int)p.div_image.d;
                const int tyi = dwt::fast_div(t2, p.div_tiles_x);
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

    const int cv = tid % CV;
    const int tx = (tid / CV) % TW;
    const int ty = tid / (CV * TW);
    const int lane = tid & 31;
    int stage = 0;
    uint32_t phase = 0;
    const int thread_off = ((ty * R * S) * C::IW + tx * S) * (CV * 16) + cv * 16;
    const int c0 = cb * C::CB + cv * VEC;

    for (int sp = sp_first; sp < p.n_spatial; sp += sp_step)
    {
        const int b = dwt::fast_div(sp, p.div_image);
        const int t2 = sp - b * (int)p.div_image.d;
        const int tyi = dwt::fast_div(t2, p.div_tiles_x);
        const int txi = t2 - tyi * p.tiles_x;

        uint4 m[R];
#pragma unroll
        for (int r = 0; r < R; r++) m[r] = MaxVec<T>::init();

        tc::mbar_wait(tc::smem_u32(&full_bar[stage]), phase);
        const uint8_t* base = smem + stage * C::stage_bytes + thread_off;
#pragma unroll
        for (int ii = 0; ii < NROWS; ii++)
        {
            // horizontal maximum of this input row's K taps, then folded into every output row whose window holds the row
            uint4 h = *reinterpret_cast<const uint4*>(base + (ii * C::IW) * (CV * 16));
#pragma unroll
            for (int kx = 1; kx < K; kx++) MaxVec<T>::acc(h, *reinterpret_cast<const uint4*>(base + (ii * C::IW + kx) * (CV * 16)));
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                const int ky = ii - r * S;
                if (ky >= 0 && ky < K) MaxVec<T>::acc(m[r], h);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty_bar[stage]));
        if (++stage == kStages)
        {
            stage = 0;
            phase ^= 1;
        }

        const int ox = txi * TW + tx;
        const int oy0 = tyi * C::TH + ty * R;
        if (ox < p.outw)
        {
            T* op = out + ((long long)b * p.out_nstep + ((long long)oy0 * p.outw + ox) * p.out_cpitch + c0);
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                if (oy0 + r < p.outh) *reinterpret_cast<uint4*>(op) = m[r];
                op += p.out_row_stride;
            }
        }
    }
}

template<typename T, int K, int S, int CV, int TW, int TY, int R>
static int launch_pool_tma(const void* in, int elemtype, const unsigned long long* gdim, const unsigned long long* gstride, T* out, Params& p, cudaStream_t stream)
{
    using C = Cfg<T, K, S, CV, TW, TY, R>;
    auto kern = maxpool_tma_kernel<T, K, S, CV, TW, TY, R>;
    static bool attr_set = false;
    if (!attr_set)
    {
        NC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
        attr_set = true;
    }
    unsigned int box[4] = {(unsigned int)C::CB, (unsigned int)C::IW, (unsigned int)C::IH, 1u};
    CUtensorMap tm;
    if (tma_encode_tiled_plain(&tm, elemtype, 4, in, gdim, gstride, box, /*oob_nan=*/1) != 0) return 1;
    p.tiles_x = (p.outw + TW - 1) / TW;
    p.tiles_y = (p.outh + C::TH - 1) / C::TH;
    const long long n_spatial = (long long)p.n * p.tiles_x * p.tiles_y;
    if (n_spatial > 0x3fffffffLL) return 1;
    p.n_spatial = (int)n_spatial;
    p.cblocks = p.C / C::CB;
    p.div_image = dwt::make_fastdiv((unsigned int)(p.tiles_x * p.tiles_y));
    p.div_tiles_x = dwt::make_fastdiv((unsigned int)p.tiles_x);
    long long groups = (2LL * sm_count()) / p.cblocks;
    if (groups < 1) groups = 1;
    if (groups > n_spatial) groups = n_spatial;
    NC_CHECK(launch_pdl(kern, dim3((unsigned int)(groups * p.cblocks)), dim3(kThreads), (size_t)C::smem_bytes, stream, tm, out, p));
    NC_LAUNCH_CHECK();
    return 0;
}

struct Call
{
    const void* in;
    void* out;
    int elemtype;
    int C, inw, inh, outw, outh, n;
    int kernel, stride;
    int pad_left, pad_top;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
};

// 0 launched, 1 not applicable (caller uses the generic kernel), < 0 error
template<typename T>
static int forward(const Call& c, cudaStream_t stream)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int es = (int)sizeof(T);
    if (!tc_available()) return 1;
    if (c.stride != 1 && c.stride != 2) return 1;
    if (c.kernel != 2 && c.kernel != 3 && c.kernel != 5) return 1;
    if (c.C % (2 * VEC) != 0) return 1;
    if (((size_t)c.in_cpitch * es) % 16 || ((size_t)c.in_nstep * es) % 16 || ((uintptr_t)c.in & 15)) return 1;
    if (((size_t)c.out_cpitch * es) % 16 || ((size_t)c.out_nstep * es) % 16 || ((uintptr_t)c.out & 15)) return 1;
    if (c.pad_left < 0 || c.pad_top < 0 || c.pad_left > 64 || c.pad_top > 64) return 1;
    const int cv = (c.C % (8 * VEC) == 0) ? 8 : ((c.C % (4 * VEC) == 0) ? 4 : 2);

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
    unsigned long long gdim[4] = {(unsigned long long)c.C, (unsigned long long)c.inw, (unsigned long long)c.inh, (unsigned long long)c.n};
    unsigned long long gstride[3] = {(unsigned long long)c.in_cpitch * es, (unsigned long long)c.in_cpitch * es * c.inw, (unsigned long long)c.in_nstep * es};

#define NC_PL(K_, S_, CV_, TW_, TY_, R_) return launch_pool_tma<T, K_, S_, CV_, TW_, TY_, R_>(c.in, c.elemtype, gdim, gstride, (T*)c.out, p, stream)
#define NC_PL_CV(K_, S_, R_)                       \
    do                                             \
    {                                              \
        if (cv == 8) NC_PL(K_, S_, 8, 16, 2, R_);  \
        if (cv == 4) NC_PL(K_, S_, 4, 16, 4, R_);  \
        NC_PL(K_, S_, 2, 32, 4, R_);               \
    } while (0)
    if (c.kernel == 2 && c.stride == 2) NC_PL_CV(2, 2, 2);
    if (c.kernel == 3 && c.stride == 2) NC_PL_CV(3, 2, 2);
    if (c.kernel == 3 && c.stride == 1) NC_PL_CV(3, 1, 4);
    if (c.kernel == 5 && c.stride == 1) NC_PL_CV(5, 1, 4);
    if (c.kernel == 2 && c.stride == 1) NC_PL_CV(2, 1, 4);
#undef NC_PL_CV
#undef NC_PL
    return 1;
}

} // namespace plt
} // namespace ncnn_cuda
