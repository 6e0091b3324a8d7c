// tc_gemm.cuh -- Blackwell (sm_100a) tensor-core implicit-GEMM for Convolution / InnerProduct / Gemm.
//
//   D[m][oc] = act( sum_k A[m][k] * W[oc][k] + bias[oc] (+ residual[m][oc]) )
//
// A = activations, channel-innermost 16-bit blob [n][P][cpitch]; rows of A are output pixels.
//   mode TILED : 1x1 stride-1 unpadded conv / InnerProduct / Gemm: A is a plain [M][C] matrix, 2-D TMA tiles.
//   mode IM2COL: any kernel/stride/dilation/zero padding: TMA *im2col mode* gathers, for filter tap
//                (ky,kx) and a 64-channel slab, the BLOCK_M consecutive output pixels' input pixels
//                straight from the NHWC blob (no im2col buffer in HBM, halo/padding = TMA OOB zero fill).
// W = weights re-packed once at create_pipeline time to [outch][taps * cblocks * BLOCK_K] K-major.
//
// Structure (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0 lane 0 : TMA producer      -- cp.async.bulk.tensor -> kStages-deep smem ring (128B-swizzled)
//   warp 1 lane 0 : MMA issuer        -- tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM,
//                                        2 accumulator stages so the epilogue of tile i overlaps tile i+1
//   warps 2..5    : epilogue          -- tcgen05.ld TMEM->registers, +bias, +residual, activation, 16-byte stores
//   full/empty mbarriers between producer and MMA, tmem_full/tmem_empty between MMA and epilogue.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace ncnn_cuda {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int kNumThreads = 192;
constexpr int kEpilogueWarp0 = 2;

struct Params
{
    long long M;   // total output pixels (n * outh * outw)
    int N;         // outch
    int num_k_blocks;
    int cblocks;   // channel slabs per filter tap
    int taps_w;    // kernel_w (taps = kernel_w * kernel_h)
    // im2col geometry
    int outw, outh;
    int stride_w, stride_h, dil_w, dil_h, pad_left, pad_top;
    // epilogue
    const float* bias; // padded to a multiple of BLOCK_N, never NULL
    void* out;
    int out_cpitch;
    const void* residual; // same type/shape as out, or NULL
    int res_cpitch;
    int act_type;
    float act_p0, act_p1;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Spin on try_wait; a bounded spin turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 26))
        {
            printf("[ncnn_cuda tc_gemm] mbarrier wait timed out: block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)map),
                 "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n, uint16_t off_w, uint16_t off_h)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_wait_ld()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in smem, rows of BLOCK_K 16-bit elements = SWIZZLE bytes, 8-row groups SBO apart.
// Bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) (2 = 128B, 4 = 64B, 6 = 32B swizzle).
template<int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    constexpr uint32_t swizzle_bytes = BLOCK_K * 2;
    constexpr uint64_t layout_type = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * swizzle_bytes) >> 4;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= sbo << 32;
    d |= (uint64_t)1 << 46;
    d |= layout_type << 61;
    return d;
}

// Instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): fp32 accumulate, A/B K-major.
__host__ __device__ constexpr uint32_t make_idesc(int ab_format /*0 f16, 1 bf16*/, int M, int N)
{
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template<int BLOCK_N, int BLOCK_K>
struct SmemPlan
{
    static constexpr int a_bytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int b_bytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int stage_bytes = a_bytes + b_bytes; // both multiples of 1024 for the tile sizes used
    static constexpr int max_bytes = 200 * 1024;
    static constexpr int stages_raw = max_bytes / stage_bytes;
    static constexpr int kStages = stages_raw > 8 ? 8 : stages_raw;
    static constexpr int barrier_bytes = 256;
    static constexpr int total = kStages * stage_bytes + barrier_bytes + 1024; // + alignment slack
};

template<typename T>
struct Pack8;
template<>
struct Pack8<__nv_bfloat16>
{
    static __device__ __forceinline__ uint4 pack(const float (&v)[8])
    {
        uint4 u;
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c);
        u.w = *reinterpret_cast<uint32_t*>(&d);
        return u;
    }
    static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[8])
    {
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            float2 f = __bfloat1622float2(p[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static constexpr int ab_format = 1;
};
template<>
struct Pack8<__half>
{
    static __device__ __forceinline__ uint4 pack(const float (&v)[8])
    {
        uint4 u;
        __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c);
        u.w = *reinterpret_cast<uint32_t*>(&d);
        return u;
    }
    static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[8])
    {
        const __half2* p = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            float2 f = __half22float2(p[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static constexpr int ab_format = 0;
};

// ---------------------------------------------------------------- the kernel
template<typename T, int BLOCK_N, int BLOCK_K, bool IM2COL>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p)
{
    using Plan = SmemPlan<BLOCK_N, BLOCK_K>;
    constexpr int kStages = Plan::kStages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N; // power of two for BLOCK_N in {16..256}

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Plan::a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Plan::stage_bytes);
    uint64_t* full_bar = bars;                  // [kStages]
    uint64_t* empty_bar = bars + kStages;       // [kStages]
    uint64_t* tmem_full_bar = bars + 2 * kStages;  // [2]
    uint64_t* tmem_empty_bar = bars + 2 * kStages + 2; // [2]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m_blocks = (int)((p.M + BLOCK_M - 1) / BLOCK_M);
    const int num_n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m_blocks * num_n_blocks;

    if (warp == 0 && lane == 0)
    {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0)
    {
        for (int i = 0; i < kStages; i++)
        {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; i++)
        {
            mbar_init(smem_u32(&tmem_full_bar[i]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[i]), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2)
    {
        tmem_alloc(smem_u32(tmem_base_slot), kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0)
    {
        if (lane == 0)
        {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                const int n_blk = tile % num_n_blocks;
                const int m_blk = tile / num_n_blocks;
                const long long m0 = (long long)m_blk * BLOCK_M;
                int base_w = 0, base_h = 0, base_n = 0;
                if (IM2COL)
                {
                    const int opix = p.outw * p.outh;
                    base_n = (int)(m0 / opix);
                    int rem = (int)(m0 - (long long)base_n * opix);
                    int oy = rem / p.outw;
                    int ox = rem - oy * p.outw;
                    base_w = ox * p.stride_w - p.pad_left;
                    base_h = oy * p.stride_h - p.pad_top;
                }
                for (int kb = 0; kb < p.num_k_blocks; kb++)
                {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, Plan::stage_bytes);
                    const int tap = kb / p.cblocks;
                    const int cb = kb - tap * p.cblocks;
                    if (IM2COL)
                    {
                        const int ky = tap / p.taps_w;
                        const int kx = tap - ky * p.taps_w;
                        tma_load_im2col_4d(smem_u32(smem_a + stage * Plan::a_bytes), &tmap_a, fb, cb * BLOCK_K, base_w, base_h, base_n,
                                           (uint16_t)(kx * p.dil_w), (uint16_t)(ky * p.dil_h));
                    }
                    else
                    {
                        tma_load_2d(smem_u32(smem_a + stage * Plan::a_bytes), &tmap_a, fb, kb * BLOCK_K, (int)m0);
                    }
                    tma_load_2d(smem_u32(smem_b + stage * Plan::b_bytes), &tmap_b, fb, kb * BLOCK_K, n_blk * BLOCK_N);
                    if (++stage == kStages)
                    {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        if (lane == 0)
        {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc(Pack8<T>::ab_format, BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < p.num_k_blocks; kb++)
                {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc<BLOCK_K>(smem_u32(smem_a + stage * Plan::a_bytes));
                    const uint64_t bdesc = make_smem_desc<BLOCK_K>(smem_u32(smem_b + stage * Plan::b_bytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; k++)
                    {
                        // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr>>4) field
                        umma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(smem_u32(&empty_bar[stage])); // frees the smem slot when these MMAs retire
                    if (kb == p.num_k_blocks - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
                    if (++stage == kStages)
                    {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    }
    else
    {
        // ===================== epilogue (warps 2..5) =====================
        const int lane_group = warp & 3; // TMEM lanes [32*lane_group, +32) are the ones this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        T* out = reinterpret_cast<T*>(p.out);
        const T* res = reinterpret_cast<const T*>(p.residual);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            const int n_blk = tile % num_n_blocks;
            const int m_blk = tile / num_n_blocks;
            const long long m = (long long)m_blk * BLOCK_M + lane_group * 32 + lane;
            const bool row_ok = m < p.M;
            const int n0 = n_blk * BLOCK_N;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            T* orow = out + m * p.out_cpitch;
            const T* rrow = res ? res + m * p.res_cpitch : nullptr;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32)
            {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + (uint32_t)c, r);
                tmem_wait_ld();
                if (row_ok)
                {
#pragma unroll
                    for (int g8 = 0; g8 < 4; g8++)
                    {
                        const int col = n0 + c + g8 * 8;
                        if (col < p.out_cpitch)
                        {
                            float v[8];
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                            v[0] = __uint_as_float(r[g8 * 8 + 0]) + b0.x;
                            v[1] = __uint_as_float(r[g8 * 8 + 1]) + b0.y;
                            v[2] = __uint_as_float(r[g8 * 8 + 2]) + b0.z;
                            v[3] = __uint_as_float(r[g8 * 8 + 3]) + b0.w;
                            v[4] = __uint_as_float(r[g8 * 8 + 4]) + b1.x;
                            v[5] = __uint_as_float(r[g8 * 8 + 5]) + b1.y;
                            v[6] = __uint_as_float(r[g8 * 8 + 6]) + b1.z;
                            v[7] = __uint_as_float(r[g8 * 8 + 7]) + b1.w;
                            if (rrow)
                            {
                                float rv[8];
                                const uint4 ru = *reinterpret_cast<const uint4*>(rrow + col);
                                Pack8<T>::unpack(ru, rv);
#pragma unroll
                                for (int j = 0; j < 8; j++) v[j] += rv[j];
                            }
                            if (p.act_type != 0)
                            {
#pragma unroll
                                for (int j = 0; j < 8; j++) v[j] = apply_activation(v[j], p.act_type, p.act_p0, p.act_p1);
                            }
                            *reinterpret_cast<uint4*>(orow + col) = Pack8<T>::pack(v);
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2)
    {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

} // namespace tc

// ---------------------------------------------------------------- host side
// Packed weights + geometry for one layer; tensor maps for A are encoded per forward call (they bake
// the activation pointer and the input shape), the weight map once.
struct TcPlan
{
    int elemtype;
    int block_n, block_k;
    int cblocks, taps, num_k_blocks;
    int Kp;            // packed K length (elements)
    int outch, outch_pad;
    void* w_packed;    // device, [outch_pad][Kp] 16-bit
    float* bias_pad;   // device, [outch_pad + 256] fp32 (zeros when no bias)
    CUtensorMap tmap_b;
};

int tc_available(); // 1 when the driver exposes cuTensorMapEncode* and the device is sm_100
int tc_pick_block_k(int inch);
int tc_pick_block_n(int outch);
// weights_k_major: fp32 host [outch][taps][inch] (already permuted by the caller to tap-major, channel-innermost)
int tc_plan_create(TcPlan* plan, int elemtype, int inch, int outch, int taps, const float* weights_tap_major, const float* bias, cudaStream_t stream);
void tc_plan_destroy(TcPlan* plan);

struct TcConvCall
{
    const void* in;    // [n][inh*inw][in_cpitch]
    int n, inh, inw, inch, in_cpitch;
    int outh, outw;
    int kernel_w, kernel_h, stride_w, stride_h, dil_w, dil_h, pad_left, pad_top, pad_right, pad_bottom;
    void* out;
    int out_cpitch;
    const void* residual;
    int res_cpitch;
    int act_type;
    float act_p0, act_p1;
    int tiled; // 1: A is a plain [M][inch] matrix (1x1 s1 p0 / linear)
};

// returns 0 ok, -1 if the geometry cannot be expressed as a TMA im2col descriptor (caller falls back)
int tc_conv_forward(const TcPlan* plan, const TcConvCall* call, cudaStream_t stream);
int tc_conv_supported(const TcPlan* plan, const TcConvCall* call);

} // namespace ncnn_cuda
