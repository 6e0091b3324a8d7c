// stem_pool.cuh -- small-channel stem convolution (+bias, ReLU) and the 3x3 stride-2 max pooling that follows it, as ONE kernel.
//
// Replaces Convolution::forward (src/layer/convolution.cpp:113-184) + Pooling::forward max branch (src/layer/pooling.cpp:188-253)
// for the ResNet / SqueezeNet stems (conv k x k stride 2 on <= 4 channels -> <= 64 channels, then max 3x3 s2).  Unfused, the
// conv writes its 112 x 112 x 64 map (411 MB per batch of 256) and the pooling reads it back to write a quarter of it; here the
// full-resolution map never leaves the SM.
//
// The convolution is the A_ROWS operand mode of tc_gemm.cuh (the zero-padded 4-channel copy of the image, one cp.async.bulk row
// segment per filter row, overlapping no-swizzle UMMA descriptors, weights resident in shared memory); a tile is ONE conv output
// row (<= 128 columns) x 64 channels in one TMEM accumulator stage.  What differs is the work order and the epilogue:
//   * a CTA walks BANDS of pooled rows of one image (item = image x band), conv rows in ascending order, so that the three conv
//     rows of a pooling window are consecutive tiles of the same CTA (one conv row per band is computed twice: the overlap);
//   * epilogue warp (lane quarter q, column half h) owns conv columns 32q..32q+31 x channels 32h..32h+31 of EVERY tile: the
//     vertical 3-row maximum is taken in its registers -- rows (s+1, s+2) of window s arrive as two accumulator stages, row s is
//     the packed `carry` the same lane kept from the previous window -- after +bias and rounding to the storage type (rounding
//     is monotonic, so max-then-round == round-then-max and the result is bit-identical to the unfused layers);
//   * the row of vertical maxima goes to shared memory (swizzled, conflict-free STS.128), one named barrier, and the 256 epilogue
//     threads take the horizontal 3-column maximum, apply ReLU (commutes with max) and write the pooled row with coalesced 16-byte
//     stores.  Two row slots, one barrier per pooled row.
#pragma once
#include "tc_gemm.cuh"

namespace ncnn_cuda {
namespace tc {

struct StemPoolParams
{
    // conv geometry (A_ROWS)
    const unsigned char* rows_src;
    long long rows_img_bytes;
    int rows_row_bytes, rows_seg_bytes, rows_seg_pitch, rows_stage_bytes;
    int rows_copy_bytes; // one tile's operand: taps_h padded rows + the last row's segment overhang
    int taps_h, stride_h;
    int outw, outh, N; // conv output size, output channels (<= 64)
    const float* bias; // padded to 64 floats
    int relu;
    // pooling geometry: window 3, stride 2, leading pads (rows / columns outside the conv map are ignored)
    int pad_left, pad_top;
    int pw, ph;
    int band_rows, bands; // pooled rows per band, bands per image
    int num_items;        // images * bands
    FastDiv div_bands;
    void* out; // [n][ph][pw][out_cpitch]
    int out_cpitch;
    int num_stages;
    int dbg; // development ablations (NCNN_B200_STEM_DBG): 1 skip the pooling pass, 2 skip the epilogue arithmetic, 4 skip the barrier
};

template<typename T>
struct NegInf2;
template<>
struct NegInf2<__half>
{
    static constexpr uint32_t value = 0xFC00FC00u;
};
template<>
struct NegInf2<__nv_bfloat16>
{
    static constexpr uint32_t value = 0xFF80FF80u;
};

template<typename T>
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b);
template<>
__device__ __forceinline__ uint32_t hmax2_u32<__half>(uint32_t a, uint32_t b)
{
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
template<>
__device__ __forceinline__ uint32_t hmax2_u32<__nv_bfloat16>(uint32_t a, uint32_t b)
{
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

constexpr int kStemN = 64;
constexpr int kStemAccStages = 8;               // 8 x 64 = all 512 TMEM columns
constexpr int kStemRowSlotBytes = 128 * 128;    // 128 conv columns x 64 channels x 2 bytes
constexpr int kStemRowSlots = 2;
// warp 0 producer, warp 1 MMA issuer of the even conv rows, warps 2..9 epilogue, warp 10 MMA issuer of the odd conv rows.  One
// issuing warp spends ~750 cycles per tile on barrier waits, fences and descriptor bookkeeping, and the tensor pipe only queues a
// few MMAs: the 14 MMAs of a conv row (~640 cycles) and that overhead ran back to back.  Two warps on alternate tiles overlap them.
constexpr int kStemMma2Warp = kEpilogueWarp0 + kEpilogueWarps;
constexpr int kStemMmaWarps = 3; // issuing warps: warp 1 and the warps behind the epilogue warps
constexpr int kStemThreads = (kStemMma2Warp + kStemMmaWarps - 1) * 32;

template<typename T, int BLOCK_K>
__global__ void __launch_bounds__(kStemThreads, 1) stem_pool_kernel(const __grid_constant__ CUtensorMap tmap_b, const StemPoolParams p)
{
    constexpr int b_bytes = kStemN * BLOCK_K * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kStages = p.num_stages;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + kStages * p.rows_stage_bytes;           // resident weights: taps_h k-blocks of [64][BLOCK_K]
    uint8_t* smem_rows = smem_w + ((p.taps_h * b_bytes + 1023) & ~1023); // vertical-maximum row slots
    float* smem_bias = reinterpret_cast<float*>(smem_rows + kStemRowSlots * kStemRowSlotBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + kStemN);
    uint64_t* full_bar = bars;             // [16]
    uint64_t* empty_bar = bars + 16;       // [16]
    uint64_t* tmem_full_bar = bars + 32;   // [8]
    uint64_t* tmem_empty_bar = bars + 40;  // [8]
    uint64_t* w_bar = bars + 48;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 49);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) prefetch_tmap(&tmap_b);
    if (warp == 1 && lane == 0)
    {
        for (int i = 0; i < kStages; i++)
        {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < kStemAccStages; i++)
        {
            mbar_init(smem_u32(&tmem_full_bar[i]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[i]), kEpilogueWarps);
        }
        mbar_init(smem_u32(w_bar), 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_base_slot), 512);
    if (threadIdx.x < kStemN) smem_bias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    pdl_wait();

    // conv rows of item (image, band): from the first window's top row to the last window's bottom row, clipped to the map
    auto item_rows = [&](int item, int& img, int& py0, int& py1, int& r0, int& r1) {
        img = fast_div(item, p.div_bands);
        const int band = item - img * p.bands;
        py0 = band * p.band_rows;
        py1 = py0 + p.band_rows < p.ph ? py0 + p.band_rows : p.ph;
        r0 = 2 * py0 - p.pad_top;
        r1 = 2 * (py1 - 1) - p.pad_top + 2;
        if (r0 < 0) r0 = 0;
        if (r1 > p.outh - 1) r1 = p.outh - 1;
    };

    if (warp == 0)
    {
        // ===================== producer: one stage = the taps_h row segments of one conv output row =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t smem_a0 = smem_u32(smem_a);
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        {
            const uint32_t wb = smem_u32(w_bar);
            if (elect_one())
            {
                mbar_expect_tx(wb, (uint32_t)(p.taps_h * b_bytes));
                for (int kb = 0; kb < p.taps_h; kb++) tma_load_2d(smem_u32(smem_w) + kb * b_bytes, &tmap_b, wb, kb * BLOCK_K, 0);
            }
            __syncwarp();
        }
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x)
        {
            int img, py0, py1, r0, r1;
            item_rows(item, img, py0, py1, r0, r1);
            const unsigned char* src_row = p.rows_src + (long long)img * p.rows_img_bytes + (long long)r0 * p.stride_h * p.rows_row_bytes;
            for (int r = r0; r <= r1; r++, src_row += (long long)p.stride_h * p.rows_row_bytes)
            {
                mbar_wait(empty0 + stage * 8, phase ^ 1);
                if (elect_one())
                {
                    // the taps_h filter rows of a conv row are CONSECUTIVE rows of the padded copy (128-byte aligned pitch): one
                    // contiguous bulk copy per tile; the shared-memory pitch of the rows is the global one
                    const uint32_t fb = full0 + stage * 8;
                    mbar_expect_tx(fb, (uint32_t)p.rows_copy_bytes);
                    bulk_load(smem_a0 + stage * p.rows_stage_bytes, src_row, (uint32_t)p.rows_copy_bytes, fb);
                }
                if (++stage == kStages)
                {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    }
    else if (warp == 1 || warp >= kStemMma2Warp)
    {
        // ===================== MMA issuers: taps_h x (BLOCK_K / 16) MMAs per conv row, one commit =====================
        // both warps walk every tile (stage / accumulator counters stay in step); each issues the tiles of its parity
        const int my_parity = warp == 1 ? 0 : warp - kStemMma2Warp + 1;
        // several issuers only when a ring stage is reused no sooner than kStemAccStages tiles later: then a tile's own accumulator
        // wait guarantees that the stage's previous use has landed (see the dual-issue rule in tc_gemm.cuh); else warp 1 issues all
        const bool multi_issue = kStages >= kStemAccStages;
        int tile_seq = 0;
        constexpr uint32_t idesc = make_idesc(Pack8<T>::ab_format, BLOCK_M, kStemN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t smem_a0 = smem_u32(smem_a);
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const uint32_t b_base = smem_u32(smem_w);
        // Only the 14-bit (address >> 4) field of the two descriptors changes from MMA to MMA: the high words and the flag bits are
        // loop constants, the low words are computed warp-uniformly OUTSIDE the elected branch (uniform registers) and the filter
        // rows are fully unrolled for the common heights -- an MMA then costs one or two 32-bit adds instead of ~20 instructions
        // (measured on the first version: the issuing lane, not the tensor pipe, paced the kernel at 1.8 k cycles per conv row).
        const uint64_t a_hi = make_smem_desc_overlap16(0) & 0xFFFFFFFF00000000ull;
        const uint32_t a_flags = (uint32_t)(make_smem_desc_overlap16(0) & 0xFFFFC000ull);
        const uint64_t b_hi = make_smem_desc<BLOCK_K>(0) & 0xFFFFFFFF00000000ull;
        const uint32_t b_flags = (uint32_t)(make_smem_desc<BLOCK_K>(0) & 0xFFFFC000ull);
        const uint32_t b_lo0 = ((b_base & 0x3FFFF) >> 4) | b_flags;
        const uint32_t a_step = (uint32_t)p.rows_seg_pitch >> 4;
        constexpr uint32_t b_step = (uint32_t)b_bytes >> 4;
        const uint32_t stage_step = (uint32_t)p.rows_stage_bytes >> 4;
        const uint32_t a_lo_first = ((smem_a0 & 0x3FFFF) >> 4) | a_flags;
        uint32_t a_lo0 = a_lo_first;
        mbar_wait(smem_u32(w_bar), 0);
        const int taps_h = p.taps_h;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x)
        {
            int img, py0, py1, r0, r1;
            item_rows(item, img, py0, py1, r0, r1);
            for (int r = r0; r <= r1; r++)
            {
                if (multi_issue ? ((tile_seq++ % kStemMmaWarps) == my_parity) : (my_parity == 0))
                {
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kStemN);
                mbar_wait(full0 + stage * 8, phase);
                tc_fence_after();
                const uint32_t commit_a = empty0 + stage * 8, commit_d = smem_u32(&tmem_full_bar[acc]);
                auto mma_row = [&](int ky) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; k++)
                        umma_f16(tmem_d, a_hi | (uint64_t)(a_lo0 + (uint32_t)ky * a_step + (uint32_t)(k * 2)), b_hi | (uint64_t)(b_lo0 + (uint32_t)ky * b_step + (uint32_t)(k * 2)),
                                 idesc, (uint32_t)((ky | k) != 0));
                };
                if (taps_h == 7)
                {
                    if (elect_one())
                    {
#pragma unroll
                        for (int ky = 0; ky < 7; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                else if (taps_h == 3)
                {
                    if (elect_one())
                    {
#pragma unroll
                        for (int ky = 0; ky < 3; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                else
                {
                    if (elect_one())
                    {
                        for (int ky = 0; ky < taps_h; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                __syncwarp();
                }
                a_lo0 += stage_step;
                if (++stage == kStages)
                {
                    stage = 0;
                    phase ^= 1;
                    a_lo0 = a_lo_first;
                }
                if (++acc == kStemAccStages)
                {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    }
    else
    {
        // ===================== epilogue + pooling (warps 2..9, 256 threads) =====================
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const int half = (warp - kEpilogueWarp0) >> 2; // channels [32 * half, +32)
        const int col = q * 32 + lane;                 // conv column of this lane
        const int et = threadIdx.x - kEpilogueWarp0 * 32; // 0..255
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t rows0 = smem_u32(smem_rows);
        const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
        const int n8 = (p.N + 7) >> 3; // 16-byte channel units that exist
        T* const outp = reinterpret_cast<T*>(p.out);
        const int outw = p.outw, pw = p.pw, pad_left = p.pad_left, pad_top = p.pad_top, out_cpitch = p.out_cpitch;
        const bool relu = p.relu != 0;
        const int dbg = p.dbg;
        // this warp's 32 bias values stay in registers for the whole kernel
        float bias[32];
#pragma unroll
        for (int j = 0; j < 32; j++) bias[j] = smem_bias[half * 32 + j];
        // where this lane's row of vertical maxima goes: a 128-byte row per conv column, 16-byte units XOR-swizzled with the column
        uint32_t st_off[4];
#pragma unroll
        for (int u = 0; u < 4; u++) st_off[u] = (uint32_t)(col * 128 + (((half * 4 + u) ^ (col & 7)) * 16));

        auto load_row = [&](uint32_t (&raw)[32]) {
            mbar_wait(tfull0 + acc * 8, acc_phase);
            tc_fence_after();
            tmem_ld_32x32b_x32(taddr0 + (uint32_t)(acc * kStemN), raw);
        };
        auto next_acc = [&]() {
            if (++acc == kStemAccStages)
            {
                acc = 0;
                acc_phase ^= 1;
            }
        };
        // + bias, rounded to the storage type
        auto pack_row = [&](const uint32_t (&raw)[32], uint32_t (&o)[16]) {
#pragma unroll
            for (int j = 0; j < 16; j++)
            {
                float v0 = __uint_as_float(raw[2 * j]), v1 = __uint_as_float(raw[2 * j + 1]);
                add2(v0, v1, bias[2 * j], bias[2 * j + 1]);
                o[j] = Pack8<T>::pack2(v0, v1);
            }
        };

        int slot = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x)
        {
            int img, py0, py1, r0, r1;
            item_rows(item, img, py0, py1, r0, r1);
            uint32_t carry[16];
#pragma unroll
            for (int j = 0; j < 16; j++) carry[j] = NegInf2<T>::value;
            {
                // the first window's top row, when it lies inside the map: computed for the carry only
                const int s0 = 2 * py0 - pad_top;
                if (s0 >= 0 && s0 <= r1)
                {
                    uint32_t raw[32];
                    load_row(raw);
                    tmem_wait_ld_pin(raw);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + acc * 8);
                    next_acc();
                    pack_row(raw, carry);
                }
            }
            T* orow = outp + ((long long)img * p.ph + py0) * (long long)pw * out_cpitch;
            for (int py = py0; py < py1; py++, orow += (long long)pw * out_cpitch)
            {
                const int s = 2 * py - pad_top;
                // rows s+1 and s+2 of this window that exist
                const bool has1 = (s + 1 >= r0) && (s + 1 <= r1);
                const bool has2 = (s + 2 >= r0) && (s + 2 <= r1);
                uint32_t v[16];
                if (has1 && has2)
                {
                    uint32_t ra[32], rb[32];
                    load_row(ra);
                    const int a0 = acc;
                    next_acc();
                    load_row(rb);
                    const int a1 = acc;
                    next_acc();
                    tmem_wait_ld_pin(ra);
                    tmem_wait_ld_pin(rb);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                    {
                        mbar_arrive(tempty0 + a0 * 8);
                        mbar_arrive(tempty0 + a1 * 8);
                    }
                    if (!(dbg & 2))
                    {
                        uint32_t oa[16], ob[16];
                        pack_row(ra, oa);
                        pack_row(rb, ob);
#pragma unroll
                        for (int j = 0; j < 16; j++)
                        {
                            v[j] = hmax2_u32<T>(hmax2_u32<T>(carry[j], oa[j]), ob[j]);
                            carry[j] = ob[j];
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = ra[j] ^ rb[j];
                    }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = carry[j];
                    if (has1 || has2)
                    {
                        uint32_t ra[32];
                        load_row(ra);
                        tmem_wait_ld_pin(ra);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + acc * 8);
                        next_acc();
                        pack_row(ra, carry);
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = hmax2_u32<T>(v[j], carry[j]);
                    }
                }
                const uint32_t slot_base = rows0 + (uint32_t)(slot * kStemRowSlotBytes);
#pragma unroll
                for (int u = 0; u < 4; u++) st_shared_v4(slot_base + st_off[u], &v[u * 4]);
                if (!(dbg & 4)) asm volatile("bar.sync 1, 256;" ::: "memory");
                // horizontal 3-column maximum, ReLU, store: item = (pooled column, 16-byte channel unit).  A window column outside
                // the map is replaced by the nearest one inside (which is in the window too, and max is idempotent): three
                // unconditional loads in flight instead of three predicated round trips
                if (!(dbg & 1))
                    for (int it = et; it < pw * 8; it += 256)
                    {
                        const int px = it >> 3, u = it & 7;
                        if (u >= n8) continue;
                        const int c0 = 2 * px - pad_left;
                        uint4 t[3];
#pragma unroll
                        for (int k = 0; k < 3; k++)
                        {
                            int c = c0 + k;
                            c = c < 0 ? 0 : (c > outw - 1 ? outw - 1 : c);
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(t[k].x), "=r"(t[k].y), "=r"(t[k].z), "=r"(t[k].w)
                                         : "r"(slot_base + (uint32_t)(c * 128 + ((u ^ (c & 7)) * 16))));
                        }
                        uint4 m;
                        m.x = hmax2_u32<T>(hmax2_u32<T>(t[0].x, t[1].x), t[2].x);
                        m.y = hmax2_u32<T>(hmax2_u32<T>(t[0].y, t[1].y), t[2].y);
                        m.z = hmax2_u32<T>(hmax2_u32<T>(t[0].z, t[1].z), t[2].z);
                        m.w = hmax2_u32<T>(hmax2_u32<T>(t[0].w, t[1].w), t[2].w);
                        if (relu)
                        {
                            m.x = hmax2_u32<T>(m.x, 0u);
                            m.y = hmax2_u32<T>(m.y, 0u);
                            m.z = hmax2_u32<T>(m.z, 0u);
                            m.w = hmax2_u32<T>(m.w, 0u);
                        }
                        *reinterpret_cast<uint4*>(orow + (long long)px * out_cpitch + u * 8) = m;
                    }
                slot ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncwarp();
    __syncthreads();
    if (warp == 2)
    {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

} // namespace tc
} // namespace ncnn_cuda
