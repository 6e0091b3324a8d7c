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
//   warp 0        : TMA producer      -- cp.async.bulk.tensor -> kStages-deep smem ring (128B-swizzled), one elected lane issues
//   warps 1, 10   : MMA issuers       -- tcgen05.mma.cta_group::1/2.kind::f16 on alternate tiles, fp32 accumulators in TMEM,
//                                        up to 8 accumulator stages (all 512 columns) so epilogues overlap the next tiles
//   warps 2..9    : epilogue          -- tcgen05.ld TMEM->registers, +bias, +residual, activation, 32-byte stores;
//                                        two warps per TMEM lane quarter, alternating 64-column chunks, no CTA-wide sync
//   full/empty mbarriers between producer and MMA, tmem_full/tmem_empty between MMA and epilogue.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace ncnn_cuda {

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (d >= 1)
struct FastDiv
{
    unsigned int d, m, s;
};

static inline FastDiv make_fastdiv(unsigned int d)
{
    FastDiv f;
    f.d = d;
    unsigned int s = 0;
    while ((1ull << s) < d) s++;
    f.s = s;
    f.m = (unsigned int)((((1ull << s) - d) << 32) / d + 1);
    return f;
}

__device__ __forceinline__ int fast_div(int n, const FastDiv& f)
{
    return (int)((__umulhi((unsigned int)n, f.m) + (unsigned int)n) >> f.s);
}

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int kEpilogueWarp0 = 2;
constexpr int kEpilogueWarps = 8; // two per TMEM lane quarter
// warp 0 TMA producer, warp 1 MMA issuer of the even tiles, warps 2..9 epilogue, warp 10 MMA issuer of the odd tiles.  An issuing
// warp spends ~750 cycles per tile on barrier waits, fences, election and descriptor bookkeeping while the tensor pipe, which only
// queues a few MMAs, drains: on tiles with little K (stems, 1x1 layers with <= 128 input channels, MobileNetV2) that overhead and
// the tile's MMAs ran back to back.  Two issuers on alternate tiles overlap one's bookkeeping with the other's MMAs.
constexpr int kMma2Warp = kEpilogueWarp0 + kEpilogueWarps;
constexpr int kNumThreads = (kMma2Warp + 1) * 32;
constexpr int kBiasSmemFloats = 4096; // layers up to this many (padded) output channels keep their bias in shared memory

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
    // A_ROWS geometry: chunks of 128 output columns per output row; rows = n * outh
    int chunks_per_row;
    // epilogue
    const float* bias; // padded to a multiple of BLOCK_N (+256), never NULL
    void* out;
    int out_cpitch;
    const void* residual; // same type/shape as out, or NULL
    int res_cpitch;
    int act_type;
    float act_p0, act_p1;
    int num_stages; // depth of the A/B ring (SmemPlan::stages_for(residual != NULL))
    int bias_smem_bytes; // shared memory reserved for the bias vector (multiple of 1024; 0 = read it from global memory)
    int v8_ok;      // output rows are 32-byte aligned: 256-bit stores
    int tma_store;  // the tile leaves through shared-memory staging + per-warp TMA stores (tmap_out) instead of per-lane sector stores
    int taps_h;     // kernel_h (A_IM2COL / A_ROWS k-block nest: filter row, filter column, channel slab)
    // A_ROWS: the zero-padded small-channel copy and its pitches (bytes)
    const unsigned char* rows_src;
    long long rows_img_bytes;
    int rows_row_bytes;
    int rows_seg_bytes; // bytes of one tile's row segment (multiple of 16)
    int rows_seg_pitch; // shared-memory pitch of the segments of one stage (multiple of 128)
    int rows_stage_bytes; // taps_h * rows_seg_pitch: one stage holds every filter row of a tile
    // A_SHIFT geometry
    int sh_bw, sh_rows, sh_colstep; // buffer row pitch (pixels), output rows per tile, output columns per column chunk
    int sh_chunks_x, sh_tiles_y;    // column chunks per row, row groups per image
    int sh_stage_bytes, sh_box_bytes;
    FastDiv div_sh_chunks_x, div_sh_tiles_y, div_sh_bw;
    // A_TILED with a SECOND A operand (a folded projection shortcut: top = act(W1 * x1 + W2 * x2 + b)): the last nk2 k-blocks of
    // the packed weight matrix multiply 64-channel slabs of a second operand map (passed in the tmap_res slot: a dual plan has no
    // fused residual, and a fifth 128-byte descriptor in the kernel's parameter block costs every instance 7-30 %, measured) -- a plain [M][C2] matrix (a2_im2col = 0) or a 1x1 strided
    // window walk over an NHWC blob through TMA im2col mode (a2_im2col = 1) -- over the SAME 128 output pixels
    int nk2, a2_im2col, a2_stride_w, a2_stride_h;
    // tile decode without integer division
    FastDiv div_n_blocks, div_opix, div_outw, div_chunks, div_outh;
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

// Monotonic "tiles observed" word of an MMA-issuing warp (see the issuers in tc_gemm_kernel): unlike a parity wait, a comparison
// against a counter cannot alias onto an earlier or later phase.
__device__ __forceinline__ void obs_publish(uint32_t addr, int seq)
{
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(addr), "r"(seq) : "memory");
}

__device__ __forceinline__ void obs_wait(uint32_t addr, int want)
{
    uint32_t spins = 0;
    while (true)
    {
        int v;
        asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v >= want) break;
        if (++spins > (1u << 26))
        {
            printf("[ncnn_cuda tc_gemm] issuer wait timed out: block %d thread %d word 0x%x want %d have %d\n", blockIdx.x, threadIdx.x, addr, want, v);
            __trap();
        }
    }
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start while
// its predecessor in the stream is still draining; everything before pdl_wait() (barrier / TMEM set-up, constant loads)
// overlaps that tail, pdl_wait() returns once the predecessor has completed and its writes are visible.
// pdl_launch_dependents() lets the successor start its own prologue as soon as SM resources free up.
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// ---- cta_group::2 (CTA pair) variants.  A pair of CTAs on the two SMs of one TPC works on one 256-row tile: each CTA stages
// its own 128 rows of A and HALF of the B tile, one tcgen05.mma.cta_group::2 issued by the leader (cluster rank 0) reads
// both halves from both shared memories (so every SM receives only half of the weights through L2 -> SM), each CTA's TMEM
// holds its own 128 accumulator rows.  The pair's TMA loads complete on the LEADER's full barrier; the MMA's commits are
// multicast to the barriers of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// the shared::cluster address of `addr` (a shared address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// arrive on an mbarrier of another CTA of the cluster (shared::cluster address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    // default (CTA-scope) semantics: a cluster-scope release compiles to MEMBAR.ALL.GPU, which waits for every store the
    // warp has in flight -- the epilogue's output stores -- before each arrive (measured: -15..-40 % on the HBM-bound layers).
    // The hand-over only has to order this warp's TMEM reads (tcgen05.wait::ld + fence::before_thread_sync) before the MMA.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__device__ __forceinline__ void tma_load_im2col_4d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c, int w, int h, int n, uint16_t off_w,
                                                       uint16_t off_h)
{
    asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
                 "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
                 : "memory");
}

// contiguous global -> shared bulk copy (16-byte granularity), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global tensor store (bulk async group); out-of-range rows / columns are clipped by the TMA unit
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// the two epilogue warps of one TMEM lane quarter (64 threads), barrier ids 2..5
__device__ __forceinline__ void quarter_bar_sync(int lane_group)
{
    asm volatile("bar.sync %0, 64;" ::"r"(lane_group + 2) : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint32_t* o)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
}

__device__ __forceinline__ void tma_store_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template<int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template<int N>
__device__ __forceinline__ void tma_store_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void epi_bar_sync()
{
    asm volatile("bar.sync 1, 128;" ::: "memory"); // the four epilogue warps only
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

__device__ __forceinline__ void tmem_alloc_cg2(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
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

__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// commit of the pair's MMAs, arriving on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
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

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_wait_ld()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tcgen05.wait::ld that also names the 32 destination registers of the load it completes as in/out operands, so that no
// use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld_pin(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// one lane writes 32 contiguous bytes = one full DRAM/L2 sector
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* o)
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]),
                 "r"(o[7])
                 : "memory");
}

__device__ __forceinline__ void st_global_v4(void* ptr, const uint32_t* o)
{
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
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

// No-swizzle ("interleaved") K-major operand: element (row m, 16-byte K chunk j) at start + (m % 8) * 16 + (m / 8) * SBO + j * LBO.
// With LBO = 16 and SBO = 128 the 8x16-byte core matrices overlap: row m begins 16 * m bytes after the start.
__device__ __forceinline__ uint64_t make_smem_desc_overlap16(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(16 >> 4) << 16;  // LBO
    d |= (uint64_t)(128 >> 4) << 32; // SBO
    d |= (uint64_t)1 << 46;          // version
    return d;                        // layout_type 0 = no swizzle
}

// Instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): fp32 accumulate, A/B K-major.
__host__ __device__ constexpr uint32_t make_idesc(int ab_format /*0 f16, 1 bf16*/, int M, int N)
{
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A-operand addressing modes
//   A_TILED  : A is a plain [M][C] matrix (1x1 stride-1 unpadded conv, InnerProduct, Gemm): 2-D TMA tiles
//   A_IM2COL : TMA im2col mode over the NHWC blob, one (tap, 64-channel slab) per k-block
//   A_ROWS   : small-channel stems (cin <= 8, e.g. 7x7 s2 / 3x3 s2 / 3x3 s1 on RGB).  The input is re-packed once per call
//              into a zero-padded [n][Hp][Wpitch][Cp] image with Cp = 4 (stride 2) or 8 (stride 1) channels per pixel, so
//              that consecutive OUTPUT columns' windows start exactly 16 bytes apart.  A k-block is one filter row.  The A
//              operand of a tile (128 consecutive output columns of one output row) is then just the raw row segment
//              (16*127 + krow*2 bytes, ONE cp.async.bulk copy): a no-swizzle K-major UMMA descriptor with LBO = 16 B and
//              SBO = 128 B addresses overlapping 8x16-byte core matrices, i.e. row m of the MMA starts 16*m bytes into the
//              segment -- the im2col expansion happens inside the tensor core's operand fetch, not in memory.
//              The layer's whole weight matrix (all filter rows) stays resident in shared memory.
//   A_SHIFT  : stride-1 k x k convolutions whose weights fit in shared memory (64 -> 64 3x3 and the like; with im2col
//              loads these are bound by the kh*kw-fold re-read of the input through L2).  A tile is R output rows x BW
//              positions of ONE image, BW = padded width (or a 128-wide column chunk); per 64-channel slab ONE 4-D tiled TMA
//              box brings the (R + kh - 1) x BW input pixels (zero OOB fill = the padding) into a 128B-swizzled [pixel][64ch]
//              buffer, and the A operand of tap (ky, kx) is that same buffer read from pixel ky * BW + kx on: the UMMA
//              descriptor's start address moves by whole 128-byte rows (the swizzle is a function of the shared-memory
//              address, so the TMA-written pattern stays consistent).  Accumulator row i is output (i / BW, i % BW); the
//              kw - 1 positions per row that wrap into the next row are computed and dropped.  kh*kw-fold less L2 traffic.
enum
{
    A_TILED = 0,
    A_IM2COL = 1,
    A_ROWS = 2,
    A_SHIFT = 3
};

template<int BLOCK_N, int BLOCK_K, int CG = 1>
struct SmemPlan
{
    static constexpr int a_bytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int b_bytes = (BLOCK_N / CG) * BLOCK_K * 2; // a CTA of a pair (CG = 2) stages half of the B tile
    static constexpr int stage_bytes = a_bytes + b_bytes; // both multiples of 1024 for the tile sizes used
    // The accumulator tile leaves the SM straight from registers (one output pixel per lane, 32-byte vector stores), so
    // the only epilogue staging is the fused residual: it arrives by TMA in EPI_N-column slots, a ring of kResSlots.
    static constexpr int EPI_N = BLOCK_N < 64 ? BLOCK_N : 64;
    static constexpr int kResSlots = 4;
    static constexpr int kMaxStages = 16;
    // accumulator stages in TMEM: as many as 512 columns hold (the MMA -> epilogue -> MMA round trip is a latency chain;
    // with few k-blocks per tile it needs more than two tiles in flight)
    static constexpr int kAccStages = (512 / BLOCK_N) > 8 ? 8 : (512 / BLOCK_N);
    static constexpr int res_slot_bytes = BLOCK_M * EPI_N * 2;
    static constexpr int barrier_bytes = 512;
    static constexpr int bias_bytes = kBiasSmemFloats * 4; // worst case; A_ROWS / A_SHIFT reserve only what the layer needs
    static constexpr int budget = 227 * 1024 - barrier_bytes - bias_bytes - 1024; // - alignment slack
    // Output staging of the TMA-store epilogue: with a fused residual the tile is written back IN PLACE into the residual slot
    // it was read from (each 16-byte unit of a slot is read and written by one warp only); without one, kStageSlots slots of
    // the same [BLOCK_M][EPI_N] shape sit where the residual ring would be (one per column half).
    static constexpr int kStageSlots = 2;
    static constexpr int aux_for(bool has_res, bool tma_store = false)
    {
        return (has_res ? kResSlots : (tma_store ? kStageSlots : 0)) * res_slot_bytes;
    }
    static constexpr int stages_for(bool has_res, bool tma_store = false)
    {
        int s = (budget - aux_for(has_res, tma_store)) / stage_bytes;
        return s > kMaxStages ? kMaxStages : s;
    }
    static constexpr int total_for(bool has_res, bool tma_store = false)
    {
        return stages_for(has_res, tma_store) * stage_bytes + aux_for(has_res, tma_store) + bias_bytes + barrier_bytes + 1024;
    }
    // A_ROWS: `aux` bytes of resident weights behind the ring
    static int stages_with_aux(int aux, int stage_sz, int bias_sz)
    {
        int s = (budget + bias_bytes - bias_sz - aux) / stage_sz;
        return s > kMaxStages ? kMaxStages : s;
    }
    static int total_with_aux(int aux, int stage_sz, int bias_sz)
    {
        return stages_with_aux(aux, stage_sz, bias_sz) * stage_sz + aux + bias_sz + barrier_bytes + 1024;
    }
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
    static __device__ __forceinline__ uint32_t pack2(float a, float b)
    {
        __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&t);
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
    static __device__ __forceinline__ uint32_t pack2(float a, float b)
    {
        __half2 t = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&t);
    }
    static constexpr int ab_format = 0;
};

// two fp32 additions in one instruction (add.rn.f32x2, FADD2 on sm_100): the epilogue's bias / residual adds
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1)
{
    asm("{\n"
        ".reg .b64 x, y;\n"
        "mov.b64 x, {%0, %1};\n"
        "mov.b64 y, {%2, %3};\n"
        "add.rn.f32x2 x, x, y;\n"
        "mov.b64 {%0, %1}, x;\n"
        "}\n"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}

// ReLU / clip on the PACKED 16-bit pair: rounding to the storage type is monotonic, so
// round(max(x, 0)) == max(round(x), 0) and round(clamp(x, lo, hi)) == clamp(round(x), round(lo), round(hi)) exactly;
// one packed min / max replaces two fp32 ones
template<typename T>
struct Packed2;
template<>
struct Packed2<__half>
{
    static __device__ __forceinline__ uint32_t relu(uint32_t v)
    {
        __half2 h = *reinterpret_cast<__half2*>(&v);
        h = __hmax2(h, __float2half2_rn(0.f));
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint32_t splat(float f)
    {
        __half2 h = __float2half2_rn(f);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint32_t clip(uint32_t v, uint32_t lo, uint32_t hi)
    {
        __half2 h = *reinterpret_cast<__half2*>(&v);
        h = __hmin2(__hmax2(h, *reinterpret_cast<__half2*>(&lo)), *reinterpret_cast<__half2*>(&hi));
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template<>
struct Packed2<__nv_bfloat16>
{
    static __device__ __forceinline__ uint32_t relu(uint32_t v)
    {
        __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
        h = __hmax2(h, __float2bfloat162_rn(0.f));
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint32_t splat(float f)
    {
        __nv_bfloat162 h = __float2bfloat162_rn(f);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint32_t clip(uint32_t v, uint32_t lo, uint32_t hi)
    {
        __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
        h = __hmin2(__hmax2(h, *reinterpret_cast<__nv_bfloat162*>(&lo)), *reinterpret_cast<__nv_bfloat162*>(&hi));
        return *reinterpret_cast<uint32_t*>(&h);
    }
};

// the rare activations (sigmoid, mish, hardswish) as an out-of-line call: keeps the unrolled epilogue small and its
// accumulators in registers
static __device__ __noinline__ float apply_activation_call(float v, int type, float p0, float p1)
{
    return apply_activation(v, type, p0, p1);
}

// ---------------------------------------------------------------- the kernel
// The residual tensor map is rank 3: (channels, columns, rows)
//   A_TILED / A_IM2COL : (C, M, 1)            tile rows m0 .. m0+127 are consecutive output pixels
//   A_ROWS             : (C, outw, n*outh)    tile rows are 128 consecutive columns of one output row
//
// CG = 2 (A_TILED / A_IM2COL): launched as clusters of two CTAs; tile = 256 output pixels x BLOCK_N, CTA `rank` of the pair
// owns m-block 2 * pair + rank (A rows, accumulator rows, epilogue, residual) and stages rows [rank * BLOCK_N / 2, +BLOCK_N / 2)
// of the B tile.  Only the leader's MMA warp runs.
template<typename T, int BLOCK_N, int BLOCK_K, int AMODE, int CG = 1>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_res,
               const __grid_constant__ CUtensorMap tmap_out, const Params p)
{
    static_assert(CG == 1 || AMODE == A_TILED || AMODE == A_IM2COL, "CTA pairs: tiled and im2col operand modes only");
    using Plan = SmemPlan<BLOCK_N, BLOCK_K, CG>;
    constexpr int EPI_N = Plan::EPI_N;
    constexpr int NCHUNK = BLOCK_N / EPI_N;    // residual slots per tile
    constexpr int SUBS = EPI_N / 32;           // 32-column TMEM loads per slot (1 or 2)
    constexpr int kResSlots = Plan::kResSlots;
    constexpr int kAccStages = Plan::kAccStages;
    constexpr uint32_t kTmemCols = kAccStages * BLOCK_N; // 512 (256 for BLOCK_N = 32): a power of two
    // swizzle of the residual staging tiles: rows of EPI_N 16-bit values = 128 / 64 bytes
    constexpr int EPI_ROW_BYTES = EPI_N * 2;
    constexpr int EPI_CHUNKS16 = EPI_ROW_BYTES / 16; // 16-byte units per row: 8 / 4

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment (pointer arithmetic keeps the shared state space visible)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kStages = p.num_stages;
    const bool has_res = p.residual != nullptr;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Plan::a_bytes;
    // behind the ring: residual slots [kResSlots][BLOCK_M][EPI_N] (has_res) or the resident stem weights (A_ROWS, whose
    // stages hold a whole tile's row segments instead of A/B k-block pairs)
    uint8_t* smem_res = smem + kStages * (AMODE == A_ROWS ? p.rows_stage_bytes : (AMODE == A_SHIFT ? p.sh_stage_bytes : Plan::stage_bytes));
    // (A_ROWS keeps the resident weights where the residual slots would be; the two never coexist)
    // A_ROWS: resident weights, then the output staging slots; A_SHIFT: resident weights (direct stores); tiled / im2col:
    // the residual ring, or the output staging slots when nothing is fused
    const int weights_bytes = AMODE == A_ROWS ? p.taps_h * Plan::b_bytes : (AMODE == A_SHIFT ? p.num_k_blocks * Plan::b_bytes : 0);
    const int aux_bytes = AMODE == A_ROWS ? weights_bytes + (p.tma_store ? Plan::kStageSlots * Plan::res_slot_bytes : 0)
                                          : (AMODE == A_SHIFT ? weights_bytes : Plan::aux_for(has_res, p.tma_store != 0));
    uint8_t* smem_stage = smem_res + weights_bytes; // output staging slots (!has_res)
    float* smem_bias = reinterpret_cast<float*>(smem_res + aux_bytes); // [kBiasSmemFloats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_bias) + p.bias_smem_bytes);
    uint64_t* full_bar = bars;                  // [16]
    uint64_t* empty_bar = bars + 16;            // [16]
    uint64_t* tmem_full_bar = bars + 32;        // [8]
    uint64_t* tmem_empty_bar = bars + 40;       // [8]
    uint64_t* res_full_bar = bars + 48;         // [kResSlots]
    uint64_t* res_empty_bar = bars + 52;        // [kResSlots]
    uint64_t* bres_bar = bars + 56;             // A_ROWS: resident weights landed
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 57);
    int* issuer_obs = reinterpret_cast<int*>(bars + 58); // [2]: last tile (per-CTA sequence number) whose operand stages each MMA issuer has seen land

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();

    const int num_n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
    int num_m_blocks;
    if (AMODE == A_ROWS)
        num_m_blocks = (int)(p.M / p.outw) * p.chunks_per_row; // rows * chunks
    else if (AMODE == A_SHIFT)
        num_m_blocks = (int)(p.M / ((long long)p.outw * p.outh)) * p.sh_tiles_y * p.sh_chunks_x; // images * row groups * column chunks
    else
        num_m_blocks = (int)((p.M + BLOCK_M - 1) / BLOCK_M);
    // CTA pairs walk (pair of m-blocks, n-block) tiles; this CTA's m-block inside the pair is its cluster rank
    const uint32_t cta_rank = CG == 2 ? (blockIdx.x & 1u) : 0u; // = %cluster_ctarank of a (2,1,1) cluster, as a value ptxas knows to be warp-uniform
    const int tile_first = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_tiles = (CG == 2 ? (num_m_blocks + 1) / 2 : num_m_blocks) * num_n_blocks;

    if (warp == 0 && lane == 0)
    {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        if (has_res || (AMODE == A_TILED && p.nk2 > 0)) prefetch_tmap(&tmap_res);
        if (p.tma_store) prefetch_tmap(&tmap_out);
    }
    if (warp == 1 && lane == 0)
    {
        for (int i = 0; i < kStages; i++)
        {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < kAccStages; i++)
        {
            mbar_init(smem_u32(&tmem_full_bar[i]), 1);
            // the leader's barrier counts the epilogue warps of both CTAs; 32-wide tiles are read by ONE half (see the epilogue)
            mbar_init(smem_u32(&tmem_empty_bar[i]), (BLOCK_N == 32 ? kEpilogueWarps / 2 : kEpilogueWarps) * CG);
        }
        mbar_init(smem_u32(bres_bar), 1);
        issuer_obs[0] = -1;
        issuer_obs[1] = -1;
        for (int i = 0; i < kResSlots; i++)
        {
            mbar_init(smem_u32(&res_full_bar[i]), 1);
            // readers of one residual slot: the four warps of the half that owns the chunk, or all eight when the tile is a
            // single 64-column chunk split between the halves
            mbar_init(smem_u32(&res_empty_bar[i]), (NCHUNK == 1 && SUBS == 2) ? 8 : 4);
        }
        fence_barrier_init();
    }
    if (warp == 2)
    {
        if (CG == 2)
            tmem_alloc_cg2(smem_u32(tmem_base_slot), kTmemCols);
        else
            tmem_alloc(smem_u32(tmem_base_slot), kTmemCols);
    }
    // the layer's (padded) bias vector: resident in shared memory when it fits
    const int bias_count = ((p.N + BLOCK_N - 1) / BLOCK_N) * BLOCK_N;
    const bool bias_in_smem = bias_count * 4 <= p.bias_smem_bytes;
    if (bias_in_smem)
        for (int i = threadIdx.x; i < bias_count; i += kNumThreads) smem_bias[i] = __ldg(p.bias + i);
    tc_fence_before();
    if (CG == 2)
        cluster_sync_all(); // the peer's barriers must be initialised before any remote arrive / TMA completion reaches them
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    // everything above touched only constants (bias, descriptors) and on-chip state; activations of the previous layer are
    // read -- and recycled pool blocks written -- only from here on
    pdl_wait();

    if (warp == 0)
    {
        // ===================== TMA producer =====================
        // The whole warp walks the loops (warp-uniform control flow: tile / k-block coordinates and shared-memory addresses
        // live in uniform registers, which is where the TMA instructions take their operands from); one elected lane issues
        // the arrive.expect_tx and the copies.  Issued from a `lane == 0` branch instead, every copy costs an ELECT /
        // R2UR.BROADCAST loop over the "divergent" lanes -- measured: the producer then never finds a free slot to wait for,
        // i.e. ITS instruction stream paces the 64-cycle MMAs of the 128-wide tiles.  No integer division in the loop: tiles
        // are decoded with multiply-shift, k-blocks by nested counters.
        int stage = 0;
        uint32_t phase = 0;
        int rslot = 0;
        uint32_t rphase = 0;
        const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const uint32_t full0_leader = CG == 2 ? mapa_shared(full0, 0) : full0;
        if (AMODE == A_ROWS || AMODE == A_SHIFT)
        {
            // the layer's weights (one n-block, every k-block): resident behind the A ring for the whole kernel
            const int nres = AMODE == A_ROWS ? p.taps_h : p.num_k_blocks;
            const uint32_t bb = smem_u32(bres_bar);
            if (elect_one())
            {
                mbar_expect_tx(bb, (uint32_t)(nres * Plan::b_bytes));
                for (int kb = 0; kb < nres; kb++) tma_load_2d(smem_u32(smem_res) + kb * Plan::b_bytes, &tmap_b, bb, kb * BLOCK_K, 0);
            }
            __syncwarp();
        }
        for (int tile = tile_first; tile < num_tiles; tile += tile_step)
        {
            const int m_grp = fast_div(tile, p.div_n_blocks);
            const int n_blk = tile - m_grp * num_n_blocks;
            const int m_blk = CG == 2 ? 2 * m_grp + (int)cta_rank : m_grp;
            const int n_coord = n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / CG); // this CTA's rows of the B tile
            int base_w = 0, base_h = 0, base_n = 0;
            int m0 = m_blk * BLOCK_M; // < 2^31: M is bounded by the host (a pair's odd m-block past the end loads zeros)
            if (AMODE == A_IM2COL)
            {
                base_n = fast_div(m0, p.div_opix);
                const int rem = m0 - base_n * (int)p.div_opix.d;
                const int oy = fast_div(rem, p.div_outw);
                const int ox = rem - oy * p.outw;
                base_w = ox * p.stride_w - p.pad_left;
                base_h = oy * p.stride_h - p.pad_top;
            }
            else if (AMODE == A_ROWS)
            {
                const int row = fast_div(m_blk, p.div_chunks); // img * outh + oy
                const int chunk = m_blk - row * p.chunks_per_row;
                base_n = fast_div(row, p.div_outh);
                base_h = (row - base_n * p.outh) * p.stride_h; // physical row of filter row 0 (top padding is materialised)
                base_w = chunk * BLOCK_M;                      // first output column of the chunk
            }
            auto advance_stage = [&]() {
                if (++stage == kStages)
                {
                    stage = 0;
                    phase ^= 1;
                }
            };
            if (AMODE == A_SHIFT)
            {
                // tile = (image, row group, column chunk); one box per 64-channel slab
                const int t1 = fast_div(m_blk, p.div_sh_chunks_x);
                const int cx = m_blk - t1 * p.sh_chunks_x;
                const int img = fast_div(t1, p.div_sh_tiles_y);
                const int ty = t1 - img * p.sh_tiles_y;
                const int w0 = cx * p.sh_colstep - p.pad_left, h0 = ty * p.sh_rows - p.pad_top;
                for (int cb = 0; cb < p.cblocks; cb++)
                {
                    mbar_wait(empty0 + stage * 8, phase ^ 1);
                    if (elect_one())
                    {
                        const uint32_t fb = full0 + stage * 8;
                        mbar_expect_tx(fb, (uint32_t)p.sh_box_bytes);
                        tma_load_4d(smem_a0 + stage * p.sh_stage_bytes, &tmap_a, fb, cb * BLOCK_K, w0, h0, img);
                    }
                    advance_stage();
                }
            }
            else if (AMODE == A_IM2COL)
            {
                int kcoord = 0;
                for (int ky = 0; ky < p.taps_h; ky++)
                    for (int kx = 0; kx < p.taps_w; kx++)
                        for (int cb = 0; cb < p.cblocks; cb++, kcoord += BLOCK_K)
                        {
                            mbar_wait(empty0 + stage * 8, phase ^ 1);
                            if (elect_one())
                            {
                                // pair: both CTAs' loads complete on the LEADER's barrier, armed by the leader for the bytes of both
                                const uint32_t fb = full0 + stage * 8;
                                if (CG == 1 || cta_rank == 0) mbar_expect_tx(fb, Plan::stage_bytes * CG);
                                if (CG == 2)
                                {
                                    tma_load_im2col_4d_cg2(smem_a0 + stage * Plan::a_bytes, &tmap_a, full0_leader + stage * 8, cb * BLOCK_K, base_w, base_h, base_n,
                                                           (uint16_t)(kx * p.dil_w), (uint16_t)(ky * p.dil_h));
                                    tma_load_2d_cg2(smem_b0 + stage * Plan::b_bytes, &tmap_b, full0_leader + stage * 8, kcoord, n_coord);
                                }
                                else
                                {
                                    tma_load_im2col_4d(smem_a0 + stage * Plan::a_bytes, &tmap_a, fb, cb * BLOCK_K, base_w, base_h, base_n, (uint16_t)(kx * p.dil_w),
                                                       (uint16_t)(ky * p.dil_h));
                                    tma_load_2d(smem_b0 + stage * Plan::b_bytes, &tmap_b, fb, kcoord, n_coord);
                                }
                            }
                            advance_stage();
                        }
            }
            else if (AMODE == A_ROWS)
            {
                // one stage = the tile's kh row segments (output column j's window starts 16*j bytes into each): one barrier
                // round trip per tile, kh contiguous bulk copies
                const unsigned char* src = p.rows_src + (long long)base_n * p.rows_img_bytes + (long long)base_h * p.rows_row_bytes + (long long)base_w * 16;
                mbar_wait(empty0 + stage * 8, phase ^ 1);
                if (elect_one())
                {
                    const uint32_t fb = full0 + stage * 8;
                    mbar_expect_tx(fb, (uint32_t)(p.taps_h * p.rows_seg_bytes));
                    uint32_t dst = smem_a0 + stage * p.rows_stage_bytes;
                    for (int ky = 0; ky < p.taps_h; ky++, src += (long long)p.dil_h * p.rows_row_bytes, dst += p.rows_seg_pitch)
                        bulk_load(dst, src, (uint32_t)p.rows_seg_bytes, fb);
                }
                advance_stage();
            }
            else
            {
                const int nk1 = p.num_k_blocks - p.nk2;
                for (int kb = 0, kcoord = 0; kb < nk1; kb++, kcoord += BLOCK_K)
                {
                    mbar_wait(empty0 + stage * 8, phase ^ 1);
                    if (elect_one())
                    {
                        const uint32_t fb = full0 + stage * 8;
                        if (CG == 1 || cta_rank == 0) mbar_expect_tx(fb, Plan::stage_bytes * CG);
                        if (CG == 2)
                        {
                            tma_load_2d_cg2(smem_a0 + stage * Plan::a_bytes, &tmap_a, full0_leader + stage * 8, kcoord, m0);
                            tma_load_2d_cg2(smem_b0 + stage * Plan::b_bytes, &tmap_b, full0_leader + stage * 8, kcoord, n_coord);
                        }
                        else
                        {
                            tma_load_2d(smem_a0 + stage * Plan::a_bytes, &tmap_a, fb, kcoord, m0);
                            tma_load_2d(smem_b0 + stage * Plan::b_bytes, &tmap_b, fb, kcoord, n_coord);
                        }
                    }
                    advance_stage();
                }
                if (p.nk2 > 0)
                {
                    // the folded shortcut's operand: same output pixels, its own channels; weights continue at column nk1 * BLOCK_K
                    int w2 = 0, h2 = 0, n2 = 0;
                    if (p.a2_im2col)
                    {
                        n2 = fast_div(m0, p.div_opix);
                        const int rem = m0 - n2 * (int)p.div_opix.d;
                        const int oy = fast_div(rem, p.div_outw);
                        w2 = (rem - oy * p.outw) * p.a2_stride_w;
                        h2 = oy * p.a2_stride_h;
                    }
                    for (int kb = 0, kc2 = 0, kcoord = nk1 * BLOCK_K; kb < p.nk2; kb++, kc2 += BLOCK_K, kcoord += BLOCK_K)
                    {
                        mbar_wait(empty0 + stage * 8, phase ^ 1);
                        if (elect_one())
                        {
                            const uint32_t fb = full0 + stage * 8;
                            if (CG == 1 || cta_rank == 0) mbar_expect_tx(fb, Plan::stage_bytes * CG);
                            if (CG == 2)
                            {
                                if (p.a2_im2col)
                                    tma_load_im2col_4d_cg2(smem_a0 + stage * Plan::a_bytes, &tmap_res, full0_leader + stage * 8, kc2, w2, h2, n2, (uint16_t)0, (uint16_t)0);
                                else
                                    tma_load_2d_cg2(smem_a0 + stage * Plan::a_bytes, &tmap_res, full0_leader + stage * 8, kc2, m0);
                                tma_load_2d_cg2(smem_b0 + stage * Plan::b_bytes, &tmap_b, full0_leader + stage * 8, kcoord, n_coord);
                            }
                            else
                            {
                                if (p.a2_im2col)
                                    tma_load_im2col_4d(smem_a0 + stage * Plan::a_bytes, &tmap_res, fb, kc2, w2, h2, n2, (uint16_t)0, (uint16_t)0);
                                else
                                    tma_load_2d(smem_a0 + stage * Plan::a_bytes, &tmap_res, fb, kc2, m0);
                                tma_load_2d(smem_b0 + stage * Plan::b_bytes, &tmap_b, fb, kcoord, n_coord);
                            }
                        }
                        advance_stage();
                    }
                }
            }
            if (has_res)
            {
                // the fused residual of this tile, EPI_N columns per ring slot; issued after the tile's operand loads so
                // that waiting for a free slot (the epilogue is at most one tile behind) never delays the MMA feed
                int c1, c2;
                if (AMODE == A_ROWS)
                {
                    c2 = fast_div(m_blk, p.div_chunks);
                    c1 = (m_blk - c2 * p.chunks_per_row) * BLOCK_M;
                }
                else
                {
                    c1 = m_blk * BLOCK_M; // < 2^31: M is bounded by the host
                    c2 = 0;
                }
                for (int cc = 0; cc < NCHUNK; cc++)
                {
                    const int c0 = n_blk * BLOCK_N + cc * EPI_N;
                    if (c0 >= p.N) break;
                    mbar_wait(smem_u32(&res_empty_bar[rslot]), rphase ^ 1);
                    if (elect_one())
                    {
                        const uint32_t rb = smem_u32(&res_full_bar[rslot]);
                        mbar_expect_tx(rb, Plan::res_slot_bytes);
                        tma_load_3d(smem_u32(smem_res + rslot * Plan::res_slot_bytes), &tmap_res, rb, c0, c1, c2);
                    }
                    if (++rslot == kResSlots)
                    {
                        rslot = 0;
                        rphase ^= 1;
                    }
                }
            }
        }
    }
    else if (warp == 1 || warp == kMma2Warp)
    {
        // ===================== MMA issuers =====================
        // The whole warp walks the loop (warp-uniform control flow keeps the descriptors in uniform registers); one elected
        // lane issues tcgen05.mma / tcgen05.commit.  Both issuing warps walk EVERY tile so that their stage / accumulator counters
        // stay in step; each issues the tiles of its parity (independent accumulator stages; a commit tracks the MMAs of its own
        // thread).
        // pair: one 256 x BLOCK_N MMA per K step, issued by the leader only (the peer's MMA warp idles)
        constexpr uint32_t idesc = make_idesc(Pack8<T>::ab_format, BLOCK_M * CG, BLOCK_N);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        const uint64_t adesc0 = make_smem_desc<BLOCK_K>(smem_a0), bdesc0 = make_smem_desc<BLOCK_K>(smem_b0);
        if (AMODE == A_ROWS || AMODE == A_SHIFT) mbar_wait(smem_u32(bres_bar), 0);
        const int my_parity = warp == 1 ? 0 : 1;
        int tile_seq = 0;
        const int stages_per_tile = AMODE == A_ROWS ? 1 : (AMODE == A_SHIFT ? p.cblocks : p.num_k_blocks);
        // Two issuers on alternate tiles when a tile takes at most half of the ring (tiles with many k-blocks do not need a second
        // issuer: their MMAs outlast the bookkeeping, warp 1 takes them all).  What makes it safe: a parity wait separates "use n of a
        // stage has landed" from "use n - 1 has landed" only for a waiter at most one phase ahead of its barrier, and TMA loads land out
        // of order, so an issuer that skips the other's tiles could reach the barrier of a stage whose previous use has not landed yet
        // and sail through on stale data (this deadlocked MobileNetV2's 144 -> 24 + residual layer at full batch); observing the other
        // issuer's stages with more parity waits aliases the other way when the observer is late (that deadlocked the default bench).
        // So the issuers tell each other through a MONOTONIC word: after the last `full` wait of its tile number q an issuer publishes
        // q (release); before the first `full` wait of tile q the other issuer waits until it reads >= q - 1 (acquire).  Then every
        // tile before q has been seen to land by one of the two (each walks its own tiles in order), in particular the previous use
        // of every stage tile q touches, so each parity wait of tile q is at most one phase ahead.  No cycle: tile q - 1 never waits
        // for anything tile q does.  The accumulator barriers are untouched by this: kAccStages is even, so a TMEM stage's previous
        // use belongs to the same issuer (and, for 32-wide tiles, to the same epilogue half).
        static_assert(kAccStages % 2 == 0, "two issuers: an accumulator stage must come back to the issuer that used it last");
        const bool dual_issue = 2 * stages_per_tile <= kStages;
        const uint32_t obs_mine = smem_u32(issuer_obs + my_parity), obs_other = smem_u32(issuer_obs + (my_parity ^ 1));
        for (int tile = tile_first; tile < (CG == 2 && cta_rank != 0 ? 0 : num_tiles); tile += tile_step)
        {
            const int seq = tile_seq++;
            if (dual_issue ? ((seq & 1) != my_parity) : (my_parity != 0))
            {
                // the other issuer's tile: step over its ring stages and its accumulator stage
                stage += stages_per_tile;
                while (stage >= kStages)
                {
                    stage -= kStages;
                    phase ^= 1;
                }
                if (++acc == kAccStages)
                {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
            tc_fence_after();
            if (dual_issue && seq > 0) obs_wait(obs_other, seq - 1);
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
            if (AMODE == A_ROWS)
            {
                // the whole tile sits in one stage: kh filter rows x (BLOCK_K / 16) MMAs, one commit.  As in the shifted-window
                // mode only the 14-bit address field of the descriptors moves: low words computed warp-uniformly outside the
                // elected branch, filter rows unrolled for the common heights (measured on the ResNet stem: ~20 instructions per
                // MMA made the issuing lane, not the tensor pipe, pace the kernel at 1.8 k cycles per tile)
                mbar_wait(full0 + stage * 8, phase);
                if (dual_issue) obs_publish(obs_mine, seq);
                tc_fence_after();
                const uint64_t a_hi = make_smem_desc_overlap16(0) & 0xFFFFFFFF00000000ull;
                const uint32_t a_flags = (uint32_t)(make_smem_desc_overlap16(0) & 0xFFFFC000ull);
                const uint64_t b_hi = make_smem_desc<BLOCK_K>(0) & 0xFFFFFFFF00000000ull;
                const uint32_t b_flags = (uint32_t)(make_smem_desc<BLOCK_K>(0) & 0xFFFFC000ull);
                const uint32_t a_lo0 = (((smem_a0 + (uint32_t)(stage * p.rows_stage_bytes)) & 0x3FFFF) >> 4) | a_flags;
                const uint32_t b_lo0 = ((smem_u32(smem_res) & 0x3FFFF) >> 4) | b_flags;
                const uint32_t a_step = (uint32_t)p.rows_seg_pitch >> 4;
                constexpr uint32_t b_step = (uint32_t)Plan::b_bytes >> 4;
                const uint32_t commit_a = empty0 + stage * 8, commit_d = smem_u32(&tmem_full_bar[acc]);
                auto mma_row = [&](int ky) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; k++)
                        umma_f16(tmem_d, a_hi | (uint64_t)(a_lo0 + (uint32_t)ky * a_step + (uint32_t)(k * 2)), b_hi | (uint64_t)(b_lo0 + (uint32_t)ky * b_step + (uint32_t)(k * 2)),
                                 idesc, (uint32_t)((ky | k) != 0));
                };
                if (p.taps_h == 3)
                {
                    if (elect_one())
                    {
#pragma unroll
                        for (int ky = 0; ky < 3; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                else if (p.taps_h == 7)
                {
                    if (elect_one())
                    {
#pragma unroll
                        for (int ky = 0; ky < 7; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                else
                {
                    if (elect_one())
                    {
                        for (int ky = 0; ky < p.taps_h; ky++) mma_row(ky);
                        umma_commit(commit_a);
                        umma_commit(commit_d);
                    }
                }
                __syncwarp();
                if (++stage == kStages)
                {
                    stage = 0;
                    phase ^= 1;
                }
            }
            else if (AMODE == A_SHIFT)
            {
                // per 64-channel slab: every tap reads the same staged pixels, shifted by ky * BW + kx rows of 128 bytes.
                // Only the 14-bit (address >> 4) field of the two descriptors changes from MMA to MMA; everything is computed
                // warp-uniformly OUTSIDE the elected branch (uniform registers), and the 3 x 3 case is fully unrolled: 36
                // MMAs per slab cost one 32-bit add each.  (Measured before: ~20 instructions per MMA, the issuing lane --
                // not the tensor pipe, the loads or the epilogue -- paced the 64 -> 64 layers.)
                // (the descriptor's base-offset field stays 0: measured on B200, the 128B swizzle is applied to the
                // absolute shared-memory address, so a start that is not 1024-byte aligned needs no correction --
                // setting the field to (start >> 7) & 7 produces wrong results)
                const uint64_t desc_hi = make_smem_desc<BLOCK_K>(0) & 0xFFFFFFFF00000000ull;
                const uint32_t desc_lo_flags = (uint32_t)(make_smem_desc<BLOCK_K>(0) & 0xFFFFC000ull); // LBO field
                const uint32_t b_lo0 = ((smem_u32(smem_res) & 0x3FFFF) >> 4) | desc_lo_flags;
                const uint32_t bw8 = (uint32_t)p.sh_bw * 8u;                  // one buffer row of pixels, in 16-byte units
                const uint32_t tap_b_step = (uint32_t)(p.cblocks * (Plan::b_bytes >> 4));
                auto issue_tap = [&](uint32_t a_lo, uint32_t b_lo, uint32_t first) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; k++)
                        umma_f16(tmem_d, desc_hi | (uint64_t)(a_lo + (uint32_t)(k * 2)), desc_hi | (uint64_t)(b_lo + (uint32_t)(k * 2)), idesc, (k == 0) ? first : 1u);
                };
                for (int cb = 0; cb < p.cblocks; cb++)
                {
                    const bool last_cb = cb == p.cblocks - 1;
                    mbar_wait(full0 + stage * 8, phase);
                    if (dual_issue && last_cb) obs_publish(obs_mine, seq);
                    tc_fence_after();
                    const uint32_t a_lo0 = (((smem_a0 + (uint32_t)(stage * p.sh_stage_bytes)) & 0x3FFFF) >> 4) | desc_lo_flags;
                    const uint32_t b_lo_cb = b_lo0 + (uint32_t)(cb * (Plan::b_bytes >> 4));
                    const uint32_t commit_bar = empty0 + stage * 8;
                    if (p.taps_h == 3 && p.taps_w == 3)
                    {
                        if (elect_one())
                        {
#pragma unroll
                            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                                for (int kx = 0; kx < 3; kx++)
                                    issue_tap(a_lo0 + (uint32_t)ky * bw8 + (uint32_t)(kx * 8), b_lo_cb + (uint32_t)(ky * 3 + kx) * tap_b_step, (uint32_t)((cb | ky | kx) != 0));
                            umma_commit(commit_bar);
                            if (last_cb) umma_commit(smem_u32(&tmem_full_bar[acc]));
                        }
                    }
                    else
                    {
                        if (elect_one())
                        {
                            uint32_t a_row = a_lo0, b_tap = b_lo_cb;
                            for (int ky = 0; ky < p.taps_h; ky++, a_row += bw8)
                                for (int kx = 0; kx < p.taps_w; kx++, b_tap += tap_b_step) issue_tap(a_row + (uint32_t)(kx * 8), b_tap, (uint32_t)((cb | ky | kx) != 0));
                            umma_commit(commit_bar);
                            if (last_cb) umma_commit(smem_u32(&tmem_full_bar[acc]));
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages)
                    {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            else
            {
                for (int kb = 0; kb < p.num_k_blocks; kb++)
                {
                    mbar_wait(full0 + stage * 8, phase);
                    if (dual_issue && kb == p.num_k_blocks - 1) obs_publish(obs_mine, seq);
                    tc_fence_after();
                    if (elect_one())
                    {
                        // descriptors of stage 0 are built once; a stage (and a 16-element K step inside the swizzle atom:
                        // 32 bytes) only moves the 14-bit (address >> 4) field, which cannot carry out below 256 KB
                        const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(stage * Plan::a_bytes) >> 4);
                        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)(stage * Plan::b_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; k++)
                        {
                            if (CG == 2)
                                umma_f16_cg2(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                            else
                                umma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                        }
                        // frees the smem slot (of both CTAs of a pair) when these MMAs retire
                        if (CG == 2)
                        {
                            umma_commit_cg2(empty0 + stage * 8);
                            if (kb == p.num_k_blocks - 1) umma_commit_cg2(smem_u32(&tmem_full_bar[acc]));
                        }
                        else
                        {
                            umma_commit(empty0 + stage * 8);
                            if (kb == p.num_k_blocks - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages)
                    {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (++acc == kAccStages)
            {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }
    else
    {
        // ===================== epilogue (warps 2..9): eight independent warps, no CTA-wide synchronisation =====================
        // Warp w may read TMEM lanes [32*(w&3), +32) = 32 output pixels, one per lane; the two warps of a lane quarter
        // ("halves") split the tile's columns: alternate 64-column chunks (BLOCK_N >= 128) or the two 32-column groups of
        // the single chunk (BLOCK_N = 64).  Per 32 accumulator columns: tcgen05.ld -> +bias (shared-memory broadcast,
        // fetched while the TMEM load is in flight) -> +residual (swizzled TMA slot) -> activation -> 16-bit pack -> two
        // 32-byte st.global.v8 per lane (each one full sector of the pixel's channel run).
        const int lane_group = warp & 3;
        const int half = (warp - kEpilogueWarp0) >> 2;
        const int row = lane_group * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t slots_seen = 0; // residual slots consumed by the CTA so far (both halves count every slot)
        const int n8 = (p.N + 7) & ~7; // the blob's padding lanes up to the next 16-byte unit may be written
        const int act = p.act_type;
        const float act_p0 = p.act_p0, act_p1 = p.act_p1;
        const int sw_row = (EPI_CHUNKS16 == 8) ? (row & 7) : ((row >> 1) & 3);
        const bool v8ok = p.v8_ok != 0;
        const int N = p.N;
        const uint32_t clip_lo2 = Packed2<T>::splat(act_p0), clip_hi2 = Packed2<T>::splat(act_p1);
        // Optional TMA-store epilogue (NCNN_B200_TC_TMASTORE=1; tiled / im2col / rows operands).  A lane that stores its own
        // pixel's 32-byte sectors touches 32 different 128-byte lines per instruction and the L1 data pipe spends ~2 cycles on
        // each (64 -> 256 layer: lsu data-pipe wavefronts at 83 % of peak).  Here the warp writes its 32 rows x EPI_N columns into
        // 128B-swizzled shared memory (conflict-free STS.128) and one lane issues a TMA tensor store of the box; rows / columns
        // outside the blob are clipped by the TMA unit; with a fused residual the rows go back into the slot they were read from.
        // MEASURED (profiles/r2/tma_store_vs_direct.txt): the data-pipe load disappears but the write-dominated layers only gain
        // 3-4 % (they sit at the HBM write rate, tools/hbm_probe.py), while the 32 KB of staging cost the compute-bound layers a
        // ring stage: ResNet-50 convs 3221 us against 3133 us with direct stores.  Kept selectable, off by default.
        const bool use_tma_store = AMODE != A_SHIFT && p.tma_store != 0;
        constexpr bool kSharedChunk = (NCHUNK == 1 && SUBS == 2); // both halves of a lane quarter fill one 64-column box
        int pending_release = -2; // -2: no store in flight; -1: a store is reading the staging slot; >= 0: ... a residual slot to release
        auto flush_store = [&]() {
            // the bulk store(s) this warp has issued must have READ their shared-memory source before it is overwritten / recycled
            if (pending_release != -2)
            {
                if (lane == 0)
                {
                    tma_store_wait_read<0>();
                    if (pending_release >= 0) mbar_arrive(smem_u32(&res_empty_bar[pending_release]));
                }
                __syncwarp();
                pending_release = -2;
            }
        };
        // the MMA warp that reuses the accumulator stage is the leader's: a pair's epilogue warps all arrive there
        const uint32_t tmem_empty_leader = CG == 2 ? mapa_shared(smem_u32(tmem_empty_bar), 0) : smem_u32(tmem_empty_bar);
        auto arrive_tmem_empty = [&](int a) {
            if (CG == 2)
                mbar_arrive_cluster(tmem_empty_leader + a * 8);
            else
                mbar_arrive(tmem_empty_leader + a * 8);
        };

        // A 32-wide tile is a single 32-column group: with the usual split one half of the warps would have nothing to read on
        // every tile, and layers with <= 32 output channels (stems, MobileNetV2 linear bottlenecks, YOLOv8 C2f halves) are bound by
        // exactly this walk.  The halves take ALTERNATE tiles instead: each accumulator stage is read -- and handed back -- by the
        // four warps of its owner only (tmem_empty counts four arrivals), the other half skips the tile without touching a barrier.
        constexpr bool kAltTiles = BLOCK_N == 32;
        int tile_seq = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step)
        {
            const int tseq = tile_seq++;
            if (kAltTiles && ((tseq & 1) != half))
            {
                slots_seen += 1u; // (one residual slot per tile at this width)
                if (++acc == kAccStages)
                {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            const int m_grp = fast_div(tile, p.div_n_blocks);
            const int n_blk = tile - m_grp * num_n_blocks;
            const int m_blk = CG == 2 ? 2 * m_grp + (int)cta_rank : m_grp;
            const int n0 = n_blk * BLOCK_N;
            long long pix;
            bool row_ok;
            if (AMODE == A_ROWS)
            {
                const int orow_idx = fast_div(m_blk, p.div_chunks);
                const int col = (m_blk - orow_idx * p.chunks_per_row) * BLOCK_M + row;
                row_ok = col < p.outw;
                pix = (long long)orow_idx * p.outw + col;
            }
            else if (AMODE == A_SHIFT)
            {
                const int t1 = fast_div(m_blk, p.div_sh_chunks_x);
                const int cx = m_blk - t1 * p.sh_chunks_x;
                const int img = fast_div(t1, p.div_sh_tiles_y);
                const int ty = t1 - img * p.sh_tiles_y;
                const int r = fast_div(row, p.div_sh_bw); // accumulator row -> (tile row, position in the buffer row)
                const int x = row - r * p.sh_bw;
                const int oy = ty * p.sh_rows + r, ox = cx * p.sh_colstep + x;
                row_ok = r < p.sh_rows && x < p.sh_colstep && oy < p.outh && ox < p.outw;
                pix = ((long long)img * p.outh + oy) * p.outw + ox;
            }
            else
            {
                pix = (long long)m_blk * BLOCK_M + row;
                row_ok = pix < p.M;
            }
            T* const orow = reinterpret_cast<T*>(p.out) + pix * p.out_cpitch + n0;
            const bool fast_store = row_ok && v8ok && (n0 + BLOCK_N <= n8);
            // TMA-store coordinates of this warp's 32 rows: (channel, column, row) as in the residual map
            int store_c1, store_c2;
            if (AMODE == A_ROWS)
            {
                store_c2 = fast_div(m_blk, p.div_chunks);
                store_c1 = (m_blk - store_c2 * p.chunks_per_row) * BLOCK_M + lane_group * 32;
            }
            else
            {
                store_c1 = m_blk * BLOCK_M + lane_group * 32;
                store_c2 = 0;
            }

            // bias of this tile's columns in shared memory: the layer's resident copy, or -- for the rare layer whose bias vector
            // does not fit -- a per-warp copy of the tile's BLOCK_N values made here, so that the group loop has ONE bias path
            int bias_base = n0;
            if (!bias_in_smem)
            {
                bias_base = (warp - kEpilogueWarp0) * BLOCK_N;
                __syncwarp();
                for (int i = lane; i < BLOCK_N; i += 32) smem_bias[bias_base + i] = __ldg(p.bias + n0 + i);
                __syncwarp();
            }
            mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * BLOCK_N);

            // 32-column groups of this tile that exist (col < N), and the ones this half owns: chunk parity selects the half
            // (group parity when the tile is a single 64-column chunk)
            const int ncols = (N - n0) < BLOCK_N ? (N - n0) : BLOCK_N;
            const int ngroups = (ncols + 31) >> 5;
            int g_begin, g_last; // first / last owned group (g_last < g_begin: none)
            if (NCHUNK == 1)
            {
                g_begin = g_last = kAltTiles ? 0 : half;
                if (g_begin >= ngroups) g_last = -1;
            }
            else
            {
                // chunk parity selects the half, and the parity flips from tile to tile: a tile with an odd number of 32-column
                // groups (96, 144, 160 ... output channels; the short last n-block of 320 or 576) would otherwise give the same
                // half the extra group every time (MobileNetV2 16 -> 96: 2 groups against 1 on every tile)
                const int nchunks = (ngroups + SUBS - 1) / SUBS;
                const int hsel = half ^ (tseq & 1);
                g_begin = hsel * SUBS;
                if (hsel >= nchunks)
                    g_last = -1;
                else
                {
                    const int last_cc = hsel + ((nchunks - 1 - hsel) & ~1);
                    const int e = last_cc * SUBS + SUBS - 1;
                    g_last = e < ngroups ? e : ngroups - 1;
                }
            }
            if (g_last < 0)
            {
                // nothing to read for this half: hand the accumulator stage back right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tmem_empty(acc);
                // single-chunk tiles count all eight warps as readers of the residual slot (TMA-store mode: the storing warp
                // arrives for both halves of its lane quarter once the box has been read)
                if (has_res && kSharedChunk && !use_tma_store && lane == 0) mbar_arrive(smem_u32(&res_empty_bar[slots_seen % kResSlots]));
            }
            // owned groups: g_begin, +1 within a chunk, then the chunk after next.  The TMEM load of the NEXT owned group is
            // issued before the current one is processed (two register buffers, ping-pong), so its latency hides behind the
            // arithmetic and the stores; the accumulator stage goes back to the MMA warp as soon as the last load has landed.
            auto next_group = [&](int g) { return (NCHUNK == 1 || (g % SUBS) != SUBS - 1) ? g + 1 : g + 1 + SUBS; };
            auto release_acc = [&]() {
                // every column group of this half sits in registers: the accumulator stage can be overwritten
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tmem_empty(acc);
            };
            auto process_group = [&](uint32_t (&r)[32], const int g) {
                const int cc = g / SUBS;         // chunk = residual slot of the tile
                const int slot_sub = g % SUBS;   // which 32-column half of the slot
                const int col0 = g * 32;
                flush_store(); // (the store issued one group ago has long read its box: no wait in practice)
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
                // + bias: LDS broadcasts (bias_base, see the tile set-up), packed fp32 adds
                {
                    const float* bs = smem_bias + bias_base + col0;
#pragma unroll
                    for (int q = 0; q < 8; q++)
                    {
                        const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * q);
                        add2(v[4 * q + 0], v[4 * q + 1], b4.x, b4.y);
                        add2(v[4 * q + 2], v[4 * q + 3], b4.z, b4.w);
                    }
                }
                if (has_res)
                {
                    const uint32_t slot_no = slots_seen + (uint32_t)cc;
                    const int rslot = (int)(slot_no % kResSlots);
                    if (slot_sub == 0 || NCHUNK == 1) mbar_wait(smem_u32(&res_full_bar[rslot]), (slot_no / kResSlots) & 1);
                    const uint8_t* rbuf = smem_res + rslot * Plan::res_slot_bytes + row * EPI_ROW_BYTES;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                    {
                        const int unit = slot_sub * 4 + u; // 16-byte unit of the slot row
                        float rv[8];
                        const uint4 ru = *reinterpret_cast<const uint4*>(rbuf + ((unit ^ sw_row) * 16));
                        Pack8<T>::unpack(ru, rv);
#pragma unroll
                        for (int j = 0; j < 8; j += 2) add2(v[u * 8 + j], v[u * 8 + j + 1], rv[j], rv[j + 1]);
                    }
                    // last read of the slot by this warp: after its final 32-column group (or the only one it owns)
                    const bool slot_done = (NCHUNK == 1) || (slot_sub == SUBS - 1) || (g + 1 >= ngroups);
                    if (slot_done && !use_tma_store)
                    {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&res_empty_bar[rslot]));
                    }
                }
                // activation: one uniform branch per group, not per element; ReLU and clip run on the packed pairs below
                if (act == 7)
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));
                }
                else if (act == 2)
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : v[j] * act_p0;
                }
                else if (act > 3)
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = apply_activation_call(v[j], act, act_p0, act_p1);
                }
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; j++) o[j] = Pack8<T>::pack2(v[2 * j], v[2 * j + 1]);
                if (act == 1)
                {
#pragma unroll
                    for (int j = 0; j < 16; j++) o[j] = Packed2<T>::relu(o[j]);
                }
                else if (act == 3)
                {
#pragma unroll
                    for (int j = 0; j < 16; j++) o[j] = Packed2<T>::clip(o[j], clip_lo2, clip_hi2);
                }
                if (use_tma_store)
                {
                    // this lane's row of the staging box: the residual slot it just read (in place), or the staging slot of its half
                    const uint32_t slot_no = slots_seen + (uint32_t)cc;
                    const int rslot = (int)(slot_no % kResSlots);
                    const uint32_t slot_base = has_res ? smem_u32(smem_res + rslot * Plan::res_slot_bytes)
                                                       : smem_u32(smem_stage + (kSharedChunk ? 0 : half) * Plan::res_slot_bytes);
                    const uint32_t srow = slot_base + (uint32_t)(row * EPI_ROW_BYTES);
#pragma unroll
                    for (int u = 0; u < 4; u++) st_shared_v4(srow + (uint32_t)((((slot_sub * 4 + u) ^ sw_row)) * 16), &o[u * 4]);
                    const bool chunk_done = (NCHUNK == 1) || (slot_sub == SUBS - 1) || (g + 1 >= ngroups);
                    if (chunk_done && !kSharedChunk)
                    {
                        fence_proxy_async(); // generic-proxy writes -> visible to the TMA unit
                        __syncwarp();
                        if (lane == 0)
                        {
                            tma_store_3d(&tmap_out, slot_base + (uint32_t)(lane_group * 32 * EPI_ROW_BYTES), n0 + cc * EPI_N, store_c1, store_c2);
                            tma_store_commit();
                        }
                        pending_release = has_res ? rslot : -1;
                    }
                }
                else if (fast_store)
                {
                    // the whole tile is inside the blob and 32-byte aligned: two unconditional sector stores
                    st_global_v8(orow + col0, &o[0]);
                    st_global_v8(orow + col0 + 16, &o[8]);
                }
                else if (row_ok)
                {
#pragma unroll
                    for (int h = 0; h < 2; h++)
                    {
                        const int c = n0 + col0 + h * 16;
                        T* dst = orow + col0 + h * 16;
                        if (v8ok && c + 16 <= n8)
                            st_global_v8(dst, &o[h * 8]);
                        else
                        {
                            if (c < n8) st_global_v4(dst, &o[h * 8]);
                            if (c + 8 < n8) st_global_v4(dst + 8, &o[h * 8 + 4]);
                        }
                    }
                }
            };
            if (g_begin <= g_last)
            {
                uint32_t ra[32], rb[32];
                int g = g_begin;
                tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 32), ra);
#pragma unroll 1
                while (true)
                {
                    tmem_wait_ld_pin(ra);
                    const int g1 = next_group(g);
                    if (g1 <= g_last)
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(g1 * 32), rb);
                    else
                        release_acc();
                    process_group(ra, g);
                    if (g1 > g_last) break;
                    tmem_wait_ld_pin(rb);
                    g = next_group(g1);
                    if (g <= g_last)
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(g * 32), ra);
                    else
                        release_acc();
                    process_group(rb, g1);
                    if (g > g_last) break;
                }
            }
            if (use_tma_store)
            {
                if (kSharedChunk)
                {
                    // both halves have staged their 32 columns of the quarter's 32 rows: one box, stored by half 0
                    const uint32_t slot_no = slots_seen;
                    const int rslot = (int)(slot_no % kResSlots);
                    const uint32_t slot_base = has_res ? smem_u32(smem_res + rslot * Plan::res_slot_bytes) : smem_u32(smem_stage);
                    fence_proxy_async();
                    quarter_bar_sync(lane_group);
                    if (half == 0 && lane == 0)
                    {
                        tma_store_3d(&tmap_out, slot_base + (uint32_t)(lane_group * 32 * EPI_ROW_BYTES), n0, store_c1, store_c2);
                        tma_store_commit();
                        tma_store_wait_read<0>();
                        if (has_res)
                        {
                            // all eight warps count as readers of a single-chunk slot: this lane arrives for both halves
                            mbar_arrive(smem_u32(&res_empty_bar[rslot]));
                            mbar_arrive(smem_u32(&res_empty_bar[rslot]));
                        }
                    }
                    quarter_bar_sync(lane_group); // the box is free again (next tile's staging / the producer's next load)
                }
                else
                    flush_store(); // do not sit on a residual slot while waiting for the next tile's accumulator
            }
            slots_seen += (uint32_t)((ngroups + SUBS - 1) / SUBS);
            if (++acc == kAccStages)
            {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    if (warp >= kEpilogueWarp0 && warp < kMma2Warp && lane == 0 && p.tma_store) tma_store_wait<0>(); // bulk stores issued by this lane have completed
    tc_fence_before();
    __syncwarp(); // (warp 0 / warp 1: the single working lane rejoins its warp before the aligned barrier)
    if (CG == 2)
        cluster_sync_all(); // neither CTA may exit (or free its TMEM) while the pair's MMAs / remote arrivals are in flight
    else
        __syncthreads();
    if (warp == 2)
    {
        tc_fence_after();
        if (CG == 2)
            tmem_dealloc_cg2(tmem_base, kTmemCols);
        else
            tmem_dealloc(tmem_base, kTmemCols);
    }
}

} // namespace tc

// ---------------------------------------------------------------- host side
// Packed weights + geometry for one layer; tensor maps for A / out / residual are encoded per forward call (they bake
// pointers and shapes), the weight maps once.
struct TcPlan
{
    int elemtype;
    int block_n, block_k;
    int cblocks, taps, num_k_blocks;
    int Kp;            // packed K length (elements)
    int outch, outch_pad;
    void* w_packed;    // device, [outch][Kp] 16-bit
    float* bias_pad;   // device, [outch_pad + 256] fp32 (zeros when no bias)
    CUtensorMap tmap_b;
    CUtensorMap tmap_b_half; // box of block_n / 2 rows: what one CTA of a cta_group::2 pair stages (valid when pair_ok)
    int pair_ok;
    CUtensorMap tmap_b_64; // box of 64 rows: narrow tiles for calls with few output pixels (valid when narrow_ok)
    int narrow_ok;
    // A_ROWS variant (small-channel stems), valid when rows_ok
    int rows_ok;
    int rows_cp, rows_wp, rows_shift; // channels per pixel of the padded copy, window width in pixels, zero taps on the left
    int rows_block_k, rows_cblocks;
    int rows_Kp;
    void* w_rows;      // device, [outch][kh * wp * cp]
    CUtensorMap tmap_b_rows;
    // folded-shortcut plan (tc_plan_create_dual): K = [k1_blocks * 64 | second operand's channels]
    int dual_k1_blocks; // 0: a plain single-operand plan
    int dual_inch2;
};

int tc_available(); // 1 when the driver exposes cuTensorMapEncode* and the device is sm_100
// cuTensorMapEncodeTiled for the bandwidth kernels (depthwise, pooling): channel-innermost blob, no swizzle, zero OOB fill.
// rank <= 5; gstride has rank-1 entries (bytes, multiples of 16).  Returns 0 / -1.
// oob_nan: out-of-bounds elements read as NaN instead of zero (max pooling: NaN never wins a NaN-ignoring maximum)
int tma_encode_tiled_plain(CUtensorMap* map, int elemtype, int rank, const void* ptr, const unsigned long long* gdim, const unsigned long long* gstride_bytes,
                           const unsigned int* box, int oob_nan = 0);
int tc_pick_block_k(int inch);
int tc_pick_block_n(int outch);
struct ncnn_cuda_conv2d_desc_fwd; // (documentation only)
// weights_tap_major: fp32 host [outch][taps][inch] (already permuted by the caller to tap-major, channel-innermost)
// kernel/stride/pad describe the layer for the A_ROWS variant (pad_left < 0: unknown until forward -> no A_ROWS)
int tc_plan_create(TcPlan* plan, int elemtype, int inch, int outch, int kernel_w, int kernel_h, int stride_w, int dil_w, int pad_left,
                   const float* weights_tap_major, const float* bias, cudaStream_t stream);
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
    // second A operand of a dual plan (NULL otherwise): [n][in2_h*in2_w][in2_cpitch], sampled with stride (in2_stride_w, in2_stride_h)
    const void* in2;
    int in2_h, in2_w, in2_ch, in2_cpitch, in2_stride_w, in2_stride_h;
    void* workspace;
    size_t workspace_size;
};

// 3x3 stride-2 max pooling fused behind a stem convolution (stem_pool.cuh): leading pads and pooled output
struct TcPoolCall
{
    int pad_left, pad_top; // rows / columns before the conv map that the first window covers (trailing ones are implied by pw / ph)
    int pw, ph;
    void* out; // [n][ph][pw][out_cpitch]
    int out_cpitch;
};
int tc_stem_pool_supported(const TcPlan* plan, const TcConvCall* call, const TcPoolCall* pool);
// the call's `out` is unused (the conv map is never written); workspace as for the A_ROWS variant.  0 ok, -1 unsupported
int tc_stem_pool_forward(const TcPlan* plan, const TcConvCall* call, const TcPoolCall* pool, cudaStream_t stream);

// returns 0 ok, -1 if the geometry cannot be expressed as a TMA descriptor (caller falls back)
int tc_conv_forward(const TcPlan* plan, const TcConvCall* call, cudaStream_t stream);
int tc_conv_supported(const TcPlan* plan, const TcConvCall* call);
// bytes of scratch the A_ROWS variant needs for this call (0 when it does not apply)
size_t tc_conv_workspace(const TcPlan* plan, const TcConvCall* call);

} // namespace ncnn_cuda
