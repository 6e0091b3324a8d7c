// reduction.cu -- Reduction (src/layer/reduction.cpp:216-752 of the reference) and LayerNorm (layernorm.cpp, second half of the file): sum / asum / sumsq / mean / max / min /
// prod / L1 / L2 / logsum / logsumexp over any subset of the (w, h, d, c) axes of a batched blob, with or without
// keepdims.  One warp per output element: the lanes stride over the reduced index space (channel fastest when c is
// reduced, so a warp reads contiguous lanes of the channel-innermost blob), combine with shuffles, and lane 0 applies
// the reference's post step (log, sqrt with the subnormal flush of :694-706, coeff / scale for the mean).  HBM-bound
// glue for squeeze-excite style graphs; accumulation is fp32.
#include "common.cuh"

using namespace ncnn_cuda;

namespace {

struct RedGeom
{
    // logical axes in the order c, w, h, d (fastest first when walking); index 0..3
    int ext[4];
    long long in_stride[4];
    long long out_stride[4];
    int reduced[4];
    long long in_nstep, out_nstep;
    int n;
    int op;      // accumulate op: 0 sum, 1 asum, 2 sumsq, 4 max, 5 min, 6 prod, 10 sumexp
    int post;    // 0 none, 1 log, 2 sqrt
    float coeff; // final multiplier (already divided by the element count for the mean)
};

__device__ __forceinline__ float red_step(float acc, float v, int op)
{
    switch (op)
    {
    case 0: return acc + v;
    case 1: return acc + fabsf(v);
    case 2: return fmaf(v, v, acc);
    case 4: return fmaxf(acc, v);
    case 5: return fminf(acc, v);
    case 6: return acc * v;
    default: return acc + expf(v);
    }
}

__device__ __forceinline__ float red_merge(float a, float b, int op)
{
    switch (op)
    {
    case 4: return fmaxf(a, b);
    case 5: return fminf(a, b);
    case 6: return a * b;
    default: return a + b;
    }
}

template<typename T>
__global__ void __launch_bounds__(256) reduction_kernel(const T* __restrict__ in, T* __restrict__ out, RedGeom g)
{
    NC_PDL_PROLOGUE();
    long long kept = 1, red = 1;
#pragma unroll
    for (int a = 0; a < 4; a++)
    {
        if (g.reduced[a])
            red *= g.ext[a];
        else
            kept *= g.ext[a];
    }
    const long long outputs = kept * g.n;
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float v0 = g.op == 4 ? -FLT_MAX : (g.op == 5 ? FLT_MAX : (g.op == 6 ? 1.f : 0.f));
    for (long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < outputs; o += warps)
    {
        // kept coordinates -> base offsets
        long long r = o;
        long long ibase = 0, obase = 0;
#pragma unroll
        for (int a = 0; a < 4; a++)
        {
            if (g.reduced[a]) continue;
            const int x = (int)(r % g.ext[a]);
            r /= g.ext[a];
            ibase += x * g.in_stride[a];
            obase += x * g.out_stride[a];
        }
        ibase += r * g.in_nstep; // what is left of r is the sample index
        obase += r * g.out_nstep;
        float acc = v0;
        for (long long i = lane; i < red; i += 32)
        {
            long long q = i, off = ibase;
#pragma unroll
            for (int a = 0; a < 4; a++)
            {
                if (!g.reduced[a]) continue;
                const int x = (int)(q % g.ext[a]);
                q /= g.ext[a];
                off += x * g.in_stride[a];
            }
            acc = red_step(acc, to_f32(in[off]), g.op);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc = red_merge(acc, __shfl_xor_sync(0xffffffffu, acc, s), g.op);
        if (lane == 0)
        {
            if (g.post == 1) acc = logf(acc);
            if (g.post == 2) acc = sqrtf(acc < FLT_MIN ? 0.f : acc);
            out[obase] = from_f32<T>(acc * g.coeff);
        }
    }
}

// storage stride (in elements) of the logical axes c, w, h, d of a blob (include/ncnn_cuda.h layout: 1-D and 2-D blobs keep
// w in the innermost lane, 3-D/4-D blobs keep c there)
static void axis_strides(const ncnn_cuda_tensor* t, long long s[4], int ext[4])
{
    // index: 0 = c, 1 = w, 2 = h, 3 = d
    ext[0] = ext[1] = ext[2] = ext[3] = 1;
    s[0] = s[1] = s[2] = s[3] = 0;
    if (t->dims == 1)
    {
        ext[1] = t->w;
        s[1] = 1;
    }
    else if (t->dims == 2)
    {
        ext[1] = t->w;
        s[1] = 1;
        ext[2] = t->h;
        s[2] = t->cpitch;
    }
    else
    {
        ext[0] = t->c;
        s[0] = 1;
        ext[1] = t->w;
        s[1] = t->cpitch;
        ext[2] = t->h;
        s[2] = (long long)t->w * t->cpitch;
        if (t->dims == 4)
        {
            ext[3] = t->d;
            s[3] = (long long)t->h * t->w * t->cpitch;
        }
    }
}

template<typename T>
static int run_reduction(const RedGeom& g, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, long long outputs, cudaStream_t stream)
{
    NC_PDL_LAUNCH((reduction_kernel<T>), grid_for(outputs * 32, 256, 16), 256, 0, stream, (const T*)bottom->data, (T*)top->data, g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" {

int ncnn_cuda_reduction(int operation, int reduce_w, int reduce_h, int reduce_d, int reduce_c, int keepdims, float coeff, const ncnn_cuda_tensor* bottom,
                        const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(bottom && top && bottom->elemtype == top->elemtype, "reduction: blobs of one element type are required");
    NC_REQUIRE(operation >= 0 && operation <= 10, "reduction: bad operation");
    NC_REQUIRE(bottom->dims >= 1 && bottom->dims <= 4, "reduction: bad rank");
    RedGeom g;
    long long os[4];
    int oext[4];
    axis_strides(bottom, g.in_stride, g.ext);
    axis_strides(top, os, oext);
    const int dims = bottom->dims;
    // axes that do not exist in this rank are extent 1 and "reduced" (they take no part in the output index)
    g.reduced[0] = dims >= 3 ? (reduce_c ? 1 : 0) : 1;
    g.reduced[1] = (dims == 1) ? 1 : (reduce_w ? 1 : 0); // a 1-D blob always reduces w (reduction.cpp:786-789)
    g.reduced[2] = dims >= 2 ? (reduce_h ? 1 : 0) : 1;
    g.reduced[3] = dims == 4 ? (reduce_d ? 1 : 0) : 1;
    // where each kept input axis lands in the top blob (reduction.cpp:811-855)
    long long kept = 1, red = 1;
    const int order[4] = {1, 2, 3, 0}; // w, h, d, c: the order the reference pushes the surviving extents
    if (keepdims)
    {
        NC_REQUIRE(top->dims == dims, "reduction: keepdims top must keep the rank");
        for (int a = 0; a < 4; a++)
        {
            g.out_stride[a] = g.reduced[a] ? 0 : os[a];
            NC_REQUIRE(oext[a] == (g.reduced[a] ? 1 : g.ext[a]), "reduction: top shape does not match");
        }
    }
    else
    {
        int surviving[4], ns = 0;
        for (int k = 0; k < 4; k++)
        {
            const int a = order[k];
            if (!g.reduced[a]) surviving[ns++] = a;
        }
        NC_REQUIRE(top->dims == (ns == 0 ? 1 : ns), "reduction: top rank does not match the surviving axes");
        // the surviving extents fill the top's (w), (w,h), (w,h,c) or (w,h,d,c) in that order
        static const int slots[5][4] = {{1, 0, 0, 0}, {1, 0, 0, 0}, {1, 2, 0, 0}, {1, 2, 0, 0}, {1, 2, 3, 0}};
        for (int a = 0; a < 4; a++) g.out_stride[a] = 0;
        for (int k = 0; k < ns; k++)
        {
            const int slot = slots[ns][k];
            NC_REQUIRE(oext[slot] == g.ext[surviving[k]], "reduction: top shape does not match");
            g.out_stride[surviving[k]] = os[slot];
        }
        if (ns == 0) NC_REQUIRE(top->w == 1, "reduction: a full reduction writes one element per sample");
    }
    for (int a = 0; a < 4; a++)
    {
        if (g.reduced[a])
            red *= g.ext[a];
        else
            kept *= g.ext[a];
    }
    g.n = bottom->n < 1 ? 1 : bottom->n;
    NC_REQUIRE((top->n < 1 ? 1 : top->n) == g.n, "reduction: batch mismatch");
    g.in_nstep = bottom->nstep;
    g.out_nstep = top->nstep;
    // accumulate op / post step / coefficient exactly as reduction.cpp:222-273, :684-749
    switch (operation)
    {
    case 0: case 3: case 9: g.op = 0; break;
    case 1: case 7: g.op = 1; break;
    case 2: case 8: g.op = 2; break;
    case 4: g.op = 4; break;
    case 5: g.op = 5; break;
    case 6: g.op = 6; break;
    default: g.op = 10; break;
    }
    g.post = (operation == 9 || operation == 10) ? 1 : (operation == 8 ? 2 : 0);
    g.coeff = operation == 3 ? coeff / (float)red : coeff;
    const long long outputs = kept * g.n;
    if (outputs == 0) return 0;
    NC_REQUIRE(red > 0, "reduction: empty reduced extent");
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32: return run_reduction<float>(g, bottom, top, outputs, as_stream(stream));
    case NCNN_CUDA_BF16: return run_reduction<__nv_bfloat16>(g, bottom, top, outputs, as_stream(stream));
    case NCNN_CUDA_F16: return run_reduction<__half>(g, bottom, top, outputs, as_stream(stream));
    }
    return -1;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------ LayerNorm
// src/layer/layernorm.cpp:38-76 per group: mean, biased variance of (x - mean), y = (x * a + b) * gamma + beta with
// a = 1/sqrt(var + eps), b = -mean * a.  One warp per group; the group's elements are `stride` apart (1 for 1-D/2-D blobs
// whose rows are contiguous, cpitch for the pixels of one channel of a 3-D/4-D blob).  Three passes over data that stays
// in L1/L2; fp32 arithmetic.
namespace {

struct LnGeom
{
    int n, groups_per_block, blocks, size; // per sample: `blocks` x `groups_per_block` groups of `size` elements
    long long block_stride, group_stride, elem_stride, in_nstep, out_nstep;
    float eps;
};

template<typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* in, T* out, const float* __restrict__ gamma, const float* __restrict__ beta, LnGeom g)
{
    NC_PDL_PROLOGUE();
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long per_sample = (long long)g.blocks * g.groups_per_block;
    const long long total = per_sample * g.n;
    for (long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < total; o += warps)
    {
        const int b = (int)(o / per_sample);
        const long long r = o - (long long)b * per_sample;
        const long long blk = r / g.groups_per_block, grp = r - blk * g.groups_per_block;
        const long long off = blk * g.block_stride + grp * g.group_stride;
        const T* src = in + b * g.in_nstep + off;
        T* dst = out + b * g.out_nstep + off;
        float sum = 0.f;
        for (int i = lane; i < g.size; i += 32) sum += to_f32(src[i * g.elem_stride]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const float mean = sum / g.size;
        float sq = 0.f;
        for (int i = lane; i < g.size; i += 32)
        {
            const float v = to_f32(src[i * g.elem_stride]) - mean;
            sq = fmaf(v, v, sq);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
        const float a = 1.f / sqrtf(sq / g.size + g.eps);
        const float bb = -mean * a;
        for (int i = lane; i < g.size; i += 32)
        {
            float v = to_f32(src[i * g.elem_stride]) * a + bb;
            if (gamma) v = v * gamma[i] + beta[i];
            dst[i * g.elem_stride] = from_f32<T>(v);
        }
    }
}

template<typename T>
static int run_layernorm(const LnGeom& g, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const float* gamma, const float* beta, cudaStream_t stream)
{
    const long long total = (long long)g.blocks * g.groups_per_block * g.n;
    NC_PDL_LAUNCH((layernorm_kernel<T>), grid_for(total * 32, 256, 16), 256, 0, stream, (const T*)bottom->data, (T*)top->data, gamma, beta, g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" int ncnn_cuda_layernorm(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int group_size, float eps, const float* gamma_dev, const float* beta_dev, void* stream)
{
    NC_REQUIRE(bottom && top && bottom->elemtype == top->elemtype && bottom->dims == top->dims && bottom->dims >= 1 && bottom->dims <= 4, "layernorm: blob pair of one type and rank required");
    NC_REQUIRE(bottom->w == top->w && bottom->h == top->h && bottom->d == top->d && bottom->c == top->c && bottom->cpitch == top->cpitch, "layernorm: shapes differ");
    NC_REQUIRE((gamma_dev == 0) == (beta_dev == 0), "layernorm: gamma and beta come together");
    NC_REQUIRE(group_size > 0, "layernorm: empty group");
    LnGeom g;
    g.n = bottom->n < 1 ? 1 : bottom->n;
    NC_REQUIRE((top->n < 1 ? 1 : top->n) == g.n, "layernorm: batch mismatch");
    g.in_nstep = bottom->nstep;
    g.out_nstep = top->nstep;
    g.eps = eps;
    g.size = group_size;
    if (bottom->dims <= 2)
    {
        // rows are contiguous (w innermost): one group per row
        NC_REQUIRE(group_size == bottom->w, "layernorm: a 1-D / 2-D blob is normalised over w");
        g.blocks = bottom->dims == 2 ? bottom->h : 1;
        g.groups_per_block = 1;
        g.block_stride = bottom->cpitch;
        g.group_stride = 0;
        g.elem_stride = 1;
    }
    else
    {
        // the pixels of one channel are cpitch apart; a group is `group_size` consecutive pixels of a channel
        const long long pixels = (long long)bottom->w * bottom->h * (bottom->dims == 4 ? bottom->d : 1);
        NC_REQUIRE(pixels % group_size == 0 && (group_size == bottom->w || group_size == bottom->w * bottom->h || group_size == pixels), "layernorm: group is not w, w*h or w*h*d");
        g.blocks = (int)(pixels / group_size);
        g.groups_per_block = bottom->c;
        g.block_stride = (long long)group_size * bottom->cpitch;
        g.group_stride = 1;
        g.elem_stride = bottom->cpitch;
    }
    if ((long long)g.blocks * g.groups_per_block * g.n == 0) return 0;
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32: return run_layernorm<float>(g, bottom, top, gamma_dev, beta_dev, as_stream(stream));
    case NCNN_CUDA_BF16: return run_layernorm<__nv_bfloat16>(g, bottom, top, gamma_dev, beta_dev, as_stream(stream));
    case NCNN_CUDA_F16: return run_layernorm<__half>(g, bottom, top, gamma_dev, beta_dev, as_stream(stream));
    }
    return -1;
}
