// glue.cu -- the non-GEMM operators the named models use: ReLU/Sigmoid/Swish/... (unary), Eltwise, BinaryOp,
// Concat/Slice copies, Interp, Softmax, Padding.  All are single memory sweeps (HBM-bound): 16-byte vector
// accesses, grid-stride loops sized to a few waves of 148 SMs, fp32 math on 16-bit storage.
// Reference bodies: src/layer/{relu,sigmoid,swish,eltwise,binaryop,concat,slice,interp,softmax,padding}.cpp
#include "common.cuh"

using namespace ncnn_cuda;

namespace {

// ---------------------------------------------------------------- helpers
struct Flat
{
    long long rows;   // n * P  (only valid when dense)
    int C, cpitch;
};

// "dense" for the flat (whole-buffer) kernels: samples back to back AND no foreign data between the pixels -- a channel-range
// view into a wider blob (Slice views, producers writing into their Concat's buffer) has cpitch far beyond its own channels and
// the lanes in between belong to its neighbours
static bool is_dense(const ncnn_cuda_tensor* t)
{
    TView v = make_view(t);
    if ((long long)(v.cpitch - v.C) * (long long)elem_size(t->elemtype) >= 16) return false;
    return v.n == 1 || v.nstep == (long long)v.P * v.cpitch;
}

// pixel-dense: element (b, p, q) at (b * P + p) * cpitch + q -- true for compact blobs and for channel-range views alike
static bool is_pixel_dense(const ncnn_cuda_tensor* t)
{
    TView v = make_view(t);
    return v.n == 1 || v.nstep == (long long)v.P * v.cpitch;
}

// the pitched vector kernels below: same logical shape, C a whole number of 16-byte vectors, every operand pixel-dense with a
// 16-byte aligned base and pitch
static bool pitched_vec_ok(const ncnn_cuda_tensor* t, int vec)
{
    TView v = make_view(t);
    return v.C > 0 && v.C % vec == 0 && t->cpitch % vec == 0 && (((uintptr_t)t->data) & 15) == 0 && is_pixel_dense(t);
}

// total addressable elements of a dense blob including pad lanes
static long long flat_count(const ncnn_cuda_tensor* t)
{
    TView v = make_view(t);
    return (long long)v.n * v.P * v.cpitch;
}

static bool same_layout(const ncnn_cuda_tensor* a, const ncnn_cuda_tensor* b)
{
    return same_shape(a, b) && a->cpitch == b->cpitch && a->nstep == b->nstep && a->elemtype == b->elemtype;
}

static bool aligned16(const void* p)
{
    return ((uintptr_t)p & 15) == 0;
}

__device__ __forceinline__ float unary_apply(int op, float v, float p0, float p1)
{
    switch (op)
    {
    case NCNN_CUDA_UNARY_RELU: return p0 == 0.f ? fmaxf(v, 0.f) : (v < 0.f ? v * p0 : v); // relu.cpp:16-50
    case NCNN_CUDA_UNARY_CLIP: return fminf(fmaxf(v, p0), p1);
    case NCNN_CUDA_UNARY_SIGMOID: return 1.f / (1.f + expf(-v));                           // sigmoid.cpp
    case NCNN_CUDA_UNARY_MISH: return v * tanhf(logf(expf(v) + 1.f));
    case NCNN_CUDA_UNARY_HARDSWISH: return apply_activation(v, 6, p0, p1);
    case NCNN_CUDA_UNARY_SWISH: return v / (1.f + expf(-v));                               // swish.cpp:14-35
    case NCNN_CUDA_UNARY_SCALE: return v * p0;
    case NCNN_CUDA_UNARY_TANH: return tanhf(v);
    case NCNN_CUDA_UNARY_HARDSIGMOID:
    {
        float lower = -p1 / p0, upper = (1.f / p0) + lower;
        return v < lower ? 0.f : (v > upper ? 1.f : v * p0 + p1);
    }
    case NCNN_CUDA_UNARY_GELU: // gelu.cpp:37, :52
        return p0 != 0.f ? 0.5f * v * (1.0f + tanhf(0.79788452f * (v + 0.044715f * v * v * v))) : 0.5f * v * erfcf(-0.70710678f * v);
    }
    return v;
}

// ---------------------------------------------------------------- unary
template<typename T, int VEC>
__global__ void __launch_bounds__(256) unary_flat_kernel(const T* __restrict__ in, T* __restrict__ out, long long count, int op, float p0, float p1)
{
    NC_PDL_PROLOGUE();
    const long long nvec = count / VEC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x)
    {
        float v[VEC];
        load_vec_f32<T, VEC>(in + i * VEC, v);
#pragma unroll
        for (int k = 0; k < VEC; k++) v[k] = unary_apply(op, v[k], p0, p1);
        store_vec_f32<T, VEC>(out + i * VEC, v);
    }
    // tail
    for (long long i = nvec * VEC + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = from_f32<T>(unary_apply(op, to_f32(in[i]), p0, p1));
}

// one thread per (pixel, 16-byte channel vector); in / out may be channel-range views with their own pitches
template<typename T, int VEC>
__global__ void __launch_bounds__(256) unary_pitched_kernel(const T* __restrict__ in, T* __restrict__ out, long long npix, int CV, int icp, int ocp, int op, float p0, float p1)
{
    NC_PDL_PROLOGUE();
    const long long total = npix * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        const long long pix = i / CV;
        const int q = (int)(i - pix * CV) * VEC;
        float v[VEC];
        load_vec_f32<T, VEC>(in + pix * icp + q, v);
#pragma unroll
        for (int k = 0; k < VEC; k++) v[k] = unary_apply(op, v[k], p0, p1);
        store_vec_f32<T, VEC>(out + pix * ocp + q, v);
    }
}

template<typename T>
__global__ void unary_strided_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int P, int C, int icp, long long ins, int ocp, long long ons, int op, float p0, float p1)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * P * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % C);
        long long r = i / C;
        int p = (int)(r % P);
        int b = (int)(r / P);
        out[b * ons + (long long)p * ocp + q] = from_f32<T>(unary_apply(op, to_f32(in[b * ins + (long long)p * icp + q]), p0, p1));
    }
}

template<typename T>
static int run_unary(int op, float p0, float p1, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    if (same_layout(bottom, top) && is_dense(bottom) && aligned16(bottom->data) && aligned16(top->data))
    {
        long long count = flat_count(bottom);
        if (count == 0) return 0;
        NC_PDL_LAUNCH((unary_flat_kernel<T, VEC>), grid_for(count / VEC + 1, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, count, op, p0, p1);
    }
    else if (pitched_vec_ok(bottom, VEC) && pitched_vec_ok(top, VEC))
    {
        TView b = make_view(bottom);
        const long long npix = (long long)b.n * b.P;
        if (npix == 0) return 0;
        NC_PDL_LAUNCH((unary_pitched_kernel<T, VEC>), grid_for(npix * (b.C / VEC), 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, npix, b.C / VEC, bottom->cpitch,
                      top->cpitch, op, p0, p1);
    }
    else
    {
        TView b = make_view(bottom), t = make_view(top);
        long long total = (long long)b.n * b.P * b.C;
        if (total == 0) return 0;
        NC_PDL_LAUNCH((unary_strided_kernel<T>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, b.n, b.P, b.C, b.cpitch, b.nstep, t.cpitch, t.nstep, op, p0, p1);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- eltwise (same-shape, up to 8 inputs per pass)
struct EltArgs
{
    const void* in[8];
    float coeff[8];
    int count;
    int op;       // 0 prod 1 sum 2 max
    int has_coeff;
    int relu;
};

template<typename T, int VEC>
__global__ void __launch_bounds__(256) eltwise_flat_kernel(EltArgs a, T* __restrict__ out, long long count)
{
    NC_PDL_PROLOGUE();
    const long long nvec = (count + VEC - 1) / VEC; // buffers are padded to VEC multiples (cpitch % VEC == 0)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x)
    {
        float acc[VEC];
        load_vec_f32<T, VEC>((const T*)a.in[0] + i * VEC, acc);
        if (a.op == 1 && a.has_coeff)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[k] *= a.coeff[0];
        }
        for (int j = 1; j < a.count; j++)
        {
            float v[VEC];
            load_vec_f32<T, VEC>((const T*)a.in[j] + i * VEC, v);
#pragma unroll
            for (int k = 0; k < VEC; k++)
            {
                if (a.op == 0)
                    acc[k] *= v[k];
                else if (a.op == 1)
                    acc[k] = a.has_coeff ? acc[k] + v[k] * a.coeff[j] : acc[k] + v[k];
                else
                    acc[k] = fmaxf(acc[k], v[k]);
            }
        }
        if (a.relu)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[k] = fmaxf(acc[k], 0.f);
        }
        store_vec_f32<T, VEC>(out + i * VEC, acc);
    }
}

struct EltStrided
{
    int cpitch[8];
    long long nstep[8];
};

template<typename T, int VEC>
__global__ void __launch_bounds__(256) eltwise_pitched_kernel(EltArgs a, EltStrided s, T* __restrict__ out, long long npix, int CV, int ocp)
{
    NC_PDL_PROLOGUE();
    const long long total = npix * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        const long long pix = i / CV;
        const int q = (int)(i - pix * CV) * VEC;
        float acc[VEC];
        load_vec_f32<T, VEC>((const T*)a.in[0] + pix * s.cpitch[0] + q, acc);
        if (a.op == 1 && a.has_coeff)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[k] *= a.coeff[0];
        }
        for (int j = 1; j < a.count; j++)
        {
            float v[VEC];
            load_vec_f32<T, VEC>((const T*)a.in[j] + pix * s.cpitch[j] + q, v);
#pragma unroll
            for (int k = 0; k < VEC; k++)
            {
                if (a.op == 0)
                    acc[k] *= v[k];
                else if (a.op == 1)
                    acc[k] = a.has_coeff ? acc[k] + v[k] * a.coeff[j] : acc[k] + v[k];
                else
                    acc[k] = fmaxf(acc[k], v[k]);
            }
        }
        if (a.relu)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) acc[k] = fmaxf(acc[k], 0.f);
        }
        store_vec_f32<T, VEC>(out + pix * ocp + q, acc);
    }
}

template<typename T>
__global__ void eltwise_strided_kernel(EltArgs a, EltStrided s, T* __restrict__ out, int n, int P, int C, int ocp, long long ons)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * P * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % C);
        long long r = i / C;
        int p = (int)(r % P);
        int b = (int)(r / P);
        float acc = to_f32(((const T*)a.in[0])[b * s.nstep[0] + (long long)p * s.cpitch[0] + q]);
        if (a.op == 1 && a.has_coeff) acc *= a.coeff[0];
        for (int j = 1; j < a.count; j++)
        {
            float v = to_f32(((const T*)a.in[j])[b * s.nstep[j] + (long long)p * s.cpitch[j] + q]);
            if (a.op == 0)
                acc *= v;
            else if (a.op == 1)
                acc = a.has_coeff ? acc + v * a.coeff[j] : acc + v;
            else
                acc = fmaxf(acc, v);
        }
        if (a.relu) acc = fmaxf(acc, 0.f);
        out[b * ons + (long long)p * ocp + q] = from_f32<T>(acc);
    }
}

template<typename T>
static int run_eltwise(EltArgs& a, const ncnn_cuda_tensor* bottoms, const ncnn_cuda_tensor* top, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    bool flat = is_dense(top) && aligned16(top->data) && (top->cpitch % VEC == 0 || flat_count(top) % VEC == 0);
    for (int j = 0; j < a.count && flat; j++) flat = same_layout(&bottoms[j], top) && aligned16(bottoms[j].data);
    if (flat && flat_count(top) % VEC == 0)
    {
        long long count = flat_count(top);
        if (count == 0) return 0;
        NC_PDL_LAUNCH((eltwise_flat_kernel<T, VEC>), grid_for(count / VEC, 256), 256, 0, stream, a, (T*)top->data, count);
    }
    else if ([&]() { bool ok = pitched_vec_ok(top, VEC); for (int j = 0; j < a.count && ok; j++) ok = pitched_vec_ok(&bottoms[j], VEC); return ok; }())
    {
        EltStrided s;
        for (int j = 0; j < a.count; j++)
        {
            s.cpitch[j] = bottoms[j].cpitch;
            s.nstep[j] = bottoms[j].nstep;
        }
        TView t = make_view(top);
        const long long npix = (long long)t.n * t.P;
        if (npix == 0) return 0;
        NC_PDL_LAUNCH((eltwise_pitched_kernel<T, VEC>), grid_for(npix * (t.C / VEC), 256), 256, 0, stream, a, s, (T*)top->data, npix, t.C / VEC, top->cpitch);
    }
    else
    {
        EltStrided s;
        for (int j = 0; j < a.count; j++)
        {
            s.cpitch[j] = bottoms[j].cpitch;
            s.nstep[j] = bottoms[j].nstep;
        }
        TView t = make_view(top);
        long long total = (long long)t.n * t.P * t.C;
        if (total == 0) return 0;
        NC_PDL_LAUNCH((eltwise_strided_kernel<T>), grid_for(total, 256), 256, 0, stream, a, s, (T*)top->data, t.n, t.P, t.C, t.cpitch, t.nstep);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- binary op with broadcasting
__device__ __forceinline__ float binary_apply(int op, float x, float y)
{
    switch (op)
    {
    case 0: return x + y;
    case 1: return x - y;
    case 2: return x * y;
    case 3: return x / y;
    case 4: return fmaxf(x, y);
    case 5: return fminf(x, y);
    case 6: return powf(x, y);
    case 7: return y - x;
    case 8: return y / x;
    case 9: return powf(y, x);
    case 10: return atan2f(x, y);
    case 11: return atan2f(y, x);
    case 12: return fmodf(x, y);
    case 13: return fmodf(y, x);
    case 14:
    {
        float mx = fmaxf(x, y), mn = fminf(x, y);
        return mx + log1pf(expf(mn - mx));
    }
    case 15: return floorf(x / y);
    case 16: return floorf(y / x);
    case 17:
    {
        float r = fmodf(x, y);
        return (r != 0.f && ((r < 0.f) != (y < 0.f))) ? r + y : r;
    }
    case 18:
    {
        float r = fmodf(y, x);
        return (r != 0.f && ((r < 0.f) != (x < 0.f))) ? r + x : r;
    }
    }
    return x;
}

struct BShape
{
    int w, h, d, c;   // logical 4-D extent (1 = broadcast)
    int cpitch;
    long long nstep;  // 0 = broadcast over batch
};

template<typename T>
__global__ void binary_broadcast_kernel(const T* __restrict__ a, BShape sa, const T* __restrict__ b, BShape sb, float scalar, int use_scalar, T* __restrict__ out,
                                        BShape so, int n, int op)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * so.d * so.h * so.w * so.c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % so.c);
        long long r = i / so.c;
        int x = (int)(r % so.w);
        r /= so.w;
        int y = (int)(r % so.h);
        r /= so.h;
        int z = (int)(r % so.d);
        int bb = (int)(r / so.d);
        long long ia = bb * sa.nstep + ((long long)((sa.d == 1 ? 0 : z) * sa.h + (sa.h == 1 ? 0 : y)) * sa.w + (sa.w == 1 ? 0 : x)) * sa.cpitch + (sa.c == 1 ? 0 : q);
        float va = to_f32(a[ia]);
        float vb;
        if (use_scalar)
            vb = scalar;
        else
        {
            long long ib = bb * sb.nstep + ((long long)((sb.d == 1 ? 0 : z) * sb.h + (sb.h == 1 ? 0 : y)) * sb.w + (sb.w == 1 ? 0 : x)) * sb.cpitch + (sb.c == 1 ? 0 : q);
            vb = to_f32(b[ib]);
        }
        long long io = bb * so.nstep + ((long long)(z * so.h + y) * so.w + x) * so.cpitch + q;
        out[io] = from_f32<T>(binary_apply(op, va, vb));
    }
}

template<typename T, int VEC>
__global__ void __launch_bounds__(256) binary_flat_kernel(const T* __restrict__ a, const T* __restrict__ b, float scalar, int use_scalar, T* __restrict__ out, long long count, int op)
{
    NC_PDL_PROLOGUE();
    const long long nvec = count / VEC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x)
    {
        float va[VEC], vb[VEC];
        load_vec_f32<T, VEC>(a + i * VEC, va);
        if (use_scalar)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) vb[k] = scalar;
        }
        else
            load_vec_f32<T, VEC>(b + i * VEC, vb);
#pragma unroll
        for (int k = 0; k < VEC; k++) va[k] = binary_apply(op, va[k], vb[k]);
        store_vec_f32<T, VEC>(out + i * VEC, va);
    }
}

template<typename T, int VEC>
__global__ void __launch_bounds__(256) binary_pitched_kernel(const T* __restrict__ a, const T* __restrict__ b, float scalar, int use_scalar, T* __restrict__ out, long long npix, int CV,
                                                            int acp, int bcp, int ocp, int op)
{
    NC_PDL_PROLOGUE();
    const long long total = npix * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        const long long pix = i / CV;
        const int q = (int)(i - pix * CV) * VEC;
        float va[VEC], vb[VEC];
        load_vec_f32<T, VEC>(a + pix * acp + q, va);
        if (use_scalar)
        {
#pragma unroll
            for (int k = 0; k < VEC; k++) vb[k] = scalar;
        }
        else
            load_vec_f32<T, VEC>(b + pix * bcp + q, vb);
#pragma unroll
        for (int k = 0; k < VEC; k++) va[k] = binary_apply(op, va[k], vb[k]);
        store_vec_f32<T, VEC>(out + pix * ocp + q, va);
    }
}

static BShape bshape(const ncnn_cuda_tensor* t)
{
    BShape s;
    // the device layout treats dims 1/2 as [P][C] with C = w: as a 4-D logical shape that is (w=1..,c=C)
    if (t->dims == 1)
    {
        s.w = 1; s.h = 1; s.d = 1; s.c = t->w;
    }
    else if (t->dims == 2)
    {
        s.w = 1; s.h = t->h; s.d = 1; s.c = t->w;
    }
    else if (t->dims == 3)
    {
        s.w = t->w; s.h = t->h; s.d = 1; s.c = t->c;
    }
    else
    {
        s.w = t->w; s.h = t->h; s.d = t->d; s.c = t->c;
    }
    s.cpitch = t->cpitch;
    s.nstep = (t->n <= 1) ? 0 : t->nstep;
    return s;
}

template<typename T>
static int run_binary(int op, const ncnn_cuda_tensor* a, const ncnn_cuda_tensor* b, float scalar, const ncnn_cuda_tensor* top, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    const int use_scalar = b ? 0 : 1;
    bool flat = same_layout(a, top) && is_dense(top) && aligned16(a->data) && aligned16(top->data) && (flat_count(top) % VEC == 0);
    if (b) flat = flat && same_layout(b, top) && aligned16(b->data);
    if (flat)
    {
        long long count = flat_count(top);
        if (count == 0) return 0;
        NC_PDL_LAUNCH((binary_flat_kernel<T, VEC>), grid_for(count / VEC, 256), 256, 0, stream, (const T*)a->data, b ? (const T*)b->data : 0, scalar, use_scalar, (T*)top->data, count, op);
        NC_LAUNCH_CHECK();
        return 0;
    }
    if (same_shape(a, top) && (!b || same_shape(b, top)) && pitched_vec_ok(a, VEC) && pitched_vec_ok(top, VEC) && (!b || pitched_vec_ok(b, VEC)))
    {
        // same logical shape, operands may be channel-range views with their own pitches
        TView t = make_view(top);
        const long long npix = (long long)t.n * t.P;
        if (npix == 0) return 0;
        NC_PDL_LAUNCH((binary_pitched_kernel<T, VEC>), grid_for(npix * (t.C / VEC), 256), 256, 0, stream, (const T*)a->data, b ? (const T*)b->data : 0, scalar, use_scalar,
                      (T*)top->data, npix, t.C / VEC, a->cpitch, b ? b->cpitch : 0, top->cpitch, op);
        NC_LAUNCH_CHECK();
        return 0;
    }
    BShape sa = bshape(a), so = bshape(top), sb = b ? bshape(b) : so;
    // 1-D/2-D operands are stored like (.., c = w); callers pass rank-matched views (see BinaryOp host layer)
    int n = top->n < 1 ? 1 : top->n;
    so.nstep = top->nstep;
    long long total = (long long)n * so.d * so.h * so.w * so.c;
    if (total == 0) return 0;
    NC_PDL_LAUNCH((binary_broadcast_kernel<T>), grid_for(total, 256), 256, 0, stream, (const T*)a->data, sa, b ? (const T*)b->data : 0, sb, scalar, use_scalar, (T*)top->data, so, n, op);
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- concat / slice along an axis
// Logical 5-D index space [n][c][d][h][w]; `axis_kind`: 0 = c, 1 = d, 2 = h, 3 = w.
struct AxisCopy
{
    int sw, sh, sd, sc;      // extents of the SMALL blob
    int s_cpitch, b_cpitch;  // small / big
    long long s_nstep, b_nstep;
    int bw, bh, bd;          // extents of the big blob (c not needed)
    int axis_kind, offset;
    int to_big;              // 1: small -> big (concat), 0: big -> small (slice)
};

template<typename T, int VEC>
__global__ void __launch_bounds__(256) axis_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, AxisCopy a, int n)
{
    NC_PDL_PROLOGUE();
    const int CV = (a.sc + VEC - 1) / VEC;
    const long long total = (long long)n * a.sd * a.sh * a.sw * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int cv = (int)(i % CV);
        long long r = i / CV;
        int x = (int)(r % a.sw);
        r /= a.sw;
        int y = (int)(r % a.sh);
        r /= a.sh;
        int z = (int)(r % a.sd);
        int b = (int)(r / a.sd);
        int q = cv * VEC;
        int bx = x + (a.axis_kind == 3 ? a.offset : 0);
        int by = y + (a.axis_kind == 2 ? a.offset : 0);
        int bz = z + (a.axis_kind == 1 ? a.offset : 0);
        int bq = q + (a.axis_kind == 0 ? a.offset : 0);
        long long so = b * a.s_nstep + ((long long)(z * a.sh + y) * a.sw + x) * a.s_cpitch + q;
        long long bo = b * a.b_nstep + ((long long)(bz * a.bh + by) * a.bw + bx) * a.b_cpitch + bq;
        if (VEC == 1)
        {
            if (a.to_big)
                dst[bo] = src[so];
            else
                dst[so] = src[bo];
        }
        else
        {
            if (a.to_big)
                *reinterpret_cast<Vec<T, VEC>*>(dst + bo) = *reinterpret_cast<const Vec<T, VEC>*>(src + so);
            else
                *reinterpret_cast<Vec<T, VEC>*>(dst + so) = *reinterpret_cast<const Vec<T, VEC>*>(src + bo);
        }
    }
}

// map the reference's per-rank positive axis to (c,d,h,w) kind, in THIS backend's storage convention
// (dims 1: w is stored as channels; dims 2: h = pixels, w = channels)
static int axis_kind_of(int dims, int axis)
{
    if (dims == 1) return 0;                       // w -> channel lane
    if (dims == 2) return axis == 0 ? 2 : 0;       // 0 = h (pixel rows), 1 = w (channel lane)
    if (dims == 3) return axis == 0 ? 0 : (axis == 1 ? 2 : 3);
    return axis == 0 ? 0 : (axis == 1 ? 1 : (axis == 2 ? 2 : 3));
}

template<typename T>
static int run_axis_copy(const ncnn_cuda_tensor* small_t, const ncnn_cuda_tensor* big_t, int axis, int offset, int to_big, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    BShape ss = bshape(small_t), bs = bshape(big_t);
    AxisCopy a;
    a.sw = ss.w; a.sh = ss.h; a.sd = ss.d; a.sc = ss.c;
    a.bw = bs.w; a.bh = bs.h; a.bd = bs.d;
    a.s_cpitch = small_t->cpitch;
    a.b_cpitch = big_t->cpitch;
    a.s_nstep = small_t->nstep;
    a.b_nstep = big_t->nstep;
    a.axis_kind = axis_kind_of(big_t->dims, axis);
    a.offset = offset;
    a.to_big = to_big;
    int n = big_t->n < 1 ? 1 : big_t->n;
    // concat of an unbatched blob into a batched one: every sample reads the same source
    if (to_big && small_t->n <= 1 && n > 1) a.s_nstep = 0;
    NC_REQUIRE(to_big ? (small_t->n <= 1 || small_t->n == n) : ((small_t->n < 1 ? 1 : small_t->n) == n), "concat/slice: batch mismatch");
    const bool chan_ok = (a.axis_kind != 0) || (offset % VEC == 0);
    // vector path: whole 16-byte channel groups, writing pad lanes of the destination is harmless only when the
    // destination's lanes [q, q+VEC) all belong to this copy: require sc % VEC == 0 unless it is the last slab
    const bool vec_ok = chan_ok && (ss.c % VEC == 0) && (a.s_cpitch % VEC == 0) && (a.b_cpitch % VEC == 0) && (a.s_nstep % VEC == 0) && (a.b_nstep % VEC == 0)
                        && aligned16(small_t->data) && aligned16(big_t->data);
    const T* src = (const T*)(to_big ? small_t->data : big_t->data);
    T* dst = (T*)(to_big ? big_t->data : small_t->data);
    if (vec_ok)
    {
        long long total = (long long)n * a.sd * a.sh * a.sw * (ss.c / VEC);
        if (total == 0) return 0;
        NC_PDL_LAUNCH((axis_copy_kernel<T, VEC>), grid_for(total, 256), 256, 0, stream, src, dst, a, n);
    }
    else
    {
        long long total = (long long)n * a.sd * a.sh * a.sw * ss.c;
        if (total == 0) return 0;
        NC_PDL_LAUNCH((axis_copy_kernel<T, 1>), grid_for(total, 256), 256, 0, stream, src, dst, a, n);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- interp
template<typename T, int VEC>
__global__ void __launch_bounds__(256) interp_nearest_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int C, int inw, int inh, int outw, int outh, float hs, float ws,
                                                            int icp, long long ins, int ocp, long long ons)
{
    NC_PDL_PROLOGUE();
    const int CV = (C + VEC - 1) / VEC;
    const long long total = (long long)n * outh * outw * CV;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int cv = (int)(i % CV);
        long long r = i / CV;
        int x = (int)(r % outw);
        r /= outw;
        int y = (int)(r % outh);
        int b = (int)(r / outh);
        int in_y = min((int)(y * hs), inh - 1);
        int in_x = min((int)(x * ws), inw - 1);
        const T* s = in + b * ins + ((long long)in_y * inw + in_x) * icp + cv * VEC;
        T* d = out + b * ons + ((long long)y * outw + x) * ocp + cv * VEC;
        if (VEC == 1)
            *d = *s;
        else
            *reinterpret_cast<Vec<T, VEC>*>(d) = *reinterpret_cast<const Vec<T, VEC>*>(s);
    }
}

__device__ __forceinline__ void linear_coeff(int w, int outw, int dx, int align_corner, int* sx_out, float* a0, float* a1)
{
    // src/layer/interp.cpp:56-90
    double scale = (double)w / outw;
    if (align_corner) scale = (double)(w - 1) / (outw - 1);
    float fx = (float)((dx + 0.5) * scale - 0.5);
    if (align_corner) fx = (float)(dx * scale);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0)
    {
        sx = 0;
        fx = 0.f;
    }
    if (sx >= w - 1)
    {
        sx = w - 2;
        fx = 1.f;
    }
    *sx_out = sx;
    *a0 = 1.f - fx;
    *a1 = fx;
}

template<typename T>
__global__ void interp_bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int C, int inw, int inh, int outw, int outh, int align_corner, int icp,
                                       long long ins, int ocp, long long ons)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * outh * outw * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % C);
        long long r = i / C;
        int x = (int)(r % outw);
        r /= outw;
        int y = (int)(r % outh);
        int b = (int)(r / outh);
        int sx, sy;
        float a0, a1, b0, b1;
        linear_coeff(inw, outw, x, align_corner, &sx, &a0, &a1);
        linear_coeff(inh, outh, y, align_corner, &sy, &b0, &b1);
        const T* base = in + b * ins + q;
        int sx1 = inw > 1 ? sx + 1 : sx, sy1 = inh > 1 ? sy + 1 : sy;
        if (inw == 1) { sx = 0; sx1 = 0; }
        if (inh == 1) { sy = 0; sy1 = 0; }
        float r0 = to_f32(base[((long long)sy * inw + sx) * icp]) * a0 + to_f32(base[((long long)sy * inw + sx1) * icp]) * a1;
        float r1 = to_f32(base[((long long)sy1 * inw + sx) * icp]) * a0 + to_f32(base[((long long)sy1 * inw + sx1) * icp]) * a1;
        out[b * ons + ((long long)y * outw + x) * ocp + q] = from_f32<T>(r0 * b0 + r1 * b1);
    }
}

// ---------------------------------------------------------------- softmax: one CTA per line along the axis
struct SoftmaxGeom
{
    int ext[4];          // extents of the 4 non-axis dims
    long long str[4];    // their element strides (in and out identical layouts required)
    int L;               // axis length
    long long astride;   // axis stride
};

template<typename T>
__global__ void __launch_bounds__(256) softmax_kernel(const T* __restrict__ in, T* __restrict__ out, SoftmaxGeom g, long long lines)
{
    NC_PDL_PROLOGUE();
    __shared__ float red[32];
    for (long long line = blockIdx.x; line < lines; line += gridDim.x)
    {
        long long r = line, base = 0;
#pragma unroll
        for (int k = 3; k >= 0; k--)
        {
            int idx = (int)(r % g.ext[k]);
            r /= g.ext[k];
            base += idx * g.str[k];
        }
        const T* src = in + base;
        T* dst = out + base;
        // max
        float m = -FLT_MAX;
        for (int i = threadIdx.x; i < g.L; i += blockDim.x) m = fmaxf(m, to_f32(src[i * g.astride]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -FLT_MAX;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (threadIdx.x == 0) red[0] = v;
        }
        __syncthreads();
        m = red[0];
        __syncthreads();
        // sum of exp
        float s = 0.f;
        for (int i = threadIdx.x; i < g.L; i += blockDim.x) s += expf(to_f32(src[i * g.astride]) - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) red[0] = v;
        }
        __syncthreads();
        s = red[0];
        for (int i = threadIdx.x; i < g.L; i += blockDim.x) dst[i * g.astride] = from_f32<T>(expf(to_f32(src[i * g.astride]) - m) / s);
        __syncthreads();
    }
}

// ---------------------------------------------------------------- padding
template<typename T>
__global__ void padding_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int inw, int inh, int inc, int outw, int outh, int outc, int top_pad, int left_pad, int front_pad,
                               int type, float value, int icp, long long ins, int ocp, long long ons)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * outh * outw * outc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % outc);
        long long r = i / outc;
        int x = (int)(r % outw);
        r /= outw;
        int y = (int)(r % outh);
        int b = (int)(r / outh);
        int sx = x - left_pad, sy = y - top_pad, sq = q - front_pad;
        float v;
        if (type == 0)
        {
            bool inside = sx >= 0 && sx < inw && sy >= 0 && sy < inh && sq >= 0 && sq < inc;
            v = inside ? to_f32(in[b * ins + ((long long)sy * inw + sx) * icp + sq]) : value;
        }
        else
        {
            if (type == 1)
            {
                sx = min(max(sx, 0), inw - 1);
                sy = min(max(sy, 0), inh - 1);
                sq = min(max(sq, 0), inc - 1);
            }
            else
            {
                if (sx < 0) sx = -sx;
                if (sx >= inw) sx = 2 * (inw - 1) - sx;
                if (sy < 0) sy = -sy;
                if (sy >= inh) sy = 2 * (inh - 1) - sy;
                if (sq < 0) sq = -sq;
                if (sq >= inc) sq = 2 * (inc - 1) - sq;
            }
            v = to_f32(in[b * ins + ((long long)sy * inw + sx) * icp + sq]);
        }
        out[b * ons + ((long long)y * outw + x) * ocp + q] = from_f32<T>(v);
    }
}

} // namespace

#define NC_DISPATCH_T(elemtype, CALL)                         \
    switch (elemtype)                                         \
    {                                                         \
    case NCNN_CUDA_F32: { typedef float T; return CALL; }     \
    case NCNN_CUDA_BF16: { typedef __nv_bfloat16 T; return CALL; } \
    case NCNN_CUDA_F16: { typedef __half T; return CALL; }    \
    }                                                         \
    return -1

template<typename T>
static int run_interp(int resize_type, int align_corner, float hs, float ws, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    int n = top->n < 1 ? 1 : top->n;
    int C = bottom->c;
    if (resize_type == 1)
    {
        const bool vec_ok = (bottom->cpitch % VEC == 0) && (top->cpitch % VEC == 0) && (bottom->nstep % VEC == 0) && (top->nstep % VEC == 0) && aligned16(bottom->data)
                            && aligned16(top->data) && ((C + VEC - 1) / VEC) * VEC <= bottom->cpitch && ((C + VEC - 1) / VEC) * VEC <= top->cpitch;
        if (vec_ok)
        {
            long long total = (long long)n * top->h * top->w * ((C + VEC - 1) / VEC);
            NC_PDL_LAUNCH((interp_nearest_kernel<T, VEC>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, n, C, bottom->w, bottom->h, top->w, top->h, hs, ws,
                                                                                  bottom->cpitch, bottom->nstep, top->cpitch, top->nstep);
        }
        else
        {
            long long total = (long long)n * top->h * top->w * C;
            NC_PDL_LAUNCH((interp_nearest_kernel<T, 1>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, n, C, bottom->w, bottom->h, top->w, top->h, hs, ws,
                                                                                bottom->cpitch, bottom->nstep, top->cpitch, top->nstep);
        }
    }
    else
    {
        long long total = (long long)n * top->h * top->w * C;
        NC_PDL_LAUNCH((interp_bilinear_kernel<T>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, n, C, bottom->w, bottom->h, top->w, top->h, align_corner,
                                                                          bottom->cpitch, bottom->nstep, top->cpitch, top->nstep);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

template<typename T>
static int run_softmax(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, cudaStream_t stream)
{
    BShape s = bshape(bottom);
    int kind = axis_kind_of(bottom->dims, axis);
    int n = bottom->n < 1 ? 1 : bottom->n;
    // dims order: [n, d, h, w, c] with strides
    int ext5[5] = {n, s.d, s.h, s.w, s.c};
    long long str5[5] = {bottom->nstep, (long long)s.h * s.w * s.cpitch, (long long)s.w * s.cpitch, (long long)s.cpitch, 1};
    int axis5 = kind == 0 ? 4 : (kind == 1 ? 1 : (kind == 2 ? 2 : 3));
    SoftmaxGeom g;
    int k = 0;
    long long lines = 1;
    for (int i = 0; i < 5; i++)
    {
        if (i == axis5) continue;
        g.ext[k] = ext5[i];
        g.str[k] = str5[i];
        lines *= ext5[i];
        k++;
    }
    g.L = ext5[axis5];
    g.astride = str5[axis5];
    if (lines == 0 || g.L == 0) return 0;
    int block = g.L >= 256 ? 256 : (g.L >= 128 ? 128 : (g.L >= 64 ? 64 : 32));
    long long grid = lines < (long long)sm_count() * 16 ? lines : (long long)sm_count() * 16;
    NC_PDL_LAUNCH((softmax_kernel<T>), (int)grid, block, 0, stream, (const T*)bottom->data, (T*)top->data, g, lines);
    NC_LAUNCH_CHECK();
    return 0;
}

template<typename T>
static int run_padding(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int top_pad, int left_pad, int front_pad, int type, float value, cudaStream_t stream)
{
    int n = top->n < 1 ? 1 : top->n;
    long long total = (long long)n * top->h * top->w * top->c;
    if (total == 0) return 0;
    NC_PDL_LAUNCH((padding_kernel<T>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, n, bottom->w, bottom->h, bottom->c, top->w, top->h, top->c, top_pad, left_pad,
                                                              front_pad, type, value, bottom->cpitch, bottom->nstep, top->cpitch, top->nstep);
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- LRN (src/layer/lrn.cpp:26-170)
namespace {

// y = x * (bias + alpha_div_size * sum_{window} x^2) ^ -beta; window = local_size channels around the element
// (region 0, zero outside [0, C)) or local_size x local_size pixels around it (region 1, zero padding)
template<typename T>
__global__ void __launch_bounds__(256) lrn_kernel(const T* __restrict__ in, T* __restrict__ out, int w, int h, int C, int cpitch_in, int cpitch_out, long long nstep_in,
                                                  long long nstep_out, int n, int region, int local_size, float alpha_div_size, float beta, float bias)
{
    NC_PDL_PROLOGUE();
    const int P = w * h;
    const long long total = (long long)n * P * C;
    const int lo = local_size / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int q = (int)(idx % C);
        long long r = idx / C;
        const int pix = (int)(r % P);
        const int b = (int)(r / P);
        const T* ib = in + (long long)b * nstep_in;
        float ss = 0.f;
        if (region == 0)
        {
            const T* px = ib + (long long)pix * cpitch_in;
            for (int c = q - lo; c <= q + lo; c++)
                if (c >= 0 && c < C)
                {
                    const float v = to_f32(px[c]);
                    ss += v * v;
                }
        }
        else
        {
            const int y = pix / w, x = pix - y * w;
            // the reference pads lo before and local_size - lo - 1 after (lrn.cpp:94-100)
            for (int dy = -lo; dy < local_size - lo; dy++)
                for (int dx = -lo; dx < local_size - lo; dx++)
                {
                    const int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
                    {
                        const float v = to_f32(ib[((long long)yy * w + xx) * cpitch_in + q]);
                        ss += v * v;
                    }
                }
        }
        const float x0 = to_f32(ib[(long long)pix * cpitch_in + q]);
        out[(long long)b * nstep_out + (long long)pix * cpitch_out + q] = from_f32<T>(x0 * powf(bias + alpha_div_size * ss, -beta));
    }
}

} // namespace

extern "C" int ncnn_cuda_lrn(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int region_type, int local_size, float alpha, float beta, float bias, void* stream)
{
    NC_REQUIRE(bottom && top && bottom->dims == 3 && same_shape(bottom, top) && bottom->elemtype == top->elemtype && bottom->data != top->data,
               "lrn: a distinct pair of 3-D blobs of one type is required");
    NC_REQUIRE((region_type == 0 || region_type == 1) && local_size > 0, "lrn: bad region_type / local_size");
    TView bv = make_view(bottom), tv = make_view(top);
    const long long total = (long long)bv.n * bv.P * bv.C;
    if (total == 0) return 0;
    const float ads = region_type == 0 ? alpha / local_size : alpha / (local_size * local_size);
    cudaStream_t st = as_stream(stream);
    switch (bottom->elemtype)
    {
    case NCNN_CUDA_F32:
        NC_PDL_LAUNCH((lrn_kernel<float>), grid_for(total, 256), 256, 0, st, (const float*)bottom->data, (float*)top->data, bottom->w, bottom->h, bv.C, bv.cpitch, tv.cpitch, bv.nstep,
                      tv.nstep, bv.n, region_type, local_size, ads, beta, bias);
        break;
    case NCNN_CUDA_BF16:
        NC_PDL_LAUNCH((lrn_kernel<__nv_bfloat16>), grid_for(total, 256), 256, 0, st, (const __nv_bfloat16*)bottom->data, (__nv_bfloat16*)top->data, bottom->w, bottom->h, bv.C, bv.cpitch,
                      tv.cpitch, bv.nstep, tv.nstep, bv.n, region_type, local_size, ads, beta, bias);
        break;
    case NCNN_CUDA_F16:
        NC_PDL_LAUNCH((lrn_kernel<__half>), grid_for(total, 256), 256, 0, st, (const __half*)bottom->data, (__half*)top->data, bottom->w, bottom->h, bv.C, bv.cpitch, tv.cpitch, bv.nstep,
                      tv.nstep, bv.n, region_type, local_size, ads, beta, bias);
        break;
    default: return -1;
    }
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- BatchNorm / Scale / ShuffleChannel
namespace {

// y = x * scale[i] + shift[i] with i = channel (dims 1, 3, 4: the innermost index) or i = row (dims 2); src/layer/batchnorm.cpp:57-120
// (value = b * value + a) and src/layer/scale.cpp:44-168.  One thread per 16-byte channel vector when the channel count allows.
template<typename T, int VEC>
__global__ void __launch_bounds__(256) channel_affine_kernel(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                                                             int P, int C, int in_cpitch, int out_cpitch, long long in_nstep, long long out_nstep, int n, int per_row)
{
    NC_PDL_PROLOGUE();
    const int CV = (C + VEC - 1) / VEC;
    const long long total = (long long)n * P * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int cv = (int)(idx % CV);
        long long r = idx / CV;
        const int pix = (int)(r % P);
        const int b = (int)(r / P);
        const int c0 = cv * VEC;
        float x[VEC];
        load_vec_f32<T, VEC>(in + (long long)b * in_nstep + (long long)pix * in_cpitch + c0, x);
#pragma unroll
        for (int v = 0; v < VEC; v++)
        {
            const int i = per_row ? pix : (c0 + v < C ? c0 + v : C - 1);
            x[v] = fmaf(x[v], scale[i], shift ? shift[i] : 0.f);
        }
        store_vec_f32<T, VEC>(out + (long long)b * out_nstep + (long long)pix * out_cpitch + c0, x);
    }
}

template<typename T>
static int run_channel_affine(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const float* scale, const float* shift, int per_row, cudaStream_t stream)
{
    TView bv = make_view(bottom), tv = make_view(top);
    constexpr int VEC = 16 / (int)sizeof(T);
    const int cround = ((bv.C + VEC - 1) / VEC) * VEC;
    const bool vec_ok = (bv.cpitch % VEC == 0) && (tv.cpitch % VEC == 0) && (bv.nstep % VEC == 0) && (tv.nstep % VEC == 0) && (((uintptr_t)bottom->data & 15) == 0)
                        && (((uintptr_t)top->data & 15) == 0) && cround <= bv.cpitch && cround <= tv.cpitch;
    if (vec_ok)
    {
        const long long total = (long long)bv.n * bv.P * (cround / VEC);
        NC_PDL_LAUNCH((channel_affine_kernel<T, VEC>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, scale, shift, bv.P, bv.C, bv.cpitch, tv.cpitch, bv.nstep,
                                                                               tv.nstep, bv.n, per_row);
    }
    else
    {
        const long long total = (long long)bv.n * bv.P * bv.C;
        NC_PDL_LAUNCH((channel_affine_kernel<T, 1>), grid_for(total, 256), 256, 0, stream, (const T*)bottom->data, (T*)top->data, scale, shift, bv.P, bv.C, bv.cpitch, tv.cpitch, bv.nstep,
                                                                             tv.nstep, bv.n, per_row);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

// out[pixel][group * j + i] = in[pixel][cpg * i + j]  (src/layer/shufflechannel.cpp:30-57); consecutive threads write consecutive channels
template<typename T>
__global__ void __launch_bounds__(256) shuffle_channel_kernel(const T* __restrict__ in, T* __restrict__ out, int P, int C, int group, int in_cpitch, int out_cpitch,
                                                              long long in_nstep, long long out_nstep, int n)
{
    NC_PDL_PROLOGUE();
    const int cpg = C / group;
    const long long total = (long long)n * P * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int dq = (int)(idx % C);
        long long r = idx / C;
        const int pix = (int)(r % P);
        const int b = (int)(r / P);
        const int j = dq / group, i = dq - j * group;
        out[(long long)b * out_nstep + (long long)pix * out_cpitch + dq] = in[(long long)b * in_nstep + (long long)pix * in_cpitch + cpg * i + j];
    }
}

} // namespace

extern "C" {

int ncnn_cuda_channel_affine(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, const float* scale_dev, const float* shift_dev, void* stream)
{
    NC_REQUIRE(bottom && top && scale_dev && same_shape(bottom, top) && bottom->elemtype == top->elemtype, "channel_affine: shape/type mismatch");
    const int per_row = bottom->dims == 2;
    NC_DISPATCH_T(bottom->elemtype, run_channel_affine<T>(bottom, top, scale_dev, shift_dev, per_row, as_stream(stream)));
}

int ncnn_cuda_shuffle_channel(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int group, void* stream)
{
    NC_REQUIRE(bottom && top && same_shape(bottom, top) && bottom->elemtype == top->elemtype && bottom->dims >= 3, "shuffle_channel: a 3-D/4-D blob pair of one type is required");
    NC_REQUIRE(group > 0 && bottom->c % group == 0, "shuffle_channel: channels not divisible by group");
    TView bv = make_view(bottom), tv = make_view(top);
    const long long total = (long long)bv.n * bv.P * bv.C;
    if (total == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (bottom->elemtype == NCNN_CUDA_F32)
        NC_PDL_LAUNCH((shuffle_channel_kernel<float>), grid_for(total, 256), 256, 0, st, (const float*)bottom->data, (float*)top->data, bv.P, bv.C, group, bv.cpitch, tv.cpitch, bv.nstep, tv.nstep, bv.n);
    else
        NC_PDL_LAUNCH((shuffle_channel_kernel<uint16_t>), grid_for(total, 256), 256, 0, st, (const uint16_t*)bottom->data, (uint16_t*)top->data, bv.P, bv.C, group, bv.cpitch, tv.cpitch, bv.nstep,
                                                                              tv.nstep, bv.n);
    NC_LAUNCH_CHECK();
    return 0;
}

} // extern "C"

extern "C" {

int ncnn_cuda_unary(int op, float p0, float p1, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(same_shape(bottom, top) && bottom->elemtype == top->elemtype, "unary: shape/type mismatch");
    NC_DISPATCH_T(bottom->elemtype, run_unary<T>(op, p0, p1, bottom, top, as_stream(stream)));
}

int ncnn_cuda_eltwise(int op, const ncnn_cuda_tensor* bottoms, int count, const float* coeffs, int relu, const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(count >= 2, "eltwise: needs at least two bottoms");
    NC_REQUIRE(op >= 0 && op <= 2, "eltwise: bad op");
    for (int j = 0; j < count; j++) NC_REQUIRE(same_shape(&bottoms[j], top) && bottoms[j].elemtype == top->elemtype, "eltwise: shape/type mismatch");
    // more than 8 inputs: fold in passes of 8 with `top` as the running value
    int done = 0;
    while (done < count)
    {
        EltArgs a;
        ncnn_cuda_tensor group[8];
        a.count = 0;
        if (done > 0)
        {
            group[0] = *top;
            a.in[0] = top->data;
            a.coeff[0] = 1.f;
            a.count = 1;
        }
        while (a.count < 8 && done < count)
        {
            group[a.count] = bottoms[done];
            a.in[a.count] = bottoms[done].data;
            a.coeff[a.count] = coeffs ? coeffs[done] : 1.f;
            a.count++;
            done++;
        }
        a.op = op;
        a.has_coeff = (op == 1 && coeffs) ? 1 : 0;
        a.relu = (done == count) ? relu : 0;
        int r;
        switch (top->elemtype)
        {
        case NCNN_CUDA_F32: r = run_eltwise<float>(a, group, top, as_stream(stream)); break;
        case NCNN_CUDA_BF16: r = run_eltwise<__nv_bfloat16>(a, group, top, as_stream(stream)); break;
        case NCNN_CUDA_F16: r = run_eltwise<__half>(a, group, top, as_stream(stream)); break;
        default: r = -1;
        }
        if (r != 0) return r;
    }
    return 0;
}

int ncnn_cuda_binaryop(int op, const ncnn_cuda_tensor* a, const ncnn_cuda_tensor* b, float scalar, const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(op >= 0 && op <= 18, "binaryop: bad op");
    NC_REQUIRE(a->elemtype == top->elemtype && (!b || b->elemtype == top->elemtype), "binaryop: element types differ");
    NC_DISPATCH_T(top->elemtype, run_binary<T>(op, a, b, scalar, top, as_stream(stream)));
}

int ncnn_cuda_copy_into_axis(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, int offset, void* stream)
{
    NC_REQUIRE(bottom->dims == top->dims && bottom->elemtype == top->elemtype, "concat: rank/type mismatch");
    NC_DISPATCH_T(top->elemtype, run_axis_copy<T>(bottom, top, axis, offset, 1, as_stream(stream)));
}

int ncnn_cuda_copy_from_axis(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, int offset, void* stream)
{
    NC_REQUIRE(bottom->dims == top->dims && bottom->elemtype == top->elemtype, "slice: rank/type mismatch");
    NC_DISPATCH_T(top->elemtype, run_axis_copy<T>(top, bottom, axis, offset, 0, as_stream(stream)));
}


int ncnn_cuda_interp(int resize_type, int align_corner, float hs, float ws, const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(bottom->dims == 3 && top->dims == 3 && bottom->c == top->c && bottom->elemtype == top->elemtype, "interp: 3-D blobs with equal channels required");
    NC_REQUIRE(resize_type == 1 || resize_type == 2, "interp: only nearest (1) and bilinear (2) are implemented");
    NC_DISPATCH_T(top->elemtype, run_interp<T>(resize_type, align_corner, hs, ws, bottom, top, as_stream(stream)));
}


int ncnn_cuda_softmax(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int axis, void* stream)
{
    NC_REQUIRE(same_shape(bottom, top) && bottom->elemtype == top->elemtype && bottom->cpitch == top->cpitch && bottom->nstep == top->nstep, "softmax: layouts differ");
    NC_REQUIRE(axis >= 0 && axis < bottom->dims, "softmax: bad axis");
    NC_DISPATCH_T(top->elemtype, run_softmax<T>(bottom, top, axis, as_stream(stream)));
}


int ncnn_cuda_padding(const ncnn_cuda_tensor* bottom, const ncnn_cuda_tensor* top, int top_pad, int left_pad, int front_pad, int type, float value, void* stream)
{
    NC_REQUIRE(bottom->dims == 3 && top->dims == 3 && bottom->elemtype == top->elemtype, "padding: 3-D blobs required");
    NC_REQUIRE(type >= 0 && type <= 2, "padding: bad type");
    NC_DISPATCH_T(top->elemtype, run_padding<T>(bottom, top, top_pad, left_pad, front_pad, type, value, as_stream(stream)));
}

} // extern "C"
