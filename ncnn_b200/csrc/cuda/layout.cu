// layout.cu -- conversions between the reference's planar Mat layout (src/mat.cpp:299-861) and this
// backend's channel-innermost device layout, plus Reshape / Flatten / Permute / cast / clone.
#include "common.cuh"

namespace ncnn_cuda {

struct DShape
{
    int dims, w, h, d, c;
    int cpitch;
    long long nstep;

    __host__ __device__ long long logical_count() const
    {
        if (dims == 1) return w;
        if (dims == 2) return (long long)w * h;
        if (dims == 3) return (long long)w * h * c;
        return (long long)w * h * d * c;
    }
    // logical flat index (ncnn dense order: c, d, h, w) -> physical offset within a sample
    __device__ long long phys_of_logical(long long i) const
    {
        if (dims == 1) return i;
        if (dims == 2)
        {
            long long y = i / w;
            int x = (int)(i - y * w);
            return y * cpitch + x;
        }
        long long plane = (long long)w * h * (dims == 4 ? d : 1);
        long long q = i / plane;
        long long p = i - q * plane;
        return p * cpitch + q;
    }
};

static inline DShape dshape(const ncnn_cuda_tensor* t)
{
    DShape s;
    s.dims = t->dims;
    s.w = t->w;
    s.h = t->dims >= 2 ? t->h : 1;
    s.d = t->dims == 4 ? t->d : 1;
    s.c = t->dims >= 3 ? t->c : 1;
    s.cpitch = t->cpitch;
    s.nstep = t->nstep;
    return s;
}

// ---------------------------------------------------------------- planar <-> pixels x channels
// Tiled transpose through shared memory so both sides are coalesced: tile = 32 pixels x 32 channels.
template<typename TD, bool PACK>
__global__ void planar_transpose_kernel(float* __restrict__ planar, long long cstep, long long pl_nstep,
                                        TD* __restrict__ dev, int cpitch, long long dv_nstep, int P, int C)
{
    NC_PDL_PROLOGUE();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32;
    const int q0 = blockIdx.y * 32;
    float* pl = planar + (long long)b * pl_nstep;
    TD* dv = dev + (long long)b * dv_nstep;
    if (PACK)
    {
        // read planar: threads along pixels
        for (int j = threadIdx.y; j < 32; j += blockDim.y)
        {
            int q = q0 + j, p = p0 + threadIdx.x;
            if (q < C && p < P) tile[j][threadIdx.x] = pl[(long long)q * cstep + p];
        }
        __syncthreads();
        // write device: threads along channels
        for (int j = threadIdx.y; j < 32; j += blockDim.y)
        {
            int p = p0 + j, q = q0 + threadIdx.x;
            if (p < P && q < cpitch) dv[(long long)p * cpitch + q] = from_f32<TD>(q < C ? tile[threadIdx.x][j] : 0.f);
        }
    }
    else
    {
        for (int j = threadIdx.y; j < 32; j += blockDim.y)
        {
            int p = p0 + j, q = q0 + threadIdx.x;
            if (p < P && q < C) tile[j][threadIdx.x] = to_f32(dv[(long long)p * cpitch + q]);
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += blockDim.y)
        {
            int q = q0 + j, p = p0 + threadIdx.x;
            if (q < C && p < P) pl[(long long)q * cstep + p] = tile[threadIdx.x][j];
        }
    }
}

// Few channels (image inputs: C <= 8).  The 32 x 32 tile above keeps 3 of its 32 channel lanes busy on an RGB batch -- 1 ms for the
// 154 MB fp32 input of ResNet-50 bs256, a third of the network's own time on the end-to-end path.  Here a thread owns one pixel:
// C coalesced plane reads (consecutive threads, consecutive pixels), one 16 / 32-byte pixel store (or the reverse).
template<typename TD, bool PACK, int CP>
__global__ void __launch_bounds__(256) planar_fewc_kernel(float* __restrict__ planar, long long cstep, long long pl_nstep, TD* __restrict__ dev, long long dv_nstep, int P,
                                                          int C, int n)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        const int b = (int)(i / P);
        const int p = (int)(i - (long long)b * P);
        float* pl = planar + (long long)b * pl_nstep + p;
        TD* dv = dev + (long long)b * dv_nstep + (long long)p * CP;
        if (PACK)
        {
            Vec<TD, CP> t;
#pragma unroll
            for (int q = 0; q < CP; q++) t.v[q] = from_f32<TD>(q < C ? __ldg(pl + (long long)q * cstep) : 0.f);
            *reinterpret_cast<Vec<TD, CP>*>(dv) = t;
        }
        else
        {
            const Vec<TD, CP> t = *reinterpret_cast<const Vec<TD, CP>*>(dv);
#pragma unroll
            for (int q = 0; q < CP; q++)
                if (q < C) pl[(long long)q * cstep] = to_f32(t.v[q]);
        }
    }
}

// dims 1/2: both sides are row-major [P][C]; only pitch and dtype differ
template<typename TD, bool PACK>
__global__ void planar_rows_kernel(float* __restrict__ planar, long long pl_nstep, TD* __restrict__ dev, int cpitch,
                                   long long dv_nstep, int P, int C, int n)
{
    NC_PDL_PROLOGUE();
    long long total = (long long)n * P * cpitch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % cpitch);
        long long r = i / cpitch;
        int p = (int)(r % P);
        int b = (int)(r / P);
        float* pl = planar + (long long)b * pl_nstep + (long long)p * C;
        TD* dv = dev + (long long)b * dv_nstep + (long long)p * cpitch;
        if (PACK)
            dv[q] = from_f32<TD>(q < C ? pl[q] : 0.f);
        else if (q < C)
            pl[q] = to_f32(dv[q]);
    }
}

template<typename TD, bool PACK>
static int planar_convert(const ncnn_cuda_hostmat* hm, const ncnn_cuda_tensor* t, cudaStream_t stream)
{
    TView v = make_view(t);
    NC_REQUIRE(hm->dims == t->dims && hm->w == t->w && (t->dims < 2 || hm->h == t->h) && (t->dims < 3 || hm->c == t->c) && (t->dims < 4 || hm->d == t->d),
               "pack/unpack: host and device shapes differ");
    int n = v.n;
    if (t->dims <= 2)
    {
        long long total = (long long)n * v.P * v.cpitch;
        NC_PDL_LAUNCH((planar_rows_kernel<TD, PACK>), grid_for(total, 256), 256, 0, stream, (float*)hm->data, hm->nstep, (TD*)t->data, v.cpitch, v.nstep, v.P, v.C, n);
        NC_LAUNCH_CHECK();
        return 0;
    }
    // image-like blobs: one thread per pixel when the whole (padded) pixel is one vector store
    if ((v.cpitch == 4 || v.cpitch == 8) && v.cpitch * sizeof(TD) >= 16 && v.cpitch * sizeof(TD) <= 32 && (((uintptr_t)t->data) & 31) == 0 &&
            (v.nstep * (long long)sizeof(TD)) % 32 == 0)
    {
        const long long total = (long long)n * v.P;
        if (v.cpitch == 4)
            NC_PDL_LAUNCH((planar_fewc_kernel<TD, PACK, 4>), grid_for(total, 256, 16), 256, 0, stream, (float*)hm->data, hm->cstep, hm->nstep, (TD*)t->data, v.nstep, v.P, v.C, n);
        else
            NC_PDL_LAUNCH((planar_fewc_kernel<TD, PACK, 8>), grid_for(total, 256, 16), 256, 0, stream, (float*)hm->data, hm->cstep, hm->nstep, (TD*)t->data, v.nstep, v.P, v.C, n);
        NC_LAUNCH_CHECK();
        return 0;
    }
    dim3 block(32, 8);
    int cspan = PACK ? v.cpitch : v.C;
    dim3 grid(ceil_div(v.P, 32), ceil_div(cspan, 32), n);
    NC_PDL_LAUNCH((planar_transpose_kernel<TD, PACK>), grid, block, 0, stream, (float*)hm->data, hm->cstep, hm->nstep, (TD*)t->data, v.cpitch, v.nstep, v.P, v.C);
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- generic gather (reshape / permute / cast)
struct PermuteMap
{
    int src_of[4]; // for out logical dim i (0=w,1=h,2=d,3=c): which src logical dim supplies it
};

template<typename TS, typename TD, bool PERMUTE>
__global__ void gather_kernel(const TS* __restrict__ src, DShape ss, TD* __restrict__ dst, DShape ds, int n, PermuteMap pm)
{
    NC_PDL_PROLOGUE();
    const int dP = ds.dims == 1 ? 1 : (ds.dims == 2 ? ds.h : ds.w * ds.h * ds.d);
    const int dC = ds.dims == 1 ? ds.w : (ds.dims == 2 ? ds.w : ds.c);
    long long total = (long long)n * dP * dC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        int q = (int)(i % dC);
        long long r = i / dC;
        int p = (int)(r % dP);
        int b = (int)(r / dP);
        // out logical coords
        int ox, oy, oz, oc;
        if (ds.dims == 1)
        {
            ox = q; oy = 0; oz = 0; oc = 0;
        }
        else if (ds.dims == 2)
        {
            ox = q; oy = p; oz = 0; oc = 0;
        }
        else
        {
            ox = p % ds.w;
            int t = p / ds.w;
            oy = t % ds.h;
            oz = t / ds.h;
            oc = q;
        }
        long long sphys;
        if (PERMUTE)
        {
            int oc4[4] = {ox, oy, oz, oc};
            int sc[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 4; k++) sc[pm.src_of[k]] = oc4[k];
            // src logical coords (x,y,z,ch)
            if (ss.dims == 1)
                sphys = sc[0];
            else if (ss.dims == 2)
                sphys = (long long)sc[1] * ss.cpitch + sc[0];
            else
                sphys = ((long long)(sc[2] * ss.h + sc[1]) * ss.w + sc[0]) * ss.cpitch + sc[3];
        }
        else
        {
            long long flat = (((long long)oc * ds.d + oz) * ds.h + oy) * ds.w + ox;
            sphys = ss.phys_of_logical(flat);
        }
        long long dphys = ds.dims == 1 ? q : (long long)p * ds.cpitch + q;
        dst[(long long)b * ds.nstep + dphys] = from_f32<TD>(to_f32(src[(long long)b * ss.nstep + sphys]));
    }
}

// Batched 2-D transpose dst[b][j][i] = src[b][i][j] (rows x cols -> cols x rows), 32x32 shared-memory tiles: both sides
// coalesced.  In the channel-innermost device layout a Reshape (w,h,c) -> (w*h, c) and a 2-D Permute are exactly this
// (YOLOv8's detection head does both on its largest blobs).
template<typename T>
__global__ void __launch_bounds__(256) transpose2d_kernel(const T* __restrict__ src, int rows, int cols, int spitch, long long snstep, T* __restrict__ dst, int dpitch,
                                                          long long dnstep)
{
    NC_PDL_PROLOGUE();
    __shared__ T tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const T* s = src + (long long)b * snstep;
    T* d = dst + (long long)b * dnstep;
    const int tx = threadIdx.x, ty = threadIdx.y; // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8)
    {
        const int r = r0 + ty + k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + k][tx] = s[(long long)r * spitch + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8)
    {
        const int c = c0 + ty + k, r = r0 + tx;
        if (r < rows && c < cols) d[(long long)c * dpitch + r] = tile[tx][ty + k];
    }
}

// src viewed as [n][rows][cols] (pitch spitch) -> dst [n][cols][rows] (pitch dpitch); same element type
static int launch_transpose2d(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, int rows, int cols, int spitch, int dpitch, cudaStream_t stream)
{
    const int n = dst->n < 1 ? 1 : dst->n;
    if (rows == 0 || cols == 0) return 0;
    dim3 block(32, 8), grid(ceil_div(cols, 32), ceil_div(rows, 32), n);
    if (grid.y > 65535 || grid.z > 65535) return 1;
    if (src->elemtype == NCNN_CUDA_F32)
        NC_PDL_LAUNCH((transpose2d_kernel<float>), grid, block, 0, stream, (const float*)src->data, rows, cols, spitch, src->nstep, (float*)dst->data, dpitch, dst->nstep);
    else
        NC_PDL_LAUNCH((transpose2d_kernel<uint16_t>), grid, block, 0, stream, (const uint16_t*)src->data, rows, cols, spitch, src->nstep, (uint16_t*)dst->data, dpitch, dst->nstep);
    NC_LAUNCH_CHECK();
    return 0;
}

template<typename TS, typename TD>
static int launch_gather(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, bool permute, const PermuteMap& pm, cudaStream_t stream)
{
    DShape ss = dshape(src), ds = dshape(dst);
    int n = dst->n < 1 ? 1 : dst->n;
    long long total = (long long)n * ds.logical_count();
    if (total == 0) return 0;
    if (permute)
        NC_PDL_LAUNCH((gather_kernel<TS, TD, true>), grid_for(total, 256), 256, 0, stream, (const TS*)src->data, ss, (TD*)dst->data, ds, n, pm);
    else
        NC_PDL_LAUNCH((gather_kernel<TS, TD, false>), grid_for(total, 256), 256, 0, stream, (const TS*)src->data, ss, (TD*)dst->data, ds, n, pm);
    NC_LAUNCH_CHECK();
    return 0;
}

static int dispatch_gather(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, bool permute, const PermuteMap& pm, cudaStream_t stream)
{
#define NC_G(TS, TD) return launch_gather<TS, TD>(src, dst, permute, pm, stream)
    int s = src->elemtype, d = dst->elemtype;
    if (s == NCNN_CUDA_F32 && d == NCNN_CUDA_F32) NC_G(float, float);
    if (s == NCNN_CUDA_F32 && d == NCNN_CUDA_BF16) NC_G(float, __nv_bfloat16);
    if (s == NCNN_CUDA_F32 && d == NCNN_CUDA_F16) NC_G(float, __half);
    if (s == NCNN_CUDA_BF16 && d == NCNN_CUDA_F32) NC_G(__nv_bfloat16, float);
    if (s == NCNN_CUDA_BF16 && d == NCNN_CUDA_BF16) NC_G(__nv_bfloat16, __nv_bfloat16);
    if (s == NCNN_CUDA_F16 && d == NCNN_CUDA_F32) NC_G(__half, float);
    if (s == NCNN_CUDA_F16 && d == NCNN_CUDA_F16) NC_G(__half, __half);
#undef NC_G
    set_last_error_msg("reshape/permute: unsupported element type pair");
    return -1;
}

} // namespace ncnn_cuda

using namespace ncnn_cuda;

// ---------------------------------------------------------------- interleaved 8-bit pixels -> device blob
// Mat::from_pixels (src/mat_pixel.cpp: from_rgb / from_rgb2bgr / from_gray / from_rgba) followed by
// Mat::substract_mean_normalize (src/mat.cpp): value = (pixel - mean[c]) * norm[c].  The interleaved HWC byte order IS the
// channel-innermost device order, so one thread converts one pixel and writes its channel vector; no transpose, and the
// fp32 planar image never exists (4x fewer bytes over PCIe than uploading the converted Mat).
struct PixelAffine
{
    float mean[4];
    float norm[4];
};

template<typename T, int CH>
__global__ void __launch_bounds__(256) pixels_to_blob_kernel(const unsigned char* __restrict__ pixels, int w, int h, int stride, long long nstride, int swap_rb, PixelAffine pa,
                                                             T* __restrict__ out, int cpitch, long long out_nstep, int n)
{
    NC_PDL_PROLOGUE();
    const long long total = (long long)n * h * w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int x = (int)(idx % w);
        long long r = idx / w;
        const int y = (int)(r % h);
        const int b = (int)(r / h);
        const unsigned char* px = pixels + (long long)b * nstride + (long long)y * stride + (long long)x * CH;
        T* o = out + (long long)b * out_nstep + ((long long)y * w + x) * cpitch;
#pragma unroll
        for (int c = 0; c < CH; c++)
        {
            // swap_rb reverses the first three channels (RGB <-> BGR; a fourth, alpha, stays in place)
            const int sc = (swap_rb && c < 3) ? 2 - c : c;
            o[c] = from_f32<T>(((float)px[sc] - pa.mean[c]) * pa.norm[c]);
        }
        for (int c = CH; c < cpitch; c++) o[c] = from_f32<T>(0.f); // padding lanes: zeros (the stem kernels read whole pixels)
    }
}

template<typename T>
static int run_pixels(const unsigned char* pixels, int ch, int w, int h, int stride, long long nstride, int swap_rb, const PixelAffine& pa, const ncnn_cuda_tensor* top,
                      cudaStream_t stream)
{
    TView tv = make_view(top);
    const long long total = (long long)tv.n * h * w;
    if (total == 0) return 0;
    const int grid = grid_for(total, 256);
    if (ch == 1)
        NC_PDL_LAUNCH((pixels_to_blob_kernel<T, 1>), grid, 256, 0, stream, pixels, w, h, stride, nstride, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    else if (ch == 3)
        NC_PDL_LAUNCH((pixels_to_blob_kernel<T, 3>), grid, 256, 0, stream, pixels, w, h, stride, nstride, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    else
        NC_PDL_LAUNCH((pixels_to_blob_kernel<T, 4>), grid, 256, 0, stream, pixels, w, h, stride, nstride, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    NC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- the same with the reference's bilinear resize in front
// Mat::from_pixels_resize (src/mat_pixel.cpp:2546-2578) = resize_bilinear_c1/c3/c4 (src/mat_pixel_resize.cpp:210-1039) on the
// 8-bit image, then from_pixels.  The reference's resize is integer arithmetic on 11-bit coefficients:
//   row value  r = (S[sx] * a0 + S[sx + 1] * a1) >> 4                       (hresize, :789-794)
//   pixel      v = (((b0 * r0) >> 16) + ((b1 * r1) >> 16) + 2) >> 2          (vresize_one, :182-186), saturated to 8 bits,
// a pure function of four source bytes per channel, so one thread produces one output pixel bit-exactly; the per-column /
// per-row source offsets and coefficients (the float -> short rounding of :617-669) come from ncnn_cuda_resize_tables,
// computed once on the host exactly as the reference computes them.
template<typename T, int CH>
__global__ void __launch_bounds__(256) pixels_resize_to_blob_kernel(const unsigned char* __restrict__ pixels, int stride, long long nstride, const int* __restrict__ tab, int w, int h,
                                                                    int swap_rb, PixelAffine pa, T* __restrict__ out, int cpitch, long long out_nstep, int n)
{
    NC_PDL_PROLOGUE();
    const int* xofs = tab;
    const int* yofs = tab + w;
    const int* alpha = tab + w + h;      // a0, a1 per column
    const int* beta = tab + w + h + 2 * w; // b0, b1 per row
    const long long total = (long long)n * h * w;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        const int x = (int)(idx % w);
        long long r = idx / w;
        const int y = (int)(r % h);
        const int b = (int)(r / h);
        const int sx = xofs[x], sy = yofs[y];
        const int a0 = alpha[2 * x], a1 = alpha[2 * x + 1], b0 = beta[2 * y], b1 = beta[2 * y + 1];
        const unsigned char* s0 = pixels + (long long)b * nstride + (long long)sy * stride + (long long)sx * CH;
        const unsigned char* s1 = s0 + stride;
        int v[CH];
#pragma unroll
        for (int c = 0; c < CH; c++)
        {
            const int r0 = (s0[c] * a0 + s0[c + CH] * a1) >> 4;
            const int r1 = (s1[c] * a0 + s1[c + CH] * a1) >> 4;
            const int q = (((b0 * r0) >> 16) + ((b1 * r1) >> 16) + 2) >> 2;
            v[c] = q < 0 ? 0 : (q > 255 ? 255 : q);
        }
        T* o = out + (long long)b * out_nstep + ((long long)y * w + x) * cpitch;
#pragma unroll
        for (int c = 0; c < CH; c++)
        {
            const int sc = (swap_rb && c < 3) ? 2 - c : c;
            o[c] = from_f32<T>(((float)v[sc] - pa.mean[c]) * pa.norm[c]);
        }
        for (int c = CH; c < cpitch; c++) o[c] = from_f32<T>(0.f);
    }
}

template<typename T>
static int run_pixels_resize(const unsigned char* pixels, int ch, int stride, long long nstride, const int* tab, int swap_rb, const PixelAffine& pa, const ncnn_cuda_tensor* top,
                             cudaStream_t stream)
{
    TView tv = make_view(top);
    const int w = top->w, h = top->h;
    const long long total = (long long)tv.n * h * w;
    if (total == 0) return 0;
    const int grid = grid_for(total, 256);
    if (ch == 1)
        NC_PDL_LAUNCH((pixels_resize_to_blob_kernel<T, 1>), grid, 256, 0, stream, pixels, stride, nstride, tab, w, h, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    else if (ch == 3)
        NC_PDL_LAUNCH((pixels_resize_to_blob_kernel<T, 3>), grid, 256, 0, stream, pixels, stride, nstride, tab, w, h, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    else
        NC_PDL_LAUNCH((pixels_resize_to_blob_kernel<T, 4>), grid, 256, 0, stream, pixels, stride, nstride, tab, w, h, swap_rb, pa, (T*)top->data, tv.cpitch, tv.nstep, tv.n);
    NC_LAUNCH_CHECK();
    return 0;
}

// round half away from zero, saturated to short: SATURATE_CAST_SHORT of src/mat_pixel_resize.cpp:619
static int resize_coef(float v)
{
    int i = (int)(v + (v >= 0.f ? 0.5f : -0.5f));
    return i < -32768 ? -32768 : (i > 32767 ? 32767 : i);
}

extern "C" {

int ncnn_cuda_resize_tables_count(int w, int h)
{
    return 3 * (w + h);
}

// host-only: [xofs(w) | yofs(h) | a0,a1 per column (2w) | b0,b1 per row (2h)]; src/mat_pixel_resize.cpp:599-669
int ncnn_cuda_resize_tables(int src_w, int src_h, int w, int h, int* tables)
{
    NC_REQUIRE(tables && src_w >= 2 && src_h >= 2 && w > 0 && h > 0, "resize_tables: the source must be at least 2 x 2");
    const double scale_x = (double)src_w / w;
    const double scale_y = (double)src_h / h;
    int* xofs = tables;
    int* yofs = tables + w;
    int* alpha = tables + w + h;
    int* beta = tables + w + h + 2 * w;
    for (int dx = 0; dx < w; dx++)
    {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floor(fx);
        fx -= sx;
        if (sx < 0)
        {
            sx = 0;
            fx = 0.f;
        }
        if (sx >= src_w - 1)
        {
            sx = src_w - 2;
            fx = 1.f;
        }
        xofs[dx] = sx;
        alpha[2 * dx] = resize_coef((1.f - fx) * 2048);
        alpha[2 * dx + 1] = resize_coef(fx * 2048);
    }
    for (int dy = 0; dy < h; dy++)
    {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floor(fy);
        fy -= sy;
        if (sy < 0)
        {
            sy = 0;
            fy = 0.f;
        }
        if (sy >= src_h - 1)
        {
            sy = src_h - 2;
            fy = 1.f;
        }
        yofs[dy] = sy;
        beta[2 * dy] = resize_coef((1.f - fy) * 2048);
        beta[2 * dy + 1] = resize_coef(fy * 2048);
    }
    return 0;
}

int ncnn_cuda_pixels_resize_to_blob(const unsigned char* pixels_dev, int channels, int src_w, int src_h, int stride, long long nstride, int swap_rb, const float* mean_vals,
                                    const float* norm_vals, const int* tables_dev, const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(pixels_dev && tables_dev && top && top->dims == 3 && top->c == channels, "pixels_resize_to_blob: the top blob must be (target w, target h, channels)");
    NC_REQUIRE(channels == 1 || channels == 3 || channels == 4, "pixels_resize_to_blob: 1, 3 or 4 interleaved channels");
    NC_REQUIRE(src_w >= 2 && src_h >= 2 && stride >= src_w * channels, "pixels_resize_to_blob: bad source geometry");
    PixelAffine pa;
    for (int c = 0; c < 4; c++)
    {
        pa.mean[c] = (mean_vals && c < channels) ? mean_vals[c] : 0.f;
        pa.norm[c] = (norm_vals && c < channels) ? norm_vals[c] : 1.f;
    }
    switch (top->elemtype)
    {
    case NCNN_CUDA_F32: return run_pixels_resize<float>(pixels_dev, channels, stride, nstride, tables_dev, swap_rb, pa, top, as_stream(stream));
    case NCNN_CUDA_BF16: return run_pixels_resize<__nv_bfloat16>(pixels_dev, channels, stride, nstride, tables_dev, swap_rb, pa, top, as_stream(stream));
    case NCNN_CUDA_F16: return run_pixels_resize<__half>(pixels_dev, channels, stride, nstride, tables_dev, swap_rb, pa, top, as_stream(stream));
    }
    return -1;
}

} // extern "C"

extern "C" {

int ncnn_cuda_pack_from_planar(const ncnn_cuda_hostmat* src, const ncnn_cuda_tensor* dst, void* stream)
{
    switch (dst->elemtype)
    {
    case NCNN_CUDA_F32: return planar_convert<float, true>(src, dst, as_stream(stream));
    case NCNN_CUDA_BF16: return planar_convert<__nv_bfloat16, true>(src, dst, as_stream(stream));
    case NCNN_CUDA_F16: return planar_convert<__half, true>(src, dst, as_stream(stream));
    }
    return -1;
}

int ncnn_cuda_unpack_to_planar(const ncnn_cuda_tensor* src, const ncnn_cuda_hostmat* dst, void* stream)
{
    switch (src->elemtype)
    {
    case NCNN_CUDA_F32: return planar_convert<float, false>(dst, src, as_stream(stream));
    case NCNN_CUDA_BF16: return planar_convert<__nv_bfloat16, false>(dst, src, as_stream(stream));
    case NCNN_CUDA_F16: return planar_convert<__half, false>(dst, src, as_stream(stream));
    }
    return -1;
}

int ncnn_cuda_pixels_to_blob(const unsigned char* pixels_dev, int channels, int w, int h, int stride, long long nstride, int swap_rb, const float* mean_vals, const float* norm_vals,
                             const ncnn_cuda_tensor* top, void* stream)
{
    NC_REQUIRE(pixels_dev && top && top->dims == 3 && top->w == w && top->h == h && top->c == channels, "pixels_to_blob: the top blob must be (w, h, channels)");
    NC_REQUIRE(channels == 1 || channels == 3 || channels == 4, "pixels_to_blob: 1, 3 or 4 interleaved channels");
    NC_REQUIRE(stride >= w * channels, "pixels_to_blob: stride shorter than a row");
    PixelAffine pa;
    for (int c = 0; c < 4; c++)
    {
        pa.mean[c] = (mean_vals && c < channels) ? mean_vals[c] : 0.f;
        pa.norm[c] = (norm_vals && c < channels) ? norm_vals[c] : 1.f;
    }
    switch (top->elemtype)
    {
    case NCNN_CUDA_F32: return run_pixels<float>(pixels_dev, channels, w, h, stride, nstride, swap_rb, pa, top, as_stream(stream));
    case NCNN_CUDA_BF16: return run_pixels<__nv_bfloat16>(pixels_dev, channels, w, h, stride, nstride, swap_rb, pa, top, as_stream(stream));
    case NCNN_CUDA_F16: return run_pixels<__half>(pixels_dev, channels, w, h, stride, nstride, swap_rb, pa, top, as_stream(stream));
    }
    return -1;
}

int ncnn_cuda_reshape(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, void* stream)
{
    DShape ss = dshape(src), ds = dshape(dst);
    NC_REQUIRE(ss.logical_count() == ds.logical_count(), "reshape: element counts differ");
    NC_REQUIRE((src->n < 1 ? 1 : src->n) == (dst->n < 1 ? 1 : dst->n), "reshape: batch differs");
    PermuteMap pm = {{0, 1, 2, 3}};
    if (src->elemtype == dst->elemtype)
    {
        // (w,h,c) -> (w*h, c): rows = pixels, cols = channels of the source; the 2-D top is [c][w*h]
        if (src->dims == 3 && dst->dims == 2 && dst->w == src->w * src->h && dst->h == src->c)
        {
            int r = launch_transpose2d(src, dst, src->w * src->h, src->c, src->cpitch, dst->cpitch, as_stream(stream));
            if (r <= 0) return r;
        }
        // (w2, h2) -> (w, h, c) with w*h = w2, c = h2: the inverse
        if (src->dims == 2 && dst->dims == 3 && src->w == dst->w * dst->h && src->h == dst->c)
        {
            int r = launch_transpose2d(src, dst, src->h, src->w, src->cpitch, dst->cpitch, as_stream(stream));
            if (r <= 0) return r;
        }
    }
    return dispatch_gather(src, dst, false, pm, as_stream(stream));
}

int ncnn_cuda_permute(const ncnn_cuda_tensor* src, const ncnn_cuda_tensor* dst, int order_type, void* stream)
{
    // tables: new (w,h,[d,]c) named in terms of the old dims, src/layer/permute.cpp:38-41, 66-73, 188-212
    // encoded as src_of[new dim] with 0=w 1=h 2=d 3=c
    static const int t2[2][2] = {{0, 1}, {1, 0}};
    static const int t3[6][3] = {{0, 1, 3}, {1, 0, 3}, {0, 3, 1}, {3, 0, 1}, {1, 3, 0}, {3, 1, 0}};
    static const int t4[24][4] = {
        {0, 1, 2, 3}, {1, 0, 2, 3}, {0, 2, 1, 3}, {2, 0, 1, 3}, {1, 2, 0, 3}, {2, 1, 0, 3},
        {0, 1, 3, 2}, {1, 0, 3, 2}, {0, 3, 1, 2}, {3, 0, 1, 2}, {1, 3, 0, 2}, {3, 1, 0, 2},
        {0, 2, 3, 1}, {2, 0, 3, 1}, {0, 3, 2, 1}, {3, 0, 2, 1}, {2, 3, 0, 1}, {3, 2, 0, 1},
        {1, 2, 3, 0}, {2, 1, 3, 0}, {1, 3, 2, 0}, {3, 1, 2, 0}, {2, 3, 1, 0}, {3, 2, 1, 0}};
    PermuteMap pm = {{0, 1, 2, 3}};
    NC_REQUIRE(src->dims == dst->dims, "permute: rank differs");
    if (src->dims == 2)
    {
        NC_REQUIRE(order_type >= 0 && order_type < 2, "permute: bad order_type");
        if (order_type == 1 && src->elemtype == dst->elemtype && dst->w == src->h && dst->h == src->w)
        {
            // 2-D blobs are [h rows][w innermost]: swapping w and h is a plain transpose
            int r = launch_transpose2d(src, dst, src->h, src->w, src->cpitch, dst->cpitch, as_stream(stream));
            if (r <= 0) return r;
        }
        pm.src_of[0] = t2[order_type][0];
        pm.src_of[1] = t2[order_type][1];
    }
    else if (src->dims == 3)
    {
        NC_REQUIRE(order_type >= 0 && order_type < 6, "permute: bad order_type");
        // 3-D blobs have no d: (w,h,c) -> logical slots 0,1,3
        pm.src_of[0] = t3[order_type][0];
        pm.src_of[1] = t3[order_type][1];
        pm.src_of[2] = 2;
        pm.src_of[3] = t3[order_type][2];
    }
    else if (src->dims == 4)
    {
        NC_REQUIRE(order_type >= 0 && order_type < 24, "permute: bad order_type");
        for (int i = 0; i < 4; i++) pm.src_of[i] = t4[order_type][i];
    }
    return dispatch_gather(src, dst, true, pm, as_stream(stream));
}

} // extern "C"
