// mat.cpp -- see mat.h
#include "mat.h"

#include <vector>

namespace ncnn {

Mat::Mat()
    : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0), n(1), nstep(0)
{
}

Mat::Mat(int _w, size_t _elemsize, Allocator* _allocator)
    : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0), n(1), nstep(0)
{
    create(_w, _elemsize, _allocator);
}

Mat::Mat(int _w, int _h, size_t _elemsize, Allocator* _allocator)
    : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0), n(1), nstep(0)
{
    create(_w, _h, _elemsize, _allocator);
}

Mat::Mat(int _w, int _h, int _c, size_t _elemsize, Allocator* _allocator)
    : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0), n(1), nstep(0)
{
    create(_w, _h, _c, _elemsize, _allocator);
}

Mat::Mat(int _w, int _h, int _d, int _c, size_t _elemsize, Allocator* _allocator)
    : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0), n(1), nstep(0)
{
    create(_w, _h, _d, _c, _elemsize, _allocator);
}

Mat::Mat(const Mat& m)
    : data(m.data), refcount(m.refcount), elemsize(m.elemsize), elempack(m.elempack), allocator(m.allocator), dims(m.dims), w(m.w), h(m.h), d(m.d), c(m.c),
      cstep(m.cstep), n(m.n), nstep(m.nstep)
{
    addref();
}

Mat::Mat(int _w, void* _data, size_t _elemsize, Allocator* _allocator)
    : data(_data), refcount(0), elemsize(_elemsize), elempack(1), allocator(_allocator), dims(1), w(_w), h(1), d(1), c(1), n(1)
{
    cstep = alignSize((size_t)w * elemsize, 16) / elemsize;
    nstep = cstep;
}

Mat::Mat(int _w, int _h, void* _data, size_t _elemsize, Allocator* _allocator)
    : data(_data), refcount(0), elemsize(_elemsize), elempack(1), allocator(_allocator), dims(2), w(_w), h(_h), d(1), c(1), n(1)
{
    cstep = alignSize((size_t)w * h * elemsize, 16) / elemsize;
    nstep = cstep;
}

Mat::Mat(int _w, int _h, int _c, void* _data, size_t _elemsize, Allocator* _allocator)
    : data(_data), refcount(0), elemsize(_elemsize), elempack(1), allocator(_allocator), dims(3), w(_w), h(_h), d(1), c(_c), n(1)
{
    cstep = alignSize((size_t)w * h * elemsize, 16) / elemsize;
    nstep = cstep * c;
}

Mat::Mat(int _w, int _h, int _d, int _c, void* _data, size_t _elemsize, Allocator* _allocator)
    : data(_data), refcount(0), elemsize(_elemsize), elempack(1), allocator(_allocator), dims(4), w(_w), h(_h), d(_d), c(_c), n(1)
{
    cstep = alignSize((size_t)w * h * d * elemsize, 16) / elemsize;
    nstep = cstep * c;
}

Mat::~Mat()
{
    release();
}

Mat& Mat::operator=(const Mat& m)
{
    if (this == &m) return *this;
    if (m.refcount) NCNN_XADD(m.refcount, 1);
    release();
    data = m.data;
    refcount = m.refcount;
    elemsize = m.elemsize;
    elempack = m.elempack;
    allocator = m.allocator;
    dims = m.dims;
    w = m.w;
    h = m.h;
    d = m.d;
    c = m.c;
    cstep = m.cstep;
    n = m.n;
    nstep = m.nstep;
    return *this;
}

void Mat::addref()
{
    if (refcount) NCNN_XADD(refcount, 1);
}

void Mat::release()
{
    if (refcount && NCNN_XADD(refcount, -1) == 1)
    {
        if (allocator)
            allocator->fastFree(data);
        else
            fastFree(data);
    }
    data = 0;
    elemsize = 0;
    elempack = 0;
    dims = 0;
    w = h = d = c = 0;
    cstep = 0;
    n = 1;
    nstep = 0;
    refcount = 0;
}

bool Mat::empty() const
{
    return data == 0 || total() == 0;
}

size_t Mat::total() const
{
    return cstep * c;
}

int Mat::elembits() const
{
    return elempack ? (int)(elemsize * 8) / elempack : 0;
}

Mat Mat::shape() const
{
    if (dims == 1) return Mat(w * elempack, (void*)0);
    if (dims == 2) return Mat(w, h * elempack, (void*)0);
    if (dims == 3) return Mat(w, h, c * elempack, (void*)0);
    if (dims == 4) return Mat(w, h, d, c * elempack, (void*)0);
    return Mat();
}

void Mat::create_dims(int _dims, int _w, int _h, int _d, int _c, int _n, size_t _elemsize, Allocator* _allocator)
{
    if (_dims < 2) _h = 1;
    if (_dims < 4) _d = 1;
    if (_dims < 3) _c = 1;
    if (_n < 1) _n = 1;
    if (dims == _dims && w == _w && h == _h && d == _d && c == _c && n == _n && elemsize == _elemsize && elempack == 1 && allocator == _allocator && data) return;
    release();
    elemsize = _elemsize;
    elempack = 1;
    allocator = _allocator;
    dims = _dims;
    w = _w;
    h = _h;
    d = _d;
    c = _c;
    n = _n;
    cstep = alignSize((size_t)w * h * d * elemsize, 16) / elemsize;
    if (n > 1)
        nstep = alignSize(cstep * c * elemsize, 4096) / elemsize; // src/mat.cpp:780
    else
        nstep = cstep * c;
    size_t totalsize = alignSize(nstep * n * elemsize, 4);
    if (totalsize > 0)
    {
        if (allocator)
            data = allocator->fastMalloc(totalsize + (int)sizeof(*refcount));
        else
            data = fastMalloc(totalsize + (int)sizeof(*refcount));
    }
    if (data)
    {
        refcount = (int*)(((unsigned char*)data) + totalsize);
        *refcount = 1;
    }
}

void Mat::create(int _w, size_t _elemsize, Allocator* _allocator)
{
    create_dims(1, _w, 1, 1, 1, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, size_t _elemsize, Allocator* _allocator)
{
    create_dims(2, _w, _h, 1, 1, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _c, size_t _elemsize, Allocator* _allocator)
{
    create_dims(3, _w, _h, 1, _c, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _d, int _c, size_t _elemsize, Allocator* _allocator)
{
    create_dims(4, _w, _h, _d, _c, 1, _elemsize, _allocator);
}
void Mat::create(int _w, size_t _elemsize, int, Allocator* _allocator)
{
    create_dims(1, _w, 1, 1, 1, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, size_t _elemsize, int, Allocator* _allocator)
{
    create_dims(2, _w, _h, 1, 1, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _c, size_t _elemsize, int, Allocator* _allocator)
{
    create_dims(3, _w, _h, 1, _c, 1, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _d, int _c, size_t _elemsize, int, Allocator* _allocator)
{
    create_dims(4, _w, _h, _d, _c, 1, _elemsize, _allocator);
}
void Mat::create(int _w, size_t _elemsize, int, int _n, Allocator* _allocator)
{
    create_dims(1, _w, 1, 1, 1, _n, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, size_t _elemsize, int, int _n, Allocator* _allocator)
{
    create_dims(2, _w, _h, 1, 1, _n, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _c, size_t _elemsize, int, int _n, Allocator* _allocator)
{
    create_dims(3, _w, _h, 1, _c, _n, _elemsize, _allocator);
}
void Mat::create(int _w, int _h, int _d, int _c, size_t _elemsize, int, int _n, Allocator* _allocator)
{
    create_dims(4, _w, _h, _d, _c, _n, _elemsize, _allocator);
}

void Mat::create_like(const Mat& m, Allocator* _allocator)
{
    create_dims(m.dims, m.w, m.h, m.d, m.c, m.n, m.elemsize, _allocator);
}

void Mat::create_like(const Mat& m, int _n, Allocator* _allocator)
{
    create_dims(m.dims, m.w, m.h, m.d, m.c, _n, m.elemsize, _allocator);
}

void Mat::fill(float v)
{
    size_t count = (size_t)(n < 1 ? 1 : n) * nstep;
    if (n <= 1) count = total();
    float* p = (float*)data;
    for (size_t i = 0; i < count; i++) p[i] = v;
}

void Mat::fill(int v)
{
    size_t count = n <= 1 ? total() : (size_t)n * nstep;
    int* p = (int*)data;
    for (size_t i = 0; i < count; i++) p[i] = v;
}

// Mat::from_pixels (src/mat_pixel.cpp:2440-2545): every conversion of the reference's table, expressed as one rule per output
// channel: copy source byte k, the 8-bit luma (77 R + 150 G + 29 B) >> 8 (from_rgb2gray, :736-800), or the constant 255 (alpha).
Mat Mat::from_pixels(const unsigned char* pixels, int type, int w, int h, int stride, Allocator* allocator)
{
    Mat m;
    const int from = type & 0xffff;
    int to = (type >> 16) & 0xffff;
    if (to == 0) to = from;
    const int in_ch = from == 3 ? 1 : (from == 4 || from == 5 ? 4 : (from == 1 || from == 2 ? 3 : 0));
    const int out_ch = to == 3 ? 1 : (to == 4 || to == 5 ? 4 : (to == 1 || to == 2 ? 3 : 0));
    if (!pixels || in_ch == 0 || out_ch == 0 || w <= 0 || h <= 0)
    {
        NCNN_LOGE("from_pixels: unsupported pixel type %d", type);
        return m;
    }
    if (stride <= 0) stride = w * in_ch;
    // position of R, G, B (and A) inside a source pixel; gray sources have a single byte
    const bool from_bgr = from == 2 || from == 5;
    const int sr = from_bgr ? 2 : 0, sg = 1, sb = from_bgr ? 0 : 2;
    // rule per output channel: >= 0 copy that source byte, -1 luma, -2 constant 255
    int rule[4] = {0, 0, 0, 0};
    if (to == 3)
        rule[0] = from == 3 ? 0 : -1;
    else
    {
        const bool to_bgr = to == 2 || to == 5;
        if (from == 3)
            rule[0] = rule[1] = rule[2] = 0; // gray replicated
        else
        {
            rule[0] = to_bgr ? sb : sr;
            rule[1] = sg;
            rule[2] = to_bgr ? sr : sb;
        }
        if (out_ch == 4) rule[3] = in_ch == 4 ? 3 : -2;
    }
    m.create(w, h, out_ch, (size_t)4u, allocator);
    if (m.empty()) return m;
    for (int c = 0; c < out_ch; c++)
    {
        float* dst = m.channel(c);
        const int r = rule[c];
        for (int y = 0; y < h; y++)
        {
            const unsigned char* row = pixels + (size_t)y * stride;
            float* d = dst + (size_t)y * w;
            if (r >= 0)
                for (int x = 0; x < w; x++) d[x] = (float)row[x * in_ch + r];
            else if (r == -1)
                for (int x = 0; x < w; x++)
                {
                    const unsigned char* px = row + x * in_ch;
                    d[x] = (float)((px[sr] * 77 + px[sg] * 150 + px[sb] * 29) >> 8);
                }
            else
                for (int x = 0; x < w; x++) d[x] = 255.f;
        }
    }
    return m;
}

// Mat::to_pixels (src/mat_pixel.cpp:2710-2753): the eight conversion families of the reference as one rule per output byte:
// the Mat channel to take (saturated as SATURATE_CAST_UCHAR does, :147: truncate toward zero, clamp to 0..255) or the
// constant 255 for a synthesised alpha.  Conversions the reference does not implement (e.g. RGB2GRAY) are refused the same way.
void Mat::to_pixels(unsigned char* pixels, int type, int stride) const
{
    if (empty() || !pixels || dims != 3) return;
    const int from = type & 0xffff;
    int to = (type >> 16) & 0xffff;
    if (to == 0) to = from;
    const int in_ch = from == 3 ? 1 : (from == 4 || from == 5 ? 4 : (from == 1 || from == 2 ? 3 : 0));
    const int out_ch = to == 3 ? 1 : (to == 4 || to == 5 ? 4 : (to == 1 || to == 2 ? 3 : 0));
    const bool swap = ((from == 1 || from == 4) && (to == 2 || to == 5)) || ((from == 2 || from == 5) && (to == 1 || to == 4));
    // implemented families: same type; RGB<->BGR; RGB/BGR -> RGBA/BGRA (either order); GRAY -> RGBA/BGRA; RGBA<->BGRA
    const bool ok = in_ch != 0 && out_ch != 0 && c >= in_ch
                    && (to == from || (in_ch == 3 && out_ch == 3) || (in_ch == 3 && out_ch == 4) || (in_ch == 1 && out_ch == 4) || (in_ch == 4 && out_ch == 4));
    if (!ok)
    {
        NCNN_LOGE("unimplemented convert type %d", type);
        return;
    }
    int rule[4] = {0, 1, 2, 3}; // -2: constant 255
    if (in_ch == 1)
        rule[0] = rule[1] = rule[2] = 0;
    else if (swap)
    {
        rule[0] = 2;
        rule[2] = 0;
    }
    if (out_ch == 4 && in_ch != 4) rule[3] = -2;
    if (stride <= 0) stride = w * out_ch;
    for (int k = 0; k < out_ch; k++)
    {
        const int r = rule[k];
        const float* src = r >= 0 ? (const float*)channel(r) : 0;
        for (int y = 0; y < h; y++)
        {
            unsigned char* row = pixels + (size_t)y * stride + k;
            if (!src)
            {
                for (int x = 0; x < w; x++) row[x * out_ch] = 255;
                continue;
            }
            const float* sp = src + (size_t)y * w;
            for (int x = 0; x < w; x++)
            {
                int v = (int)sp[x];
                row[x * out_ch] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}

// resize_bilinear_c1/c3/c4 of the reference (src/mat_pixel_resize.cpp:210-1039) on the host: the same offset / coefficient tables
// the device kernel uses (ncnn_cuda_resize_tables, computed as :599-669) and the same integer formula per output byte.
static int resize_bilinear_u8(const unsigned char* src, int ch, int srcw, int srch, int srcstride, unsigned char* dst, int w, int h, int stride)
{
    std::vector<int> tab((size_t)ncnn_cuda_resize_tables_count(w, h));
    int ret = ncnn_cuda_resize_tables(srcw, srch, w, h, tab.data());
    if (ret != 0) return ret;
    const int* xofs = tab.data();
    const int* yofs = xofs + w;
    const int* alpha = yofs + h;
    const int* beta = alpha + 2 * w;
    std::vector<short> rows0((size_t)w * ch), rows1((size_t)w * ch);
    int prev_sy = -2;
    for (int y = 0; y < h; y++)
    {
        const int sy = yofs[y];
        if (sy != prev_sy)
        {
            const unsigned char* s0 = src + (size_t)sy * srcstride;
            const unsigned char* s1 = s0 + srcstride;
            for (int x = 0; x < w; x++)
            {
                const int a0 = alpha[2 * x], a1 = alpha[2 * x + 1];
                const unsigned char* p0 = s0 + xofs[x] * ch;
                const unsigned char* p1 = s1 + xofs[x] * ch;
                for (int c = 0; c < ch; c++)
                {
                    rows0[(size_t)x * ch + c] = (short)((p0[c] * a0 + p0[c + ch] * a1) >> 4);
                    rows1[(size_t)x * ch + c] = (short)((p1[c] * a0 + p1[c + ch] * a1) >> 4);
                }
            }
            prev_sy = sy;
        }
        const int b0 = beta[2 * y], b1 = beta[2 * y + 1];
        unsigned char* d = dst + (size_t)y * stride;
        for (int i = 0; i < w * ch; i++)
        {
            int q = (((b0 * rows0[i]) >> 16) + ((b1 * rows1[i]) >> 16) + 2) >> 2;
            d[i] = (unsigned char)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
    }
    return 0;
}

static int pixel_channels(int fmt)
{
    return fmt == 3 ? 1 : (fmt == 4 || fmt == 5 ? 4 : (fmt == 1 || fmt == 2 ? 3 : 0));
}

// src/mat_pixel.cpp:2546-2578
Mat Mat::from_pixels_resize(const unsigned char* pixels, int type, int w, int h, int stride, int target_width, int target_height, Allocator* allocator)
{
    if (w == target_width && h == target_height) return from_pixels(pixels, type, w, h, stride, allocator);
    const int ch = pixel_channels(type & 0xffff);
    if (!pixels || ch == 0 || target_width <= 0 || target_height <= 0)
    {
        NCNN_LOGE("unknown convert type %d", type);
        return Mat();
    }
    if (stride <= 0) stride = w * ch;
    std::vector<unsigned char> resized((size_t)target_width * target_height * ch);
    if (resize_bilinear_u8(pixels, ch, w, h, stride, resized.data(), target_width, target_height, target_width * ch) != 0) return Mat();
    return from_pixels(resized.data(), type, target_width, target_height, target_width * ch, allocator);
}

// src/mat_pixel.cpp:2608-2632
Mat Mat::from_pixels_roi(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, Allocator* allocator)
{
    const int ch = pixel_channels(type & 0xffff);
    if (roix < 0 || roiy < 0 || roiw <= 0 || roih <= 0 || roix + roiw > w || roiy + roih > h)
    {
        NCNN_LOGE("roi %d %d %d %d out of image %d %d", roix, roiy, roiw, roih, w, h);
        return Mat();
    }
    if (ch == 0) return Mat();
    if (stride <= 0) stride = w * ch;
    return from_pixels(pixels + (size_t)roiy * stride + (size_t)roix * ch, type, roiw, roih, stride, allocator);
}

// src/mat_pixel.cpp:2664-2690
Mat Mat::from_pixels_roi_resize(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, int target_width, int target_height,
                                Allocator* allocator)
{
    const int ch = pixel_channels(type & 0xffff);
    if (roix < 0 || roiy < 0 || roiw <= 0 || roih <= 0 || roix + roiw > w || roiy + roih > h)
    {
        NCNN_LOGE("roi %d %d %d %d out of image %d %d", roix, roiy, roiw, roih, w, h);
        return Mat();
    }
    if (ch == 0) return Mat();
    if (stride <= 0) stride = w * ch;
    return from_pixels_resize(pixels + (size_t)roiy * stride + (size_t)roix * ch, type, roiw, roih, stride, target_width, target_height, allocator);
}

// src/mat_pixel.cpp:2773-2806 (an equal target size writes tightly packed rows, as the reference does)
void Mat::to_pixels_resize(unsigned char* pixels, int type, int target_width, int target_height, int target_stride) const
{
    if (w == target_width && h == target_height) return to_pixels(pixels, type);
    int to = (type >> 16) & 0xffff;
    if (to == 0) to = type & 0xffff;
    const int ch = pixel_channels(to);
    if (empty() || !pixels || ch == 0 || target_width <= 0 || target_height <= 0) return;
    if (target_stride <= 0) target_stride = target_width * ch;
    std::vector<unsigned char> full((size_t)w * h * ch);
    to_pixels(full.data(), type, w * ch);
    resize_bilinear_u8(full.data(), ch, w, h, w * ch, pixels, target_width, target_height, target_stride);
}

// Mat::substract_mean_normalize (src/mat.cpp): (x - mean[c]) * norm[c]; either array may be NULL
void Mat::substract_mean_normalize(const float* mean_vals, const float* norm_vals)
{
    if (empty() || (!mean_vals && !norm_vals)) return;
    const int chs = dims == 1 ? 1 : (dims == 2 ? 1 : c);
    const size_t size = dims == 1 ? (size_t)w : (dims == 2 ? (size_t)w * h : (size_t)w * h * d);
    for (int q = 0; q < chs; q++)
    {
        float* ptr = dims >= 3 ? (float*)channel(q) : (float*)data;
        const float mean = mean_vals ? mean_vals[q] : 0.f;
        const float norm = norm_vals ? norm_vals[q] : 1.f;
        if (mean_vals && norm_vals)
            for (size_t i = 0; i < size; i++) ptr[i] = (ptr[i] - mean) * norm;
        else if (mean_vals)
            for (size_t i = 0; i < size; i++) ptr[i] -= mean;
        else
            for (size_t i = 0; i < size; i++) ptr[i] *= norm;
    }
}

Mat Mat::clone(Allocator* _allocator) const
{
    if (empty()) return Mat();
    Mat m;
    m.create_dims(dims, w, h, d, c, n, elemsize, _allocator);
    if (m.empty()) return m;
    size_t bytes = (n <= 1 ? total() : (size_t)n * nstep) * elemsize;
    memcpy(m.data, data, bytes);
    return m;
}

void Mat::clone_from(const Mat& mat, Allocator* _allocator)
{
    *this = mat.clone(_allocator);
}

// reshape keeps the logical element order (c, d, h, w); planes are re-aligned when cstep padding differs
static void copy_logical(const Mat& src, Mat& dst)
{
    // walk both in logical order
    size_t total = (size_t)src.w * src.h * src.d * src.c;
    size_t splane = (size_t)src.w * src.h * src.d, dplane = (size_t)dst.w * dst.h * dst.d;
    const unsigned char* s = (const unsigned char*)src.data;
    unsigned char* t = (unsigned char*)dst.data;
    size_t es = src.elemsize;
    size_t i = 0;
    while (i < total)
    {
        size_t sq = i / splane, so = i % splane;
        size_t dq = i / dplane, dof = i % dplane;
        size_t run = splane - so;
        if (dplane - dof < run) run = dplane - dof;
        memcpy(t + (dq * dst.cstep + dof) * es, s + (sq * src.cstep + so) * es, run * es);
        i += run;
    }
}

Mat Mat::reshape(int _w, Allocator* _allocator) const
{
    if ((size_t)w * h * d * c != (size_t)_w) return Mat();
    Mat m;
    m.create(_w, elemsize, _allocator);
    if (!m.empty()) copy_logical(*this, m);
    return m;
}

Mat Mat::reshape(int _w, int _h, Allocator* _allocator) const
{
    if ((size_t)w * h * d * c != (size_t)_w * _h) return Mat();
    Mat m;
    m.create(_w, _h, elemsize, _allocator);
    if (!m.empty()) copy_logical(*this, m);
    return m;
}

Mat Mat::reshape(int _w, int _h, int _c, Allocator* _allocator) const
{
    if ((size_t)w * h * d * c != (size_t)_w * _h * _c) return Mat();
    Mat m;
    m.create(_w, _h, _c, elemsize, _allocator);
    if (!m.empty()) copy_logical(*this, m);
    return m;
}

Mat Mat::reshape(int _w, int _h, int _d, int _c, Allocator* _allocator) const
{
    if ((size_t)w * h * d * c != (size_t)_w * _h * _d * _c) return Mat();
    Mat m;
    m.create(_w, _h, _d, _c, elemsize, _allocator);
    if (!m.empty()) copy_logical(*this, m);
    return m;
}

Mat Mat::channel(int _c)
{
    Mat m(w, h, d, (unsigned char*)data + cstep * _c * elemsize, elemsize, allocator);
    m.dims = dims - 1;
    if (dims == 4) m.cstep = (size_t)w * h;
    return m;
}

const Mat Mat::channel(int _c) const
{
    Mat m(w, h, d, (unsigned char*)data + cstep * _c * elemsize, elemsize, allocator);
    m.dims = dims - 1;
    if (dims == 4) m.cstep = (size_t)w * h;
    return m;
}

Mat Mat::batch(int b)
{
    return ((const Mat*)this)->batch(b);
}

const Mat Mat::batch(int b) const
{
    Mat m;
    m.data = (unsigned char*)data + nstep * b * elemsize;
    m.refcount = 0;
    m.elemsize = elemsize;
    m.elempack = elempack;
    m.allocator = allocator;
    m.dims = dims;
    m.w = w;
    m.h = h;
    m.d = d;
    m.c = c;
    m.cstep = cstep;
    m.n = 1;
    m.nstep = cstep * c;
    return m;
}

Mat Mat::batch_range(int b, int batches)
{
    return ((const Mat*)this)->batch_range(b, batches);
}

const Mat Mat::batch_range(int b, int batches) const
{
    Mat m = batch(b);
    m.n = batches;
    m.nstep = nstep;
    return m;
}

float* Mat::row(int y)
{
    return (float*)((unsigned char*)data + (size_t)w * y * elemsize);
}

const float* Mat::row(int y) const
{
    return (const float*)((unsigned char*)data + (size_t)w * y * elemsize);
}

// ------------------------------------------------------------------ CudaMat
CudaMat::CudaMat()
    : data(0), base(0), refcount(0), allocator(0), elemtype(NCNN_CUDA_F32), dims(0), w(0), h(0), d(0), c(0), n(1), cpitch(0), nstep(0)
{
}

CudaMat::CudaMat(const CudaMat& m)
    : data(m.data), base(m.base), refcount(m.refcount), allocator(m.allocator), elemtype(m.elemtype), dims(m.dims), w(m.w), h(m.h), d(m.d), c(m.c), n(m.n), cpitch(m.cpitch),
      nstep(m.nstep)
{
    addref();
}

CudaMat::~CudaMat()
{
    release();
}

CudaMat& CudaMat::operator=(const CudaMat& m)
{
    if (this == &m) return *this;
    if (m.refcount) NCNN_XADD(m.refcount, 1);
    release();
    data = m.data;
    base = m.base;
    refcount = m.refcount;
    allocator = m.allocator;
    elemtype = m.elemtype;
    dims = m.dims;
    w = m.w;
    h = m.h;
    d = m.d;
    c = m.c;
    n = m.n;
    cpitch = m.cpitch;
    nstep = m.nstep;
    return *this;
}

void CudaMat::addref()
{
    if (refcount) NCNN_XADD(refcount, 1);
}

void CudaMat::release()
{
    if (refcount && NCNN_XADD(refcount, -1) == 1)
    {
        if (allocator && base) allocator->fastFree(base);
        delete refcount;
    }
    data = 0;
    base = 0;
    refcount = 0;
    dims = 0;
    w = h = d = c = 0;
    n = 1;
    cpitch = 0;
    nstep = 0;
}

int CudaMat::pixels() const
{
    if (dims == 1) return 1;
    if (dims == 2) return h;
    if (dims == 3) return h * w;
    return d * h * w;
}

int CudaMat::channels() const
{
    return dims <= 2 ? w : c;
}

void CudaMat::create_dims(int _dims, int _w, int _h, int _d, int _c, int _elemtype, int _n, CudaAllocator* _allocator)
{
    if (_dims < 2) _h = 1;
    if (_dims < 4) _d = 1;
    if (_dims < 3) _c = 1;
    if (_n < 1) _n = 1;
    release();
    allocator = _allocator;
    elemtype = _elemtype;
    dims = _dims;
    w = _w;
    h = _h;
    d = _d;
    c = _c;
    n = _n;
    const int vec = _elemtype == NCNN_CUDA_F32 ? 4 : 8; // 16-byte channel vectors; TMA needs 16-byte pixel strides
    const int C = channels();
    cpitch = (C + vec - 1) / vec * vec;
    nstep = (size_t)pixels() * cpitch;
    size_t bytes = nstep * n * elemsize();
    if (bytes == 0 || !allocator) return;
    // a placement allocator may answer with a view into a buffer it manages (it fills data / base / refcount / cpitch / nstep and
    // the owning allocator); otherwise the blob is a fresh allocation owned by the real allocator behind it
    if (allocator->place(*this)) return;
    allocator = allocator->real();
    data = allocator->fastMalloc(bytes);
    base = data;
    if (data)
    {
        refcount = new int;
        *refcount = 1;
    }
}

CudaMat CudaMat::channel_range(int c0, int count) const
{
    CudaMat v;
    if (!data || dims < 3 || c0 < 0 || count <= 0 || c0 + count > c) return v;
    if (((size_t)c0 * elemsize()) % 16 != 0) return v;
    v = *this; // shares the refcount and the base allocation
    v.data = (unsigned char*)data + (size_t)c0 * elemsize();
    v.c = count;
    return v;
}

void CudaMat::create(int _w, int _elemtype, int _n, CudaAllocator* _allocator)
{
    create_dims(1, _w, 1, 1, 1, _elemtype, _n, _allocator);
}
void CudaMat::create(int _w, int _h, int _elemtype, int _n, CudaAllocator* _allocator)
{
    create_dims(2, _w, _h, 1, 1, _elemtype, _n, _allocator);
}
void CudaMat::create(int _w, int _h, int _c, int _elemtype, int _n, CudaAllocator* _allocator)
{
    create_dims(3, _w, _h, 1, _c, _elemtype, _n, _allocator);
}
void CudaMat::create(int _w, int _h, int _d, int _c, int _elemtype, int _n, CudaAllocator* _allocator)
{
    create_dims(4, _w, _h, _d, _c, _elemtype, _n, _allocator);
}
void CudaMat::create_like(const CudaMat& m, CudaAllocator* _allocator)
{
    create_dims(m.dims, m.w, m.h, m.d, m.c, m.elemtype, m.n, _allocator);
}
void CudaMat::create_like(const Mat& m, int _elemtype, CudaAllocator* _allocator)
{
    create_dims(m.dims, m.w, m.h, m.d, m.c, _elemtype, m.n, _allocator);
}

ncnn_cuda_tensor CudaMat::view() const
{
    ncnn_cuda_tensor t;
    t.data = data;
    t.dims = dims;
    t.w = w;
    t.h = h;
    t.d = d;
    t.c = c;
    t.n = n < 1 ? 1 : n;
    t.elemtype = elemtype;
    t.cpitch = cpitch;
    t.nstep = (long long)nstep;
    return t;
}

} // namespace ncnn
