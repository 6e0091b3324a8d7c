// mat.h -- host tensor `Mat` and device tensor `CudaMat`.
//
// Mat keeps the reference's layout contract bit for bit (src/mat.h:50-382, src/mat.cpp:299-861): planar, intrusive
// refcount placed after the data, cstep aligned to 16 bytes per channel plane, batch axis n with nstep aligned to
// 4096 bytes (src/mat.cpp:779-780), Mat::batch(b) a zero-copy view (src/mat.h:1786-1802).
//
// CudaMat is the CUDA sibling of VkMat (src/mat.h:387-555): a refcounted handle on a device blob owned by a
// CudaAllocator.  Its layout is PRIVATE to the backend (channel-innermost [n][pixels][cpitch], see
// include/ncnn_cuda.h); host code only sees it through upload/download (command.h).
#ifndef NCNN_B200_MAT_H
#define NCNN_B200_MAT_H

#include <stddef.h>
#include <string.h>

#include "allocator.h"
#include "ncnn_cuda.h"
#include "platform.h"

namespace ncnn {

class NCNN_EXPORT Mat
{
public:
    Mat();
    Mat(int w, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, int c, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, int d, int c, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(const Mat& m);
    // external data (not owned, refcount == 0)
    Mat(int w, void* data, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, void* data, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, int c, void* data, size_t elemsize = 4u, Allocator* allocator = 0);
    Mat(int w, int h, int d, int c, void* data, size_t elemsize = 4u, Allocator* allocator = 0);
    ~Mat();
    Mat& operator=(const Mat& m);

    void fill(float v);
    void fill(int v);
    Mat clone(Allocator* allocator = 0) const;

    // src/mat_pixel.cpp / src/mat.cpp: 8-bit interleaved pixels -> planar fp32 (type = from | (to << 16), codes 1 RGB 2 BGR 3 GRAY
    // 4 RGBA 5 BGRA), and the per-channel (x - mean) * norm pass.  Host versions; the device path is Extractor::input_pixels.
    static Mat from_pixels(const unsigned char* pixels, int type, int w, int h, int stride, Allocator* allocator = 0);
    void substract_mean_normalize(const float* mean_vals, const float* norm_vals);
    // src/mat_pixel.cpp:2692-2753; stride 0 = w * channels of the target type
    void to_pixels(unsigned char* pixels, int type, int stride = 0) const;
    // src/mat_pixel.cpp:2546-2690: the resize / roi forms (resize = the reference's 8-bit bilinear, bit-exact)
    static Mat from_pixels_resize(const unsigned char* pixels, int type, int w, int h, int stride, int target_width, int target_height, Allocator* allocator = 0);
    static Mat from_pixels_roi(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, Allocator* allocator = 0);
    static Mat from_pixels_roi_resize(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, int target_width,
                                      int target_height, Allocator* allocator = 0);
    // src/mat_pixel.cpp:2773-2806
    void to_pixels_resize(unsigned char* pixels, int type, int target_width, int target_height, int target_stride = 0) const;
    void clone_from(const Mat& mat, Allocator* allocator = 0);
    Mat reshape(int w, Allocator* allocator = 0) const;
    Mat reshape(int w, int h, Allocator* allocator = 0) const;
    Mat reshape(int w, int h, int c, Allocator* allocator = 0) const;
    Mat reshape(int w, int h, int d, int c, Allocator* allocator = 0) const;

    void create(int w, size_t elemsize = 4u, Allocator* allocator = 0);
    void create(int w, int h, size_t elemsize = 4u, Allocator* allocator = 0);
    void create(int w, int h, int c, size_t elemsize = 4u, Allocator* allocator = 0);
    void create(int w, int h, int d, int c, size_t elemsize = 4u, Allocator* allocator = 0);
    // with elempack (kept for signature compatibility; this backend only produces elempack 1)
    void create(int w, size_t elemsize, int elempack, Allocator* allocator);
    void create(int w, int h, size_t elemsize, int elempack, Allocator* allocator);
    void create(int w, int h, int c, size_t elemsize, int elempack, Allocator* allocator);
    void create(int w, int h, int d, int c, size_t elemsize, int elempack, Allocator* allocator);
    // batch
    void create(int w, size_t elemsize, int elempack, int n, Allocator* allocator);
    void create(int w, int h, size_t elemsize, int elempack, int n, Allocator* allocator);
    void create(int w, int h, int c, size_t elemsize, int elempack, int n, Allocator* allocator);
    void create(int w, int h, int d, int c, size_t elemsize, int elempack, int n, Allocator* allocator);
    void create_like(const Mat& m, Allocator* allocator = 0);
    void create_like(const Mat& m, int n, Allocator* allocator);
    // any rank, n <= 1 gives the non-batch layout
    void create_dims(int dims, int w, int h, int d, int c, int n, size_t elemsize, Allocator* allocator);

    void addref();
    void release();
    bool empty() const;
    size_t total() const;
    int elembits() const;
    Mat shape() const;

    Mat channel(int c);
    const Mat channel(int c) const;
    Mat batch(int b);
    const Mat batch(int b) const;
    Mat batch_range(int b, int batches);
    const Mat batch_range(int b, int batches) const;
    float* row(int y);
    const float* row(int y) const;
    template<typename T>
    T* row(int y)
    {
        return (T*)((unsigned char*)data + (size_t)w * y * elemsize);
    }
    template<typename T>
    const T* row(int y) const
    {
        return (const T*)((unsigned char*)data + (size_t)w * y * elemsize);
    }
    template<typename T>
    operator T*()
    {
        return (T*)data;
    }
    template<typename T>
    operator const T*() const
    {
        return (const T*)data;
    }
    float& operator[](size_t i)
    {
        return ((float*)data)[i];
    }
    const float& operator[](size_t i) const
    {
        return ((const float*)data)[i];
    }

    void* data;
    int* refcount;
    size_t elemsize;
    int elempack;
    Allocator* allocator;
    int dims;
    int w, h, d, c;
    size_t cstep;
    int n;
    size_t nstep;
};

class CudaMat
{
public:
    CudaMat();
    CudaMat(const CudaMat& m);
    ~CudaMat();
    CudaMat& operator=(const CudaMat& m);

    // elemtype: NCNN_CUDA_F32 / BF16 / F16
    void create(int w, int elemtype, int n, CudaAllocator* allocator);
    void create(int w, int h, int elemtype, int n, CudaAllocator* allocator);
    void create(int w, int h, int c, int elemtype, int n, CudaAllocator* allocator);
    void create(int w, int h, int d, int c, int elemtype, int n, CudaAllocator* allocator);
    void create_dims(int dims, int w, int h, int d, int c, int elemtype, int n, CudaAllocator* allocator);
    void create_like(const CudaMat& m, CudaAllocator* allocator);
    void create_like(const Mat& m, int elemtype, CudaAllocator* allocator);

    void addref();
    void release();
    bool empty() const
    {
        return data == 0 || total_elems() == 0;
    }
    size_t total_elems() const
    {
        return (size_t)(n < 1 ? 1 : n) * nstep;
    }
    int pixels() const;   // P of the device layout
    int channels() const; // C of the device layout
    size_t elemsize() const
    {
        return elemtype == NCNN_CUDA_F32 ? 4u : 2u;
    }
    // same storage, other logical shape (only valid when the device layouts coincide; see Reshape layer)
    ncnn_cuda_tensor view() const;
    // channels [c0, c0 + count) of a 3-D / 4-D blob as a VIEW: same pixels, same cpitch / nstep, data moved by c0 elements; shares
    // the refcount, so the parent allocation lives as long as any view does (Slice without a copy; Concat in place).
    // c0 must keep the view 16-byte aligned.  Empty when the request is out of range.
    CudaMat channel_range(int c0, int count) const;

    void* data;
    void* base;    // the allocation this handle keeps alive (what goes back to the allocator): == data except for views
    int* refcount; // host-side counter
    CudaAllocator* allocator;
    int elemtype;
    int dims;
    int w, h, d, c, n;
    int cpitch;
    size_t nstep;
};

inline ncnn_cuda_hostmat host_view(const Mat& m, void* data_override)
{
    ncnn_cuda_hostmat hm;
    hm.data = data_override;
    hm.dims = m.dims;
    hm.w = m.w;
    hm.h = m.h;
    hm.d = m.d;
    hm.c = m.c;
    hm.n = m.n < 1 ? 1 : m.n;
    hm.cstep = (long long)m.cstep;
    hm.nstep = (long long)m.nstep;
    return hm;
}

} // namespace ncnn

#endif // NCNN_B200_MAT_H
