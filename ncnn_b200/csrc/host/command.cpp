#include "command.h"

#include <map>
#include <mutex>

namespace ncnn {

CudaContext::CudaContext(int _device_index)
    : device_index(_device_index), stream(0), blob_allocator(0), staging_allocator(0)
{
    ncnn_cuda_stream_create(&stream);
    blob_allocator = new CudaBlobAllocator(_device_index);
    staging_allocator = new CudaStagingAllocator;
}

CudaContext::~CudaContext()
{
    if (stream)
    {
        ncnn_cuda_stream_sync(stream);
        ncnn_cuda_stream_destroy(stream);
    }
    delete blob_allocator;
    delete staging_allocator;
}

namespace {
struct ContextPool
{
    std::mutex lock;
    std::map<int, std::vector<CudaContext*> > idle;
    std::map<int, CudaWeightAllocator*> weights;
    ~ContextPool()
    {
        // process teardown: the CUDA runtime may already be gone, leave device memory to the driver
    }
};
static ContextPool& pool()
{
    static ContextPool* p = new ContextPool;
    return *p;
}
} // namespace

int get_cuda_device_count()
{
    return ncnn_cuda_device_count();
}

CudaContext* acquire_cuda_context(int device_index)
{
    if (device_index < 0) device_index = ncnn_cuda_get_device();
    if (device_index < 0) return 0;
    if (ncnn_cuda_set_device(device_index) != 0) return 0;
    ContextPool& p = pool();
    {
        std::lock_guard<std::mutex> lk(p.lock);
        std::vector<CudaContext*>& v = p.idle[device_index];
        if (!v.empty())
        {
            CudaContext* c = v.back();
            v.pop_back();
            return c;
        }
    }
    CudaContext* c = new CudaContext(device_index);
    if (!c->stream)
    {
        delete c;
        return 0;
    }
    return c;
}

void reclaim_cuda_context(CudaContext* ctx)
{
    if (!ctx) return;
    ContextPool& p = pool();
    std::lock_guard<std::mutex> lk(p.lock);
    p.idle[ctx->device_index].push_back(ctx);
}

CudaWeightAllocator* get_cuda_weight_allocator(int device_index)
{
    if (device_index < 0) device_index = ncnn_cuda_get_device();
    ContextPool& p = pool();
    std::lock_guard<std::mutex> lk(p.lock);
    CudaWeightAllocator*& a = p.weights[device_index];
    if (!a) a = new CudaWeightAllocator(device_index);
    return a;
}

CudaCompute::CudaCompute(CudaContext* ctx)
    : h2d_bytes(0), d2h_bytes(0), ctx_(ctx), profiling_(false)
{
}

void CudaCompute::profile_begin(int layer_index)
{
    LayerTiming t;
    memset(&t, 0, sizeof(t));
    t.layer_index = layer_index;
    ncnn_cuda_event_create(&t.start);
    ncnn_cuda_event_create(&t.stop);
    ncnn_cuda_event_record(t.start, stream());
    timings_.push_back(t);
}

void CudaCompute::profile_end(const CudaMat& top)
{
    if (timings_.empty()) return;
    LayerTiming& t = timings_.back();
    ncnn_cuda_event_record(t.stop, stream());
    t.dims = top.dims;
    t.w = top.w;
    t.h = top.h;
    t.d = top.d;
    t.c = top.c;
    t.n = top.n;
}

void CudaCompute::clear_timings()
{
    for (size_t i = 0; i < timings_.size(); i++)
    {
        if (timings_[i].start) ncnn_cuda_event_destroy(timings_[i].start);
        if (timings_[i].stop) ncnn_cuda_event_destroy(timings_[i].stop);
    }
    timings_.clear();
}

CudaCompute::~CudaCompute()
{
    if (!downloads_.empty() || !staging_in_flight_.empty() || !keep_alive_.empty()) submit_and_wait();
    clear_timings();
}

static size_t mat_bytes(const Mat& m)
{
    return (m.n <= 1 ? m.total() : (size_t)m.n * m.nstep) * m.elemsize;
}

int CudaCompute::record_upload(const Mat& src, CudaMat& dst, const Option& opt)
{
    if (src.empty()) return -100;
    if (src.elemsize != 4u || src.elempack != 1)
    {
        NCNN_LOGE("record_upload: only fp32 elempack=1 host Mats are accepted (got elemsize %zu elempack %d)", src.elemsize, src.elempack);
        return -1;
    }
    void* st = stream();
    const size_t bytes = mat_bytes(src);
    // raw planar bytes on the device, then one conversion kernel into the backend layout
    CudaMat raw;
    raw.create((int)((bytes + 3) / 4), NCNN_CUDA_F32, 1, workspace_allocator(opt));
    if (raw.empty()) return -100;
    const void* hsrc = src.data;
    if (!ncnn_cuda_host_is_pinned(src.data))
    {
        Allocator* sa = ctx_->staging_allocator;
        void* staging = sa->fastMalloc(bytes);
        if (!staging) return -100;
        memcpy(staging, src.data, bytes);
        staging_in_flight_.push_back(staging);
        hsrc = staging;
    }
    int ret = ncnn_cuda_memcpy_h2d_async(raw.data, hsrc, bytes, st);
    if (ret != 0) return ret;
    h2d_bytes += bytes;
    dst.create_like(src, opt.cuda_elemtype(), blob_allocator(opt));
    if (dst.empty()) return -100;
    ncnn_cuda_hostmat hm = host_view(src, raw.data);
    ncnn_cuda_tensor t = dst.view();
    ret = ncnn_cuda_pack_from_planar(&hm, &t, st);
    keep_alive_.push_back(raw); // returned to the pool at the next submit; reuse before that would still be stream-ordered
    return ret;
}

// pixel type codes of the reference (src/mat.h:213-262): PIXEL_RGB 1, BGR 2, GRAY 3, RGBA 4, BGRA 5; conversion = from | (to << 16)
int CudaCompute::record_upload_pixels(const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, const float* mean_vals, const float* norm_vals,
                                      CudaMat& dst, const Option& opt, int target_w, int target_h)
{
    if (!pixels || w <= 0 || h <= 0 || n <= 0) return -100;
    // Mat::from_pixels_resize (src/mat_pixel.cpp:2546-2549): an equal target size is a plain from_pixels
    const bool resize = target_w > 0 && target_h > 0 && (target_w != w || target_h != h);
    if (resize && (w < 2 || h < 2)) return -1;
    const int from = type & 0xffff, to = (type >> 16) & 0xffff;
    int channels;
    switch (from)
    {
    case 1:
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4:
    case 5: channels = 4; break;
    default: NCNN_LOGE("record_upload_pixels: unknown pixel type %d", type); return -1;
    }
    int swap_rb = 0;
    if (to != 0 && to != from)
    {
        const bool rgb_bgr = (from == 1 && to == 2) || (from == 2 && to == 1) || (from == 4 && to == 5) || (from == 5 && to == 4);
        if (!rgb_bgr)
        {
            NCNN_LOGE("record_upload_pixels: pixel conversion %d -> %d is not on the device path", from, to);
            return -1;
        }
        swap_rb = 1;
    }
    if (stride <= 0) stride = w * channels;
    if (nstride == 0) nstride = (size_t)h * stride;
    void* st = stream();
    const size_t bytes = (size_t)(n - 1) * nstride + (size_t)h * stride;
    CudaMat raw;
    raw.create((int)((bytes + 3) / 4), NCNN_CUDA_F32, 1, workspace_allocator(opt));
    if (raw.empty()) return -100;
    const void* hsrc = pixels;
    if (!ncnn_cuda_host_is_pinned(pixels))
    {
        Allocator* sa = ctx_->staging_allocator;
        void* staging = sa->fastMalloc(bytes);
        if (!staging) return -100;
        memcpy(staging, pixels, bytes);
        staging_in_flight_.push_back(staging);
        hsrc = staging;
    }
    int ret = ncnn_cuda_memcpy_h2d_async(raw.data, hsrc, bytes, st);
    if (ret != 0) return ret;
    h2d_bytes += bytes;
    if (!resize)
    {
        dst.create_dims(3, w, h, 1, channels, opt.cuda_elemtype(), n, blob_allocator(opt));
        if (dst.empty()) return -100;
        ncnn_cuda_tensor t = dst.view();
        ret = ncnn_cuda_pixels_to_blob((const unsigned char*)raw.data, channels, w, h, stride, (long long)nstride, swap_rb, mean_vals, norm_vals, &t, st);
        keep_alive_.push_back(raw);
        return ret;
    }
    // the reference's bilinear resize on the device: its offset / coefficient tables are computed here on the host
    // (a few KB), staged through pinned memory and read by the kernel
    const int count = ncnn_cuda_resize_tables_count(target_w, target_h);
    const size_t tbytes = (size_t)count * sizeof(int);
    Allocator* sa = ctx_->staging_allocator;
    int* tab_host = (int*)sa->fastMalloc(tbytes);
    if (!tab_host) return -100;
    staging_in_flight_.push_back(tab_host);
    ret = ncnn_cuda_resize_tables(w, h, target_w, target_h, tab_host);
    if (ret != 0) return ret;
    CudaMat tab;
    tab.create(count, NCNN_CUDA_F32, 1, workspace_allocator(opt));
    if (tab.empty()) return -100;
    ret = ncnn_cuda_memcpy_h2d_async(tab.data, tab_host, tbytes, st);
    if (ret != 0) return ret;
    h2d_bytes += tbytes;
    dst.create_dims(3, target_w, target_h, 1, channels, opt.cuda_elemtype(), n, blob_allocator(opt));
    if (dst.empty()) return -100;
    ncnn_cuda_tensor t = dst.view();
    ret = ncnn_cuda_pixels_resize_to_blob((const unsigned char*)raw.data, channels, w, h, stride, (long long)nstride, swap_rb, mean_vals, norm_vals, (const int*)tab.data, &t, st);
    keep_alive_.push_back(raw);
    keep_alive_.push_back(tab);
    return ret;
}

int CudaCompute::record_download(const CudaMat& src, Mat& dst, const Option& opt)
{
    if (src.empty()) return -100;
    void* st = stream();
    dst.create_dims(src.dims, src.w, src.h, src.d, src.c, src.n, 4u, opt.blob_allocator);
    if (dst.empty()) return -100;
    const size_t bytes = mat_bytes(dst);
    CudaMat raw;
    raw.create((int)((bytes + 3) / 4), NCNN_CUDA_F32, 1, workspace_allocator(opt));
    if (raw.empty()) return -100;
    ncnn_cuda_hostmat hm = host_view(dst, raw.data);
    ncnn_cuda_tensor t = src.view();
    int ret = ncnn_cuda_unpack_to_planar(&t, &hm, st);
    if (ret != 0) return ret;
    keep_alive_.push_back(raw);
    keep_alive_.push_back(src);
    if (ncnn_cuda_host_is_pinned(dst.data))
    {
        ret = ncnn_cuda_memcpy_d2h_async(dst.data, raw.data, bytes, st);
    }
    else
    {
        Allocator* sa = ctx_->staging_allocator;
        void* staging = sa->fastMalloc(bytes);
        if (!staging) return -100;
        ret = ncnn_cuda_memcpy_d2h_async(staging, raw.data, bytes, st);
        PendingDownload pd;
        pd.staging = staging;
        pd.dst = dst.data;
        pd.bytes = bytes;
        downloads_.push_back(pd);
    }
    d2h_bytes += bytes;
    return ret;
}

int CudaCompute::record_clone(const CudaMat& src, CudaMat& dst, const Option& opt)
{
    if (src.empty()) return -100;
    dst.create_like(src, blob_allocator(opt));
    if (dst.empty()) return -100;
    // a channel-range view (Slice as a view, a Concat input written in place) has its parent's pitch: the clone is dense, so
    // the copy is a pitched gather, not one flat memcpy
    if (src.dims == 3 && (dst.cpitch != src.cpitch || dst.nstep != src.nstep))
    {
        ncnn_cuda_tensor s = src.view(), d = dst.view();
        return ncnn_cuda_copy_from_axis(&s, &d, 0, 0, stream());
    }
    return ncnn_cuda_memcpy_d2d_async(dst.data, src.data, src.total_elems() * src.elemsize(), stream());
}

int CudaCompute::submit_and_wait()
{
    int ret = ncnn_cuda_stream_sync(stream());
    for (size_t i = 0; i < timings_.size(); i++)
    {
        if (ret == 0 && timings_[i].start && timings_[i].stop && timings_[i].ms == 0.f)
            ncnn_cuda_event_elapsed_ms(timings_[i].start, timings_[i].stop, &timings_[i].ms);
    }
    Allocator* sa = ctx_->staging_allocator;
    for (size_t i = 0; i < downloads_.size(); i++)
    {
        if (ret == 0) memcpy(downloads_[i].dst, downloads_[i].staging, downloads_[i].bytes);
        sa->fastFree(downloads_[i].staging);
    }
    downloads_.clear();
    for (size_t i = 0; i < staging_in_flight_.size(); i++) sa->fastFree(staging_in_flight_[i]);
    staging_in_flight_.clear();
    keep_alive_.clear();
    return ret;
}

int CudaCompute::reset()
{
    return submit_and_wait();
}

} // namespace ncnn
