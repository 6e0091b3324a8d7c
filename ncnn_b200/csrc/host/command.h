// command.h -- cudaStream-based command recorder, the CUDA sibling of VkCompute (reference: src/command.h:22-88,
// src/command.cpp:358 record_upload, :439 record_download, :1249 record_clone, :1834 submit_and_wait).
//
// A CudaContext bundles what one in-flight extract needs: a stream, a pooled blob allocator whose reuse is ordered
// by that stream, and a pinned staging allocator.  Contexts are acquired/reclaimed per extract like the reference's
// vkdev->acquire_blob_allocator()/reclaim_blob_allocator() (src/net.cpp:2885-2933).
#ifndef NCNN_B200_COMMAND_H
#define NCNN_B200_COMMAND_H

#include <vector>

#include "mat.h"
#include "option.h"

namespace ncnn {

class NCNN_EXPORT CudaContext
{
public:
    explicit CudaContext(int device_index);
    ~CudaContext();
    int device_index;
    void* stream;
    CudaBlobAllocator* blob_allocator;
    CudaStagingAllocator* staging_allocator;
};

// process-wide pool of contexts per device
NCNN_EXPORT CudaContext* acquire_cuda_context(int device_index);
NCNN_EXPORT void reclaim_cuda_context(CudaContext* ctx);
NCNN_EXPORT int get_cuda_device_count();
// weights live until the process ends or Net::clear
NCNN_EXPORT CudaWeightAllocator* get_cuda_weight_allocator(int device_index);

class NCNN_EXPORT CudaCompute
{
public:
    explicit CudaCompute(CudaContext* ctx);
    ~CudaCompute();

    void* stream() const
    {
        return ctx_->stream;
    }
    CudaContext* context() const
    {
        return ctx_;
    }
    CudaAllocator* blob_allocator(const Option& opt) const
    {
        return opt.blob_cuda_allocator ? opt.blob_cuda_allocator : (CudaAllocator*)ctx_->blob_allocator;
    }
    CudaAllocator* workspace_allocator(const Option& opt) const
    {
        return opt.workspace_cuda_allocator ? opt.workspace_cuda_allocator : (CudaAllocator*)ctx_->blob_allocator;
    }

    // host planar fp32 Mat -> device blob of opt.cuda_elemtype() (H2D copy + layout/dtype conversion kernel)
    int record_upload(const Mat& src, CudaMat& dst, const Option& opt);
    // interleaved 8-bit pixels (Mat::from_pixels types PIXEL_RGB/BGR/GRAY/RGBA/BGRA and the RGB<->BGR conversions) -> device blob with
    // (pixel - mean) * norm fused; the raw bytes are what crosses PCIe
    int record_upload_pixels(const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, const float* mean_vals, const float* norm_vals,
                             CudaMat& dst, const Option& opt, int target_w = 0, int target_h = 0);
    // device blob -> host planar fp32 Mat; dst is valid after submit_and_wait()
    int record_download(const CudaMat& src, Mat& dst, const Option& opt);
    // deep copy on the device
    int record_clone(const CudaMat& src, CudaMat& dst, const Option& opt);
    // wait for everything recorded so far and complete the pending downloads
    int submit_and_wait();
    int reset();

    // per-layer device timing, the CUDA form of the reference's NCNN_BENCHMARK / Vulkan timestamp queries
    // (src/net.cpp:302-315, src/benchmark.cpp:30-152): an event pair around every layer of the walk
    struct LayerTiming
    {
        int layer_index;
        void* start;
        void* stop;
        int dims, w, h, d, c, n; // shape of the layer's first top blob
        float ms;                // valid after submit_and_wait()
    };
    void set_profiling(bool enable)
    {
        profiling_ = enable;
    }
    bool profiling() const
    {
        return profiling_;
    }
    void profile_begin(int layer_index);
    void profile_end(const CudaMat& top);
    const std::vector<LayerTiming>& timings() const
    {
        return timings_;
    }
    void clear_timings();

    // bytes moved by record_upload / record_download since construction (bench.py's e2e accounting)
    size_t h2d_bytes;
    size_t d2h_bytes;

private:
    struct PendingDownload
    {
        void* staging;
        void* dst;
        size_t bytes;
    };
    CudaContext* ctx_;
    bool profiling_;
    std::vector<LayerTiming> timings_;
    std::vector<PendingDownload> downloads_;
    std::vector<void*> staging_in_flight_;
    std::vector<CudaMat> keep_alive_;
};

} // namespace ncnn

#endif // NCNN_B200_COMMAND_H
