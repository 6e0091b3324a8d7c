// net.h -- graph loader and executor: Net / Extractor with the reference's public surface (src/net.h:27-248).
//
// load_param parses the .param text (src/net.cpp:1305-1665), load_model feeds each layer its weights in layer order
// and calls create_pipeline (src/net.cpp:2021-2155), Extractor::extract pulls the requested blob through a lazy
// depth-first walk (src/net.cpp:123-190) -- structurally the reference's NCNN_VULKAN branch (src/net.cpp:192-356,
// :886-1140, :3083-3116) with CudaMat / CudaCompute in place of VkMat / VkCompute: one upload at the first device
// layer, every layer enqueued on one stream, one download + stream sync at the end.  Batched Mats (n > 1) go through
// every layer in one launch (all CUDA layers set support_batch) instead of the per-sample loop of src/net.cpp:654-705.
#ifndef NCNN_B200_NET_H
#define NCNN_B200_NET_H

#include <string>
#include <vector>

#include "blob.h"
#include "command.h"
#include "datareader.h"
#include "layer.h"
#include "mat.h"
#include "option.h"

namespace ncnn {

class Extractor;
class NetPrivate;

class NCNN_EXPORT Net
{
public:
    Net();
    virtual ~Net();

    Option opt;

    // device the Net's weights live on and its extractors run on (reference: Net::set_vulkan_device, net.cpp:2555)
    void set_cuda_device(int device_index);
    int cuda_device() const;

    // replace or add an operator type (reference: src/net.cpp:1214-1237); looked up before the built-in registry
    int register_custom_layer(const char* type, layer_creator_func creator, layer_destroyer_func destroyer = 0, void* userdata = 0);

    int load_param(const DataReader& dr);
    int load_param(FILE* fp);
    int load_param(const char* protopath);
    int load_param_mem(const char* mem);
    int load_param_bin(const DataReader& dr);
    int load_param_bin(const char* protopath);

    int load_model(const DataReader& dr);
    int load_model(FILE* fp);
    int load_model(const char* modelpath);
    // returns bytes consumed
    // src/net.cpp:2337-2343 (Net::load_param(const unsigned char*)): a .param.bin image in memory, returns the bytes consumed
    size_t load_param_bin_mem(const unsigned char* mem);
    size_t load_model(const unsigned char* mem);

    void clear();

    Extractor create_extractor() const;

    const std::vector<int>& input_indexes() const;
    const std::vector<int>& output_indexes() const;
    const std::vector<const char*>& input_names() const;
    const std::vector<const char*>& output_names() const;
    const std::vector<Blob>& blobs() const;
    const std::vector<Layer*>& layers() const;

    int find_blob_index_by_name(const char* name) const;
    int find_layer_index_by_name(const char* name) const;

    // number of layers folded into a neighbour by the load-time fusion pass (opt.use_cuda_graph_fusion)
    int fused_layer_count() const;

protected:
    friend class Extractor;
    friend class NetPrivate;
    int load_param_text(const std::string& text);
    Layer* create_layer_by_type(const char* type, int* custom_index);

private:
    Net(const Net&);
    Net& operator=(const Net&);
    NetPrivate* const d;
};

class ExtractorPrivate;
class NCNN_EXPORT Extractor
{
public:
    virtual ~Extractor();
    Extractor(const Extractor&);
    Extractor& operator=(const Extractor&);

    void clear();
    void set_light_mode(bool enable);
    void set_blob_allocator(Allocator* allocator);
    void set_workspace_allocator(Allocator* allocator);
    void set_blob_cuda_allocator(CudaAllocator* allocator);

    int input(const char* blob_name, const Mat& in);
    int extract(const char* blob_name, Mat& feat, int type = 0);
    int input(int blob_index, const Mat& in);
    int extract(int blob_index, Mat& feat, int type = 0);

    // device-resident entry points (reference: Extractor::input/extract with VkMat, net.cpp:3036-3116)
    int input(const char* blob_name, const CudaMat& in);
    int input(int blob_index, const CudaMat& in);
    int extract(const char* blob_name, CudaMat& feat, CudaCompute& cmd);
    int extract(int blob_index, CudaMat& feat, CudaCompute& cmd);

    // device pre-processing: the blob is fed as n interleaved 8-bit images; Mat::from_pixels(type) and
    // substract_mean_normalize(mean_vals, norm_vals) run on the device at extract time.  The pixel memory must stay valid
    // until extract() returns (it is read by the upload, like a Mat view).
    int input_pixels(const char* blob_name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, const float* mean_vals,
                     const float* norm_vals);
    // the same with Mat::from_pixels_resize (src/mat_pixel.cpp:2546-2578): every image is resized to target_w x target_h by the
    // reference's 8-bit bilinear resize, on the device, before the conversion
    int input_pixels_resize(const char* blob_name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, int target_w, int target_h,
                            const float* mean_vals, const float* norm_vals);

    // device post-processing: forward to the YOLOv8 prediction blob, decode it on the device
    // (examples/yolov8.cpp:160-273 generate_proposals) and download one {x, y, w, h, prob, label} row per anchor
    int extract_yolov8_proposals(const char* blob_name, const int* strides, int num_strides, int in_w, int in_h, float prob_threshold, Mat& proposals);

    // bytes moved over PCIe by the last extract(Mat&) call
    size_t last_h2d_bytes() const;
    size_t last_d2h_bytes() const;

protected:
    friend Extractor Net::create_extractor() const;
    Extractor(const Net* net, size_t blob_count);

private:
    ExtractorPrivate* const d;
};

} // namespace ncnn

#endif // NCNN_B200_NET_H
