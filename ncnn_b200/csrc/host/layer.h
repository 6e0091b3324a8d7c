// layer.h -- the operator plugin interface, kept as the reference defines it (src/layer.h:20-142): load_param /
// load_model / create_pipeline / destroy_pipeline / forward / forward_inplace with Option, the capability flags the
// executor reads, and the registry + creator functions (src/layer.h:145-199).
//
// The device overloads take CudaMat + CudaCompute where the reference's Vulkan overloads take VkMat + VkCompute
// (src/layer.h:106-117).  Every layer of this runtime is a CUDA layer (support_cuda = true, support_batch = true):
// the host-Mat overloads are convenience wrappers that upload, run the device overload and download -- there is
// no CPU compute path.
#ifndef NCNN_B200_LAYER_H
#define NCNN_B200_LAYER_H

#include <string>
#include <vector>

#include "command.h"
#include "mat.h"
#include "modelbin.h"
#include "option.h"
#include "paramdict.h"

namespace ncnn {

class NCNN_EXPORT Layer
{
public:
    Layer();
    virtual ~Layer();

    virtual int load_param(const ParamDict& pd);
    virtual int load_model(const ModelBin& mb);
    // weights are re-packed and uploaded here, once (the reference's create_pipeline + upload_model)
    virtual int create_pipeline(const Option& opt);
    virtual int destroy_pipeline(const Option& opt);

public:
    bool one_blob_only;
    bool support_inplace;
    bool support_vulkan;  // always false
    bool support_packing; // always false: the device layout is private, host Mats are elempack 1
    bool support_bf16_storage;
    bool support_fp16_storage;
    bool support_int8_storage;
    bool support_cuda;    // takes the reference's support_reserved slot (src/layer.h:80-87)
    bool support_batch;   // CUDA layers consume the whole batch in one launch (vs src/net.cpp:654-705)
    int featmask;

public:
    // host overloads (upload -> device forward -> download)
    virtual int forward(const std::vector<Mat>& bottom_blobs, std::vector<Mat>& top_blobs, const Option& opt) const;
    virtual int forward(const Mat& bottom_blob, Mat& top_blob, const Option& opt) const;
    virtual int forward_inplace(std::vector<Mat>& bottom_top_blobs, const Option& opt) const;
    virtual int forward_inplace(Mat& bottom_top_blob, const Option& opt) const;

    // device overloads
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const;
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const;
    virtual int forward_inplace(std::vector<CudaMat>& bottom_top_blobs, CudaCompute& cmd, const Option& opt) const;
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const;

public:
    void* userdata;
    int typeindex;
    std::string type;
    std::string name;
    std::vector<int> bottoms;
    std::vector<int> tops;
    std::vector<Mat> bottom_shapes;
    std::vector<Mat> top_shapes;
    int top_count_hint; // number of tops when a multi-top layer is driven outside a Net (host wrappers)
};

typedef Layer* (*layer_creator_func)(void* userdata);
typedef void (*layer_destroyer_func)(Layer* layer, void* userdata);

struct layer_registry_entry
{
    const char* name;
    layer_creator_func creator;
};

// typeindex = position in the reference's registry order (src/CMakeLists.txt:66-175) so .param.bin files resolve
NCNN_EXPORT int layer_to_index(const char* type);
NCNN_EXPORT const char* layer_index_to_type(int typeindex);
NCNN_EXPORT Layer* create_layer(const char* type);
NCNN_EXPORT Layer* create_layer(int typeindex);
NCNN_EXPORT Layer* create_layer_cuda(const char* type);

#define DEFINE_LAYER_CREATOR(name)                          \
    ::ncnn::Layer* name##_layer_creator(void* /*userdata*/) \
    {                                                       \
        return new name;                                    \
    }

// shared helper: fused activation descriptor from (activation_type, activation_params) as fused_activation.h:10-64
ncnn_cuda_activation make_activation(int activation_type, const Mat& activation_params);

} // namespace ncnn

#endif // NCNN_B200_LAYER_H
