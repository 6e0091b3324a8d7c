// option.h -- run-time options, a flat struct as in the reference (src/option.h, defaults src/option.cpp:10-76).
// New fields for this backend take the place of the reference's Vulkan ones:
//   use_cuda_compute       <-> use_vulkan_compute     blob_cuda_allocator    <-> blob_vkallocator
//   cuda_device_index      <-> Net::set_vulkan_device workspace_cuda_allocator / staging_cuda_allocator likewise
// Storage type of device blobs follows the reference's own switches: use_bf16_storage -> bf16 blobs (tcgen05 kind::f16
// with bf16 operands), else use_fp16_storage -> fp16 blobs (tcgen05 f16 operands), else fp32 blobs on the CUDA-core path.
#ifndef NCNN_B200_OPTION_H
#define NCNN_B200_OPTION_H

#include "platform.h"

namespace ncnn {

class Allocator;
class CudaAllocator;

class NCNN_EXPORT Option
{
public:
    Option();

    bool lightmode;              // release a blob as soon as its last consumer ran (src/net.cpp:729-733)
    bool use_shader_pack8;       // unused, layout compatibility of the flag set
    bool use_subgroup_ops;       // unused
    bool use_reserved_0;
    int num_threads;             // host threads are irrelevant to the CUDA path; kept for the oracle-side tools
    Allocator* blob_allocator;      // host Mats returned by extract
    Allocator* workspace_allocator; // host scratch
    CudaAllocator* blob_cuda_allocator;
    CudaAllocator* workspace_cuda_allocator;
    Allocator* staging_cuda_allocator; // pinned host memory
    int openmp_blocktime;
    bool use_winograd_convolution; // accepted and ignored: every conv is an implicit GEMM on tcgen05
    bool use_sgemm_convolution;    // accepted and ignored
    bool use_int8_inference;       // int8 models are out of scope: load fails loudly
    bool use_vulkan_compute;       // accepted and ignored (no Vulkan on the target)
    bool use_cuda_compute;         // default true: this runtime has no CPU compute path
    int cuda_device_index;
    bool use_bf16_packed;
    bool use_fp16_packed;
    bool use_fp16_storage;
    bool use_fp16_arithmetic;
    bool use_int8_packed;
    bool use_int8_storage;
    bool use_int8_arithmetic;
    bool use_packing_layout;
    int vulkan_device_index;
    bool use_tensor_storage;
    bool use_reserved_1p;
    int flush_denormals;           // kernels are compiled with --ftz=true: same effect as the reference's DAZ|FTZ (net.cpp:2863)
    bool use_local_pool_allocator;
    bool use_shader_local_memory;
    bool use_cooperative_matrix;
    bool use_winograd23_convolution;
    bool use_winograd43_convolution;
    bool use_winograd63_convolution;
    bool use_a53_a55_optimized_kernel;
    bool use_fp16_uniform;
    bool use_int8_uniform;
    bool use_bf16_storage;
    // fold Conv -> (Eltwise SUM) -> ReLU / Conv -> Swish chains into the conv epilogue at load time
    // (the reference does this offline, tools/ncnnoptimize.cpp:1268-1419); results are unchanged up to rounding
    bool use_cuda_graph_fusion;
    // replay the recorded walk of an extract as one CUDA graph when shapes repeat
    bool use_cuda_graph;
    bool use_mapped_model_loading;

    // element type (include/ncnn_cuda.h NCNN_CUDA_*) of the device blobs these options select
    int cuda_elemtype() const;
};

} // namespace ncnn

#endif // NCNN_B200_OPTION_H
