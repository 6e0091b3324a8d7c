#include "option.h"

#include "ncnn_cuda.h"

namespace ncnn {

Option::Option()
{
    lightmode = true;
    use_shader_pack8 = false;
    use_subgroup_ops = false;
    use_reserved_0 = false;
    num_threads = 1;
    blob_allocator = 0;
    workspace_allocator = 0;
    blob_cuda_allocator = 0;
    workspace_cuda_allocator = 0;
    staging_cuda_allocator = 0;
    openmp_blocktime = 20;
    use_winograd_convolution = true;
    use_sgemm_convolution = true;
    use_int8_inference = true;
    use_vulkan_compute = false;
    use_cuda_compute = true;
    cuda_device_index = -1; // current device
    use_bf16_packed = false;
    use_fp16_packed = true;
    use_fp16_storage = true; // the reference's default (option.cpp:48): GPU blobs are 16-bit floats
    use_fp16_arithmetic = true;
    use_int8_packed = true;
    use_int8_storage = true;
    use_int8_arithmetic = false;
    use_packing_layout = true;
    vulkan_device_index = -1;
    use_tensor_storage = false;
    use_reserved_1p = false;
    flush_denormals = 3;
    use_local_pool_allocator = true;
    use_shader_local_memory = true;
    use_cooperative_matrix = true;
    use_winograd23_convolution = true;
    use_winograd43_convolution = true;
    use_winograd63_convolution = true;
    use_a53_a55_optimized_kernel = false;
    use_fp16_uniform = true;
    use_int8_uniform = true;
    use_bf16_storage = false;
    use_cuda_graph_fusion = true;
    use_cuda_graph = false;
    use_mapped_model_loading = false;
}

int Option::cuda_elemtype() const
{
    if (use_bf16_storage) return NCNN_CUDA_BF16;
    if (use_fp16_storage) return NCNN_CUDA_F16;
    return NCNN_CUDA_F32;
}

} // namespace ncnn
