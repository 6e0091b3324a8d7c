// layer_registry.cpp -- type name <-> typeindex <-> creator (reference: src/layer.cpp:408-538, generated
// layer_registry.h).  Only the hot-path operators have creators; asking for anything else returns 0 and the Net
// fails the load loudly -- there is no CPU fallback inside a graph.
#include <string.h>

#include "layer.h"
#include "layer/cuda_layers.h"
#include "layer_type_table.h"

namespace ncnn {

DEFINE_LAYER_CREATOR(Input)
DEFINE_LAYER_CREATOR(Convolution)
DEFINE_LAYER_CREATOR(ConvolutionDepthWise)
DEFINE_LAYER_CREATOR(InnerProduct)
DEFINE_LAYER_CREATOR(Pooling)
DEFINE_LAYER_CREATOR(Gemm)
DEFINE_LAYER_CREATOR(ReLU)
DEFINE_LAYER_CREATOR(Sigmoid)
DEFINE_LAYER_CREATOR(Swish)
DEFINE_LAYER_CREATOR(TanH)
DEFINE_LAYER_CREATOR(Mish)
DEFINE_LAYER_CREATOR(Clip)
DEFINE_LAYER_CREATOR(HardSwish)
DEFINE_LAYER_CREATOR(HardSigmoid)
DEFINE_LAYER_CREATOR(Dropout)
DEFINE_LAYER_CREATOR(Eltwise)
DEFINE_LAYER_CREATOR(BinaryOp)
DEFINE_LAYER_CREATOR(Split)
DEFINE_LAYER_CREATOR(Concat)
DEFINE_LAYER_CREATOR(Slice)
DEFINE_LAYER_CREATOR(Interp)
DEFINE_LAYER_CREATOR(Softmax)
DEFINE_LAYER_CREATOR(Reshape)
DEFINE_LAYER_CREATOR(Flatten)
DEFINE_LAYER_CREATOR(Permute)
DEFINE_LAYER_CREATOR(Padding)
DEFINE_LAYER_CREATOR(BatchNorm)
DEFINE_LAYER_CREATOR(Scale)
DEFINE_LAYER_CREATOR(ShuffleChannel)
DEFINE_LAYER_CREATOR(LRN)
DEFINE_LAYER_CREATOR(Noop)
DEFINE_LAYER_CREATOR(Crop)
DEFINE_LAYER_CREATOR(Reduction)
DEFINE_LAYER_CREATOR(LayerNorm)
DEFINE_LAYER_CREATOR(MultiHeadAttention)
DEFINE_LAYER_CREATOR(GELU)
DEFINE_LAYER_CREATOR(MemoryData)
DEFINE_LAYER_CREATOR(Deconvolution)
DEFINE_LAYER_CREATOR(DeconvolutionDepthWise)

static const layer_registry_entry cuda_layer_registry[] = {
    {"Input", Input_layer_creator},
    {"Convolution", Convolution_layer_creator},
    {"ConvolutionDepthWise", ConvolutionDepthWise_layer_creator},
    {"InnerProduct", InnerProduct_layer_creator},
    {"Pooling", Pooling_layer_creator},
    {"Gemm", Gemm_layer_creator},
    {"ReLU", ReLU_layer_creator},
    {"Sigmoid", Sigmoid_layer_creator},
    {"Swish", Swish_layer_creator},
    {"TanH", TanH_layer_creator},
    {"Mish", Mish_layer_creator},
    {"Clip", Clip_layer_creator},
    {"HardSwish", HardSwish_layer_creator},
    {"HardSigmoid", HardSigmoid_layer_creator},
    {"Dropout", Dropout_layer_creator},
    {"Eltwise", Eltwise_layer_creator},
    {"BinaryOp", BinaryOp_layer_creator},
    {"Split", Split_layer_creator},
    {"Concat", Concat_layer_creator},
    {"Slice", Slice_layer_creator},
    {"Interp", Interp_layer_creator},
    {"Softmax", Softmax_layer_creator},
    {"Reshape", Reshape_layer_creator},
    {"Flatten", Flatten_layer_creator},
    {"Permute", Permute_layer_creator},
    {"Padding", Padding_layer_creator},
    {"BatchNorm", BatchNorm_layer_creator},
    {"Scale", Scale_layer_creator},
    {"ShuffleChannel", ShuffleChannel_layer_creator},
    {"LRN", LRN_layer_creator},
    {"Noop", Noop_layer_creator},
    {"Crop", Crop_layer_creator},
    {"Reduction", Reduction_layer_creator},
    {"LayerNorm", LayerNorm_layer_creator},
    {"MultiHeadAttention", MultiHeadAttention_layer_creator},
    {"GELU", GELU_layer_creator},
    {"MemoryData", MemoryData_layer_creator},
    {"Deconvolution", Deconvolution_layer_creator},
    {"DeconvolutionDepthWise", DeconvolutionDepthWise_layer_creator},
};

static const int layer_type_count = (int)(sizeof(layer_type_names) / sizeof(layer_type_names[0]));

int layer_to_index(const char* type)
{
    for (int i = 0; i < layer_type_count; i++)
        if (strcmp(type, layer_type_names[i]) == 0) return i;
    return -1;
}

const char* layer_index_to_type(int typeindex)
{
    if (typeindex < 0 || typeindex >= layer_type_count) return 0;
    return layer_type_names[typeindex];
}

Layer* create_layer_cuda(const char* type)
{
    const int n = (int)(sizeof(cuda_layer_registry) / sizeof(cuda_layer_registry[0]));
    for (int i = 0; i < n; i++)
    {
        if (strcmp(type, cuda_layer_registry[i].name) == 0)
        {
            Layer* layer = cuda_layer_registry[i].creator(0);
            layer->type = type;
            layer->typeindex = layer_to_index(type);
            return layer;
        }
    }
    return 0;
}

Layer* create_layer(const char* type)
{
    return create_layer_cuda(type);
}

Layer* create_layer(int typeindex)
{
    const char* type = layer_index_to_type(typeindex);
    return type ? create_layer_cuda(type) : 0;
}

} // namespace ncnn
