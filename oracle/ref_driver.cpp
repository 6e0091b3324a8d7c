// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A few extra C entry points compiled INTO oracle/_ref/libncnn_ref_<isa>.so next to the
// reference's own c_api (src/c_api.h).  They expose things the reference keeps C++-only
// but its own tests rely on:
//   * create_layer_naive  (src/layer.h:185) -- the scalar ground truth tests/testutil.cpp:1301 uses
//   * create_layer_cpu    (src/layer.h:186)
//   * Option fields the c_api has no setter for (lightmode, flush_denormals)
//   * get_physical_big_cpu_count / omp thread control (src/cpu.h) for the CPU baseline report
// This file is ours; it only CALLS the unmodified reference.
#include "c_api.h"
#include "cpu.h"
#include "layer.h"
#include "modelbin.h"
#include "net.h"
#include "option.h"

#include <stdlib.h>
#include <vector>

extern "C" {

// Swap the implementation object inside a c_api layer handle for another Layer*.
static ncnn_layer_t wrap_as(const char* type, ncnn::Layer* impl)
{
    if (!impl) return 0;
    ncnn_layer_t l = ncnn_layer_create_by_type(type);
    if (!l)
    {
        delete impl;
        return 0;
    }
    delete (ncnn::Layer*)l->pthis;
    l->pthis = impl;
    return l;
}

ncnn_layer_t ref_layer_create_naive(const char* type)
{
    return wrap_as(type, ncnn::create_layer_naive(type));
}

ncnn_layer_t ref_layer_create_cpu(const char* type)
{
    return wrap_as(type, ncnn::create_layer_cpu(type));
}

// load_model from an array of Mats, the way tests/testutil.cpp:1331 feeds ModelBinFromMatArray.
// (The reference's own ncnn_modelbin_create_from_mat_array keeps a pointer into a vector that dies when it
// returns -- src/c_api.cpp:1115-1128 -- so the oracle goes to the C++ class directly.)
int ref_layer_load_model_from_mats(ncnn_layer_t layer, const ncnn_mat_t* weights, int n)
{
    std::vector<ncnn::Mat> mats(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) mats[i] = *(const ncnn::Mat*)weights[i];
    ncnn::ModelBinFromMatArray mb(&mats[0]);
    return ((ncnn::Layer*)layer->pthis)->load_model(mb);
}

int ref_layer_get_support_batch(const ncnn_layer_t layer)
{
    return ((const ncnn::Layer*)layer->pthis)->support_batch ? 1 : 0;
}

void ref_option_set_lightmode(ncnn_option_t opt, int enable)
{
    ((ncnn::Option*)opt)->lightmode = enable != 0;
}

void ref_option_set_flush_denormals(ncnn_option_t opt, int v)
{
    ((ncnn::Option*)opt)->flush_denormals = v;
}

int ref_cpu_count(void)
{
    return ncnn::get_cpu_count();
}

int ref_physical_big_cpu_count(void)
{
    return ncnn::get_physical_big_cpu_count();
}

void ref_set_omp_num_threads(int n)
{
    ncnn::set_omp_num_threads(n);
}

// blob index by name (Net::find_blob_index_by_name is protected; input/output name lists are public)
int ref_net_layer_count(const ncnn_net_t net)
{
    return (int)((const ncnn::Net*)net->pthis)->layers().size();
}

int ref_net_blob_count(const ncnn_net_t net)
{
    return (int)((const ncnn::Net*)net->pthis)->blobs().size();
}

const char* ref_net_blob_name(const ncnn_net_t net, int i)
{
    return ((const ncnn::Net*)net->pthis)->blobs()[i].name.c_str();
}

const char* ref_net_layer_type(const ncnn_net_t net, int i)
{
    return ((const ncnn::Net*)net->pthis)->layers()[i]->type.c_str();
}

} // extern "C"
