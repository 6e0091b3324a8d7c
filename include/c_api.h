/* c_api.h -- the flat C ABI of the host runtime in libncnn_b200.so.
 *
 * Every entry point in sections 1-9 has the SAME name, argument list and meaning as the reference's
 * src/c_api.h (Tencent/ncnn @ a4d2ea1d; the line each one replaces is cited), so a program or binding written
 * against the reference's C API (python ctypes, cgo, JNI, ...) links against this library unchanged for the
 * hot path: load .param/.bin, Extractor input/extract, custom layers through the function-pointer table.
 * Section 10 adds what a CUDA backend needs where the reference has its Vulkan-only switches.
 *
 * Opaque handles are plain pointers; return codes: 0 ok, -1 invalid/unsupported, -100 out of memory.
 */
#ifndef NCNN_B200_C_API_H
#define NCNN_B200_C_API_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NCNN_C_API __attribute__((visibility("default")))
#else
#define NCNN_C_API
#endif

/* 1. version (src/c_api.h:18-19) */
NCNN_C_API const char* ncnn_version(void);
NCNN_C_API int ncnn_version_number(void);

/* 2. allocator (src/c_api.h:21-33) */
typedef struct __ncnn_allocator_t* ncnn_allocator_t;
struct __ncnn_allocator_t
{
    void* pthis;
    void* (*fast_malloc)(ncnn_allocator_t allocator, size_t size);
    void (*fast_free)(ncnn_allocator_t allocator, void* ptr);
};
NCNN_C_API ncnn_allocator_t ncnn_allocator_create_pool_allocator(void);
NCNN_C_API ncnn_allocator_t ncnn_allocator_create_unlocked_pool_allocator(void);
NCNN_C_API void ncnn_allocator_destroy(ncnn_allocator_t allocator);

/* 3. option (src/c_api.h:55-103) */
typedef struct __ncnn_option_t* ncnn_option_t;
NCNN_C_API ncnn_option_t ncnn_option_create(void);
NCNN_C_API void ncnn_option_destroy(ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_num_threads(const ncnn_option_t opt);
NCNN_C_API void ncnn_option_set_num_threads(ncnn_option_t opt, int num_threads);
NCNN_C_API void ncnn_option_set_blob_allocator(ncnn_option_t opt, ncnn_allocator_t allocator);
NCNN_C_API void ncnn_option_set_workspace_allocator(ncnn_option_t opt, ncnn_allocator_t allocator);
NCNN_C_API int ncnn_option_get_use_vulkan_compute(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_local_pool_allocator(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_winograd_convolution(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_sgemm_convolution(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_packing_layout(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_fp16_packed(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_fp16_storage(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_fp16_arithmetic(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_int8_packed(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_int8_storage(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_int8_arithmetic(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_bf16_packed(const ncnn_option_t opt);
NCNN_C_API int ncnn_option_get_use_bf16_storage(const ncnn_option_t opt);
NCNN_C_API void ncnn_option_set_use_vulkan_compute(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_local_pool_allocator(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_winograd_convolution(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_sgemm_convolution(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_packing_layout(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_fp16_packed(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_fp16_storage(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_fp16_arithmetic(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_int8_packed(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_int8_storage(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_int8_arithmetic(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_bf16_packed(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_bf16_storage(ncnn_option_t opt, int enable);

/* 4. mat (src/c_api.h:105-165) */
typedef struct __ncnn_mat_t* ncnn_mat_t;
NCNN_C_API ncnn_mat_t ncnn_mat_create(void);
NCNN_C_API ncnn_mat_t ncnn_mat_create_1d(int w, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_2d(int w, int h, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_3d(int w, int h, int c, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_4d(int w, int h, int d, int c, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_1d_batch(int w, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_2d_batch(int w, int h, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_3d_batch(int w, int h, int c, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_4d_batch(int w, int h, int d, int c, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_1d(int w, void* data, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_2d(int w, int h, void* data, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_3d(int w, int h, int c, void* data, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_4d(int w, int h, int d, int c, void* data, ncnn_allocator_t allocator);
NCNN_C_API void ncnn_mat_destroy(ncnn_mat_t mat);
NCNN_C_API void ncnn_mat_fill_float(ncnn_mat_t mat, float v);

/* pixel pre-processing on the host, as in the reference (src/c_api.h:167-185); the device path is ncnn_extractor_input_pixels */
#define NCNN_MAT_PIXEL_RGB       1
#define NCNN_MAT_PIXEL_BGR       2
#define NCNN_MAT_PIXEL_GRAY      3
#define NCNN_MAT_PIXEL_RGBA      4
#define NCNN_MAT_PIXEL_BGRA      5
#define NCNN_MAT_PIXEL_X2Y(X, Y) (X | (Y << 16))
/* src/c_api.h:124-137: Mats of another element size / packing (fp16, int8, packed storage made by the caller) */
NCNN_C_API ncnn_mat_t ncnn_mat_create_1d_elem(int w, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_2d_elem(int w, int h, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_3d_elem(int w, int h, int c, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_4d_elem(int w, int h, int d, int c, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_1d_elem_batch(int w, size_t elemsize, int elempack, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_2d_elem_batch(int w, int h, size_t elemsize, int elempack, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_3d_elem_batch(int w, int h, int c, size_t elemsize, int elempack, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_4d_elem_batch(int w, int h, int d, int c, size_t elemsize, int elempack, int n, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_1d_elem(int w, void* data, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_2d_elem(int w, int h, void* data, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_3d_elem(int w, int h, int c, void* data, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_create_external_4d_elem(int w, int h, int d, int c, void* data, size_t elemsize, int elempack, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_from_pixels(const unsigned char* pixels, int type, int w, int h, int stride, ncnn_allocator_t allocator);
NCNN_C_API void ncnn_mat_substract_mean_normalize(ncnn_mat_t mat, const float* mean_vals, const float* norm_vals);
/* src/c_api.h:180 */
NCNN_C_API void ncnn_mat_to_pixels(const ncnn_mat_t mat, unsigned char* pixels, int type, int stride);
/* src/c_api.h:177-181: the resize / roi forms; the resize is the reference's 8-bit bilinear (src/mat_pixel_resize.cpp), bit-exact */
NCNN_C_API ncnn_mat_t ncnn_mat_from_pixels_resize(const unsigned char* pixels, int type, int w, int h, int stride, int target_width, int target_height, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_from_pixels_roi(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_from_pixels_roi_resize(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, int target_width,
                                                      int target_height, ncnn_allocator_t allocator);
NCNN_C_API void ncnn_mat_to_pixels_resize(const ncnn_mat_t mat, unsigned char* pixels, int type, int target_width, int target_height, int target_stride);
NCNN_C_API ncnn_mat_t ncnn_mat_clone(const ncnn_mat_t mat, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_reshape_1d(const ncnn_mat_t mat, int w, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_reshape_2d(const ncnn_mat_t mat, int w, int h, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_reshape_3d(const ncnn_mat_t mat, int w, int h, int c, ncnn_allocator_t allocator);
NCNN_C_API ncnn_mat_t ncnn_mat_reshape_4d(const ncnn_mat_t mat, int w, int h, int d, int c, ncnn_allocator_t allocator);
NCNN_C_API int ncnn_mat_get_dims(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_w(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_h(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_d(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_c(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_n(const ncnn_mat_t mat);
NCNN_C_API size_t ncnn_mat_get_elemsize(const ncnn_mat_t mat);
NCNN_C_API int ncnn_mat_get_elempack(const ncnn_mat_t mat);
NCNN_C_API size_t ncnn_mat_get_cstep(const ncnn_mat_t mat);
NCNN_C_API size_t ncnn_mat_get_nstep(const ncnn_mat_t mat);
NCNN_C_API void* ncnn_mat_get_data(const ncnn_mat_t mat);
NCNN_C_API void* ncnn_mat_get_batch_data(const ncnn_mat_t mat, int b);
NCNN_C_API void* ncnn_mat_get_channel_data(const ncnn_mat_t mat, int c);

/* 5. blob / paramdict (src/c_api.h:190-216) */
typedef struct __ncnn_paramdict_t* ncnn_paramdict_t;
NCNN_C_API ncnn_paramdict_t ncnn_paramdict_create(void);
NCNN_C_API void ncnn_paramdict_destroy(ncnn_paramdict_t pd);
NCNN_C_API int ncnn_paramdict_get_type(const ncnn_paramdict_t pd, int id);
NCNN_C_API int ncnn_paramdict_get_int(const ncnn_paramdict_t pd, int id, int def);
NCNN_C_API float ncnn_paramdict_get_float(const ncnn_paramdict_t pd, int id, float def);
NCNN_C_API ncnn_mat_t ncnn_paramdict_get_array(const ncnn_paramdict_t pd, int id, const ncnn_mat_t def);
NCNN_C_API void ncnn_paramdict_set_int(ncnn_paramdict_t pd, int id, int i);
NCNN_C_API void ncnn_paramdict_set_float(ncnn_paramdict_t pd, int id, float f);
NCNN_C_API void ncnn_paramdict_set_array(ncnn_paramdict_t pd, int id, const ncnn_mat_t v);

/* 6. datareader (src/c_api.h:218-235) */
typedef struct __ncnn_datareader_t* ncnn_datareader_t;
struct __ncnn_datareader_t
{
    void* pthis;
    int (*scan)(ncnn_datareader_t dr, const char* format, void* p);
    size_t (*read)(ncnn_datareader_t dr, void* buf, size_t size);
};
NCNN_C_API ncnn_datareader_t ncnn_datareader_create(void);
NCNN_C_API ncnn_datareader_t ncnn_datareader_create_from_stdio(FILE* fp);
NCNN_C_API ncnn_datareader_t ncnn_datareader_create_from_memory(const unsigned char** mem);
NCNN_C_API void ncnn_datareader_destroy(ncnn_datareader_t dr);

/* 7. modelbin (src/c_api.h:237-250) */
typedef struct __ncnn_modelbin_t* ncnn_modelbin_t;
struct __ncnn_modelbin_t
{
    void* pthis;
    ncnn_mat_t (*load_1d)(const ncnn_modelbin_t mb, int w, int type);
    ncnn_mat_t (*load_2d)(const ncnn_modelbin_t mb, int w, int h, int type);
    ncnn_mat_t (*load_3d)(const ncnn_modelbin_t mb, int w, int h, int c, int type);
};
NCNN_C_API ncnn_modelbin_t ncnn_modelbin_create_from_datareader(const ncnn_datareader_t dr);
/* the Mats are copied (refcount shared) and stay alive as long as the modelbin does */
NCNN_C_API ncnn_modelbin_t ncnn_modelbin_create_from_mat_array(const ncnn_mat_t* weights, int n);
NCNN_C_API void ncnn_modelbin_destroy(ncnn_modelbin_t mb);

/* 8. layer (src/c_api.h:252-318): the plugin table a custom operator fills in */
typedef struct __ncnn_layer_t* ncnn_layer_t;
struct __ncnn_layer_t
{
    void* pthis;
    int (*load_param)(ncnn_layer_t layer, const ncnn_paramdict_t pd);
    int (*load_model)(ncnn_layer_t layer, const ncnn_modelbin_t mb);
    int (*create_pipeline)(ncnn_layer_t layer, const ncnn_option_t opt);
    int (*destroy_pipeline)(ncnn_layer_t layer, const ncnn_option_t opt);
    int (*forward_1)(const ncnn_layer_t layer, const ncnn_mat_t bottom_blob, ncnn_mat_t* top_blob, const ncnn_option_t opt);
    int (*forward_n)(const ncnn_layer_t layer, const ncnn_mat_t* bottom_blobs, int n, ncnn_mat_t* top_blobs, int n2, const ncnn_option_t opt);
    int (*forward_inplace_1)(const ncnn_layer_t layer, ncnn_mat_t bottom_top_blob, const ncnn_option_t opt);
    int (*forward_inplace_n)(const ncnn_layer_t layer, ncnn_mat_t* bottom_top_blobs, int n, const ncnn_option_t opt);
};
NCNN_C_API ncnn_layer_t ncnn_layer_create(void);
NCNN_C_API ncnn_layer_t ncnn_layer_create_by_typeindex(int typeindex);
NCNN_C_API ncnn_layer_t ncnn_layer_create_by_type(const char* type);
NCNN_C_API int ncnn_layer_type_to_index(const char* type);
NCNN_C_API void ncnn_layer_destroy(ncnn_layer_t layer);
NCNN_C_API const char* ncnn_layer_get_name(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_typeindex(const ncnn_layer_t layer);
NCNN_C_API const char* ncnn_layer_get_type(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_one_blob_only(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_inplace(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_vulkan(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_packing(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_bf16_storage(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_fp16_storage(const ncnn_layer_t layer);
NCNN_C_API void ncnn_layer_set_one_blob_only(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_inplace(ncnn_layer_t layer, int enable);
/* src/c_api.h:294-306: the remaining capability flags.  This runtime keeps host Mats unpacked and has no Vulkan path, so the three
 * *_packing / vulkan getters always report 0 and their setters only exist so that a custom layer written against the reference's
 * C API links and runs unchanged (the values are ignored). */
NCNN_C_API int ncnn_layer_get_support_vulkan_packing(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_any_packing(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_support_vulkan_any_packing(const ncnn_layer_t layer);
NCNN_C_API void ncnn_layer_set_support_vulkan(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_packing(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_bf16_storage(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_fp16_storage(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_vulkan_packing(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_any_packing(ncnn_layer_t layer, int enable);
NCNN_C_API void ncnn_layer_set_support_vulkan_any_packing(ncnn_layer_t layer, int enable);
/* src/c_api.h:313-314: the shape hints of a layer's i-th bottom / top blob (dims 0 when the .param carries none) */
NCNN_C_API void ncnn_blob_get_bottom_shape(const ncnn_layer_t layer, int i, int* dims, int* w, int* h, int* c);
NCNN_C_API void ncnn_blob_get_top_shape(const ncnn_layer_t layer, int i, int* dims, int* w, int* h, int* c);
NCNN_C_API int ncnn_layer_get_bottom_count(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_bottom(const ncnn_layer_t layer, int i);
NCNN_C_API int ncnn_layer_get_top_count(const ncnn_layer_t layer);
NCNN_C_API int ncnn_layer_get_top(const ncnn_layer_t layer, int i);

/* 9. net / extractor (src/c_api.h:320-408) */
typedef struct __ncnn_net_t* ncnn_net_t;
struct __ncnn_net_t
{
    void* pthis;
    void* custom_layer_factory;
};
typedef struct __ncnn_extractor_t* ncnn_extractor_t;
typedef ncnn_layer_t (*ncnn_layer_creator_t)(void* userdata);
typedef void (*ncnn_layer_destroyer_t)(ncnn_layer_t layer, void* userdata);
NCNN_C_API ncnn_net_t ncnn_net_create(void);
NCNN_C_API void ncnn_net_destroy(ncnn_net_t net);
NCNN_C_API ncnn_option_t ncnn_net_get_option(ncnn_net_t net);
NCNN_C_API void ncnn_net_set_option(ncnn_net_t net, ncnn_option_t opt);
NCNN_C_API void ncnn_net_register_custom_layer_by_type(ncnn_net_t net, const char* type, ncnn_layer_creator_t creator, ncnn_layer_destroyer_t destroyer, void* userdata);
NCNN_C_API void ncnn_net_register_custom_layer_by_typeindex(ncnn_net_t net, int typeindex, ncnn_layer_creator_t creator, ncnn_layer_destroyer_t destroyer, void* userdata);
NCNN_C_API int ncnn_net_load_param(ncnn_net_t net, const char* path);
NCNN_C_API int ncnn_net_load_param_bin(ncnn_net_t net, const char* path);
NCNN_C_API int ncnn_net_load_model(ncnn_net_t net, const char* path);
NCNN_C_API int ncnn_net_load_param_memory(ncnn_net_t net, const char* mem);
/* src/c_api.h:373: a .param.bin image in memory; returns the bytes consumed (0 on failure) */
NCNN_C_API size_t ncnn_net_load_param_bin_memory(ncnn_net_t net, const unsigned char* mem);
NCNN_C_API size_t ncnn_net_load_model_memory(ncnn_net_t net, const unsigned char* mem);
NCNN_C_API int ncnn_net_load_param_datareader(ncnn_net_t net, const ncnn_datareader_t dr);
NCNN_C_API int ncnn_net_load_param_bin_datareader(ncnn_net_t net, const ncnn_datareader_t dr);
NCNN_C_API int ncnn_net_load_model_datareader(ncnn_net_t net, const ncnn_datareader_t dr);
NCNN_C_API void ncnn_net_clear(ncnn_net_t net);
NCNN_C_API int ncnn_net_get_input_count(const ncnn_net_t net);
NCNN_C_API int ncnn_net_get_output_count(const ncnn_net_t net);
NCNN_C_API const char* ncnn_net_get_input_name(const ncnn_net_t net, int i);
NCNN_C_API const char* ncnn_net_get_output_name(const ncnn_net_t net, int i);
NCNN_C_API int ncnn_net_get_input_index(const ncnn_net_t net, int i);
NCNN_C_API int ncnn_net_get_output_index(const ncnn_net_t net, int i);
NCNN_C_API ncnn_extractor_t ncnn_extractor_create(ncnn_net_t net);
NCNN_C_API void ncnn_extractor_destroy(ncnn_extractor_t ex);
NCNN_C_API void ncnn_extractor_set_option(ncnn_extractor_t ex, const ncnn_option_t opt);
NCNN_C_API int ncnn_extractor_input(ncnn_extractor_t ex, const char* name, const ncnn_mat_t mat);
NCNN_C_API int ncnn_extractor_extract(ncnn_extractor_t ex, const char* name, ncnn_mat_t* mat);
NCNN_C_API int ncnn_extractor_input_index(ncnn_extractor_t ex, int index, const ncnn_mat_t mat);
NCNN_C_API int ncnn_extractor_extract_index(ncnn_extractor_t ex, int index, ncnn_mat_t* mat);

/* 10. CUDA backend additions (where the reference has ncnn_option_set_use_vulkan_compute / ncnn_net_set_vulkan_device,
 * src/c_api.h:85, :345, and its C++-only VkMat/VkCompute extractor overloads, src/net.h:205-230) */
NCNN_C_API int ncnn_option_get_use_cuda_compute(const ncnn_option_t opt);
NCNN_C_API void ncnn_option_set_use_cuda_compute(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_lightmode(ncnn_option_t opt, int enable);
NCNN_C_API void ncnn_option_set_use_cuda_graph_fusion(ncnn_option_t opt, int enable);
/* Option::use_mapped_model_loading (src/option.h:125; C++-only in the reference): ncnn_net_load_model(path) mmap's the .bin and
 * parses it in place (src/net.cpp:2263-2301) */
NCNN_C_API void ncnn_option_set_use_mapped_model_loading(ncnn_option_t opt, int enable);
NCNN_C_API int ncnn_get_cuda_device_count(void);
NCNN_C_API void ncnn_net_set_cuda_device(ncnn_net_t net, int device_index);
NCNN_C_API int ncnn_net_get_fused_layer_count(const ncnn_net_t net);
/* zero-copy view of samples [b, b + batches) of a batched Mat (Mat::batch_range, src/mat.h:241-242): how a host batch
 * is split across per-GPU replicas; the view does not own the data (refcount NULL, as in the reference): the parent
 * must outlive it.  NULL when the range is out of bounds. */
NCNN_C_API ncnn_mat_t ncnn_mat_batch_range(const ncnn_mat_t mat, int b, int batches);
/* pinned host Mats: async H2D/D2H without a staging copy */
NCNN_C_API ncnn_allocator_t ncnn_allocator_create_cuda_staging_allocator(void);
/* device-resident tensors and the stream recorder */
typedef struct __ncnn_cuda_mat_t* ncnn_cuda_mat_t;
typedef struct __ncnn_cuda_compute_t* ncnn_cuda_compute_t;
NCNN_C_API ncnn_cuda_compute_t ncnn_cuda_compute_create(int device_index);
NCNN_C_API void ncnn_cuda_compute_destroy(ncnn_cuda_compute_t cmd);
NCNN_C_API void* ncnn_cuda_compute_get_stream(ncnn_cuda_compute_t cmd);
NCNN_C_API int ncnn_cuda_compute_record_upload(ncnn_cuda_compute_t cmd, const ncnn_mat_t src, ncnn_cuda_mat_t* dst, const ncnn_option_t opt);
NCNN_C_API int ncnn_cuda_compute_record_download(ncnn_cuda_compute_t cmd, const ncnn_cuda_mat_t src, ncnn_mat_t* dst, const ncnn_option_t opt);
NCNN_C_API int ncnn_cuda_compute_submit_and_wait(ncnn_cuda_compute_t cmd);
/* per-layer device times of the walks recorded on `cmd` (the reference's NCNN_BENCHMARK layer timing, src/net.cpp:145-180,
 * :302-315): shape = {dims, w, h, d, c, n} of the layer's first top blob; valid after submit_and_wait */
NCNN_C_API void ncnn_cuda_compute_set_profiling(ncnn_cuda_compute_t cmd, int enable);
NCNN_C_API int ncnn_cuda_compute_get_profile_count(ncnn_cuda_compute_t cmd);
NCNN_C_API int ncnn_cuda_compute_get_profile(ncnn_cuda_compute_t cmd, int i, int* layer_index, float* ms, int shape[6]);
NCNN_C_API void ncnn_cuda_compute_clear_profile(ncnn_cuda_compute_t cmd);
NCNN_C_API int ncnn_net_get_layer_count(const ncnn_net_t net);
NCNN_C_API const char* ncnn_net_get_layer_type(const ncnn_net_t net, int i);
NCNN_C_API const char* ncnn_net_get_layer_name(const ncnn_net_t net, int i);
NCNN_C_API void ncnn_cuda_mat_destroy(ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_dims(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_w(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_h(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_c(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_n(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_cuda_mat_get_elemtype(const ncnn_cuda_mat_t mat);
NCNN_C_API void* ncnn_cuda_mat_get_data(const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_extractor_input_cuda(ncnn_extractor_t ex, const char* name, const ncnn_cuda_mat_t mat);
NCNN_C_API int ncnn_extractor_extract_cuda(ncnn_extractor_t ex, const char* name, ncnn_cuda_mat_t* mat, ncnn_cuda_compute_t cmd);
/* Device pre-processing: feed `name` with n interleaved 8-bit images; the equivalent of ncnn_mat_from_pixels(pixels, type, w, h,
 * stride) (src/c_api.h:104-109, types as NCNN_MAT_PIXEL_*) + ncnn_mat_substract_mean_normalize(mean_vals, norm_vals)
 * (src/c_api.h:117) per image runs on the device at extract time; only the raw bytes cross PCIe.  stride 0 = w * channels,
 * nstride (bytes between images) 0 = h * stride; mean_vals / norm_vals may be NULL.  `pixels` must stay valid until the
 * extract call returns. */
NCNN_C_API int ncnn_extractor_input_pixels(ncnn_extractor_t ex, const char* name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride,
                                           const float* mean_vals, const float* norm_vals);
/* The same with ncnn_mat_from_pixels_resize (src/c_api.h:177): every image is brought to target_w x target_h by the reference's
 * 8-bit bilinear resize (src/mat_pixel_resize.cpp, bit-exact) on the device before the conversion; the source images may be
 * larger or smaller than the network input and only their raw bytes cross PCIe. */
NCNN_C_API int ncnn_extractor_input_pixels_resize(ncnn_extractor_t ex, const char* name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride,
                                                  int target_w, int target_h, const float* mean_vals, const float* norm_vals);
/* Device post-processing for YOLOv8-style heads: runs the graph up to blob `name` (the 2-D prediction blob, see
 * ncnn_cuda_yolov8_decode in ncnn_cuda.h), decodes it on the device as generate_proposals of the reference's
 * examples/yolov8.cpp:160-273 does on the host, and downloads only the decoded records: `*proposals` becomes a batched 2-D
 * fp32 Mat, w = 6 {x, y, width, height, prob, label}, h = anchors, one row per anchor in anchor order; rows below
 * prob_threshold have prob = 0 and label = -1.  in_w / in_h: size of the (padded) network input.  The caller sorts the
 * surviving rows and applies NMS as the example does. */
NCNN_C_API int ncnn_extractor_extract_yolov8_proposals(ncnn_extractor_t ex, const char* name, const int* strides, int num_strides, int in_w, int in_h,
                                                       float prob_threshold, ncnn_mat_t* proposals);
/* Host-side Mat helpers of the reference's C API (src/c_api.h:188, :413-415), for the letterbox / crop steps around a network
 * (examples/yolov8.cpp:331-335 pads with 114).  fp32 Mats of 1 to 3 dims; border type 0 constant `v`, 1 replicate, 2 reflect
 * (src/layer/padding.cpp:21-260).  `opt` supplies the blob allocator (may be NULL). */
NCNN_C_API void ncnn_copy_make_border(const ncnn_mat_t src, ncnn_mat_t dst, int top, int bottom, int left, int right, int type, float v, const ncnn_option_t opt);
/* src/c_api.h:414: also pads the channel axis of a 3-D Mat (front / behind) */
NCNN_C_API void ncnn_copy_make_border_3d(const ncnn_mat_t src, ncnn_mat_t dst, int top, int bottom, int left, int right, int front, int behind, int type, float v,
                                         const ncnn_option_t opt);
NCNN_C_API void ncnn_copy_cut_border(const ncnn_mat_t src, ncnn_mat_t dst, int top, int bottom, int left, int right, const ncnn_option_t opt);
NCNN_C_API void ncnn_flatten(const ncnn_mat_t src, ncnn_mat_t* dst, const ncnn_option_t opt);
/* PCIe bytes of the last ncnn_extractor_extract call */
NCNN_C_API size_t ncnn_extractor_get_last_h2d_bytes(const ncnn_extractor_t ex);
NCNN_C_API size_t ncnn_extractor_get_last_d2h_bytes(const ncnn_extractor_t ex);

#ifdef __cplusplus
}
#endif

#endif /* NCNN_B200_C_API_H */
