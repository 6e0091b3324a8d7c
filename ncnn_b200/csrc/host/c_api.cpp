// c_api.cpp -- implementation of include/c_api.h over the C++ host runtime (counterpart of the reference's
// src/c_api.cpp; same entry-point names so existing bindings keep working).
#include "c_api.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

#include "allocator.h"
#include "command.h"
#include "datareader.h"
#include "layer.h"
#include "mat.h"
#include "modelbin.h"
#include "net.h"
#include "option.h"
#include "paramdict.h"

using namespace ncnn;

// src/c_api.cpp:55-138: the Allocator object behind an ncnn_allocator_t calls back THROUGH the C function table, and the table's
// default entries call the base class non-virtually -- so a C caller may replace fast_malloc / fast_free on the struct it was
// handed and every Mat the runtime allocates from it goes through the replacement (that is the allocator plugin contract).
namespace {
template<class Base>
class TableAllocator : public Base
{
public:
    explicit TableAllocator(ncnn_allocator_t t)
        : table(t)
    {
    }
    virtual void* fastMalloc(size_t size)
    {
        return table->fast_malloc(table, size);
    }
    virtual void fastFree(void* ptr)
    {
        table->fast_free(table, ptr);
    }
    ncnn_allocator_t table;
};
template<class Base>
void* table_default_malloc(ncnn_allocator_t a, size_t size)
{
    return ((TableAllocator<Base>*)(Allocator*)a->pthis)->Base::fastMalloc(size);
}
template<class Base>
void table_default_free(ncnn_allocator_t a, void* ptr)
{
    ((TableAllocator<Base>*)(Allocator*)a->pthis)->Base::fastFree(ptr);
}
template<class Base>
ncnn_allocator_t make_table_allocator()
{
    ncnn_allocator_t a = (ncnn_allocator_t)malloc(sizeof(struct __ncnn_allocator_t));
    a->pthis = (void*)(Allocator*)(new TableAllocator<Base>(a));
    a->fast_malloc = table_default_malloc<Base>;
    a->fast_free = table_default_free<Base>;
    return a;
}
} // namespace

extern "C" {

const char* ncnn_version(void)
{
    return "ncnn-b200-1.0";
}

int ncnn_version_number(void)
{
    return 20261017;
}

// ------------------------------------------------------------------ allocator
ncnn_allocator_t ncnn_allocator_create_pool_allocator(void)
{
    return make_table_allocator<PoolAllocator>();
}
ncnn_allocator_t ncnn_allocator_create_unlocked_pool_allocator(void)
{
    return make_table_allocator<PoolAllocator>();
}
ncnn_allocator_t ncnn_allocator_create_cuda_staging_allocator(void)
{
    return make_table_allocator<CudaStagingAllocator>();
}
void ncnn_allocator_destroy(ncnn_allocator_t a)
{
    if (!a) return;
    delete (Allocator*)a->pthis;
    free(a);
}

static Allocator* alloc_of(ncnn_allocator_t a)
{
    return a ? (Allocator*)a->pthis : 0;
}

// ------------------------------------------------------------------ option
ncnn_option_t ncnn_option_create(void)
{
    return (ncnn_option_t)(new Option);
}
void ncnn_option_destroy(ncnn_option_t opt)
{
    delete (Option*)opt;
}
int ncnn_option_get_num_threads(const ncnn_option_t opt)
{
    return ((const Option*)opt)->num_threads;
}
void ncnn_option_set_num_threads(ncnn_option_t opt, int num_threads)
{
    ((Option*)opt)->num_threads = num_threads;
}
void ncnn_option_set_blob_allocator(ncnn_option_t opt, ncnn_allocator_t allocator)
{
    ((Option*)opt)->blob_allocator = alloc_of(allocator);
}
void ncnn_option_set_workspace_allocator(ncnn_option_t opt, ncnn_allocator_t allocator)
{
    ((Option*)opt)->workspace_allocator = alloc_of(allocator);
}

#define OPT_FLAG(name)                                                   \
    int ncnn_option_get_##name(const ncnn_option_t opt)                  \
    {                                                                    \
        return ((const Option*)opt)->name;                               \
    }                                                                    \
    void ncnn_option_set_##name(ncnn_option_t opt, int enable)           \
    {                                                                    \
        ((Option*)opt)->name = enable != 0;                              \
    }
OPT_FLAG(use_vulkan_compute)
OPT_FLAG(use_local_pool_allocator)
OPT_FLAG(use_winograd_convolution)
OPT_FLAG(use_sgemm_convolution)
OPT_FLAG(use_packing_layout)
OPT_FLAG(use_fp16_packed)
OPT_FLAG(use_fp16_storage)
OPT_FLAG(use_fp16_arithmetic)
OPT_FLAG(use_int8_packed)
OPT_FLAG(use_int8_storage)
OPT_FLAG(use_int8_arithmetic)
OPT_FLAG(use_bf16_packed)
OPT_FLAG(use_bf16_storage)
OPT_FLAG(use_cuda_compute)
#undef OPT_FLAG

void ncnn_option_set_lightmode(ncnn_option_t opt, int enable)
{
    ((Option*)opt)->lightmode = enable != 0;
}
void ncnn_option_set_use_cuda_graph_fusion(ncnn_option_t opt, int enable)
{
    ((Option*)opt)->use_cuda_graph_fusion = enable != 0;
}
void ncnn_option_set_use_mapped_model_loading(ncnn_option_t opt, int enable)
{
    ((Option*)opt)->use_mapped_model_loading = enable != 0;
}

// ------------------------------------------------------------------ mat
ncnn_mat_t ncnn_mat_create(void)
{
    return (ncnn_mat_t)(new Mat);
}
ncnn_mat_t ncnn_mat_create_1d(int w, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_2d(int w, int h, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_3d(int w, int h, int c, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, c, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_4d(int w, int h, int d, int c, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, d, c, (size_t)4u, alloc_of(a)));
}
static ncnn_mat_t create_batch(int dims, int w, int h, int d, int c, int n, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create_dims(dims, w, h, d, c, n, 4u, alloc_of(a));
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_1d_batch(int w, int n, ncnn_allocator_t a)
{
    return create_batch(1, w, 1, 1, 1, n, a);
}
ncnn_mat_t ncnn_mat_create_2d_batch(int w, int h, int n, ncnn_allocator_t a)
{
    return create_batch(2, w, h, 1, 1, n, a);
}
ncnn_mat_t ncnn_mat_create_3d_batch(int w, int h, int c, int n, ncnn_allocator_t a)
{
    return create_batch(3, w, h, 1, c, n, a);
}
ncnn_mat_t ncnn_mat_create_4d_batch(int w, int h, int d, int c, int n, ncnn_allocator_t a)
{
    return create_batch(4, w, h, d, c, n, a);
}
ncnn_mat_t ncnn_mat_create_external_1d(int w, void* data, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, data, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_external_2d(int w, int h, void* data, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, data, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_external_3d(int w, int h, int c, void* data, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, c, data, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_external_4d(int w, int h, int d, int c, void* data, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(w, h, d, c, data, (size_t)4u, alloc_of(a)));
}
ncnn_mat_t ncnn_mat_create_1d_elem(int w, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, elemsize, elempack, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_2d_elem(int w, int h, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, elemsize, elempack, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_3d_elem(int w, int h, int c, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, c, elemsize, elempack, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_4d_elem(int w, int h, int d, int c, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, d, c, elemsize, elempack, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_1d_elem_batch(int w, size_t elemsize, int elempack, int n, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, elemsize, elempack, n, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_2d_elem_batch(int w, int h, size_t elemsize, int elempack, int n, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, elemsize, elempack, n, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_3d_elem_batch(int w, int h, int c, size_t elemsize, int elempack, int n, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, c, elemsize, elempack, n, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_4d_elem_batch(int w, int h, int d, int c, size_t elemsize, int elempack, int n, ncnn_allocator_t a)
{
    Mat* m = new Mat;
    m->create(w, h, d, c, elemsize, elempack, n, alloc_of(a));
    m->elempack = elempack; // the layout depends on elemsize only (src/mat.cpp:299-861); the field is carried for the caller
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_external_1d_elem(int w, void* data, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat(w, data, elemsize, alloc_of(a));
    m->elempack = elempack;
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_external_2d_elem(int w, int h, void* data, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat(w, h, data, elemsize, alloc_of(a));
    m->elempack = elempack;
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_external_3d_elem(int w, int h, int c, void* data, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat(w, h, c, data, elemsize, alloc_of(a));
    m->elempack = elempack;
    return (ncnn_mat_t)m;
}
ncnn_mat_t ncnn_mat_create_external_4d_elem(int w, int h, int d, int c, void* data, size_t elemsize, int elempack, ncnn_allocator_t a)
{
    Mat* m = new Mat(w, h, d, c, data, elemsize, alloc_of(a));
    m->elempack = elempack;
    return (ncnn_mat_t)m;
}
void ncnn_mat_destroy(ncnn_mat_t mat)
{
    delete (Mat*)mat;
}
ncnn_mat_t ncnn_mat_from_pixels(const unsigned char* pixels, int type, int w, int h, int stride, ncnn_allocator_t allocator)
{
    return (ncnn_mat_t)(new Mat(Mat::from_pixels(pixels, type, w, h, stride, alloc_of(allocator))));
}
void ncnn_mat_substract_mean_normalize(ncnn_mat_t mat, const float* mean_vals, const float* norm_vals)
{
    ((Mat*)mat)->substract_mean_normalize(mean_vals, norm_vals);
}
void ncnn_mat_to_pixels(const ncnn_mat_t mat, unsigned char* pixels, int type, int stride)
{
    ((const Mat*)mat)->to_pixels(pixels, type, stride);
}
ncnn_mat_t ncnn_mat_from_pixels_resize(const unsigned char* pixels, int type, int w, int h, int stride, int target_width, int target_height, ncnn_allocator_t allocator)
{
    return (ncnn_mat_t)(new Mat(Mat::from_pixels_resize(pixels, type, w, h, stride, target_width, target_height, alloc_of(allocator))));
}
ncnn_mat_t ncnn_mat_from_pixels_roi(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, ncnn_allocator_t allocator)
{
    return (ncnn_mat_t)(new Mat(Mat::from_pixels_roi(pixels, type, w, h, stride, roix, roiy, roiw, roih, alloc_of(allocator))));
}
ncnn_mat_t ncnn_mat_from_pixels_roi_resize(const unsigned char* pixels, int type, int w, int h, int stride, int roix, int roiy, int roiw, int roih, int target_width,
                                           int target_height, ncnn_allocator_t allocator)
{
    return (ncnn_mat_t)(new Mat(Mat::from_pixels_roi_resize(pixels, type, w, h, stride, roix, roiy, roiw, roih, target_width, target_height, alloc_of(allocator))));
}
void ncnn_mat_to_pixels_resize(const ncnn_mat_t mat, unsigned char* pixels, int type, int target_width, int target_height, int target_stride)
{
    ((const Mat*)mat)->to_pixels_resize(pixels, type, target_width, target_height, target_stride);
}
void ncnn_mat_fill_float(ncnn_mat_t mat, float v)
{
    ((Mat*)mat)->fill(v);
}
ncnn_mat_t ncnn_mat_clone(const ncnn_mat_t mat, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(((const Mat*)mat)->clone(alloc_of(a))));
}
ncnn_mat_t ncnn_mat_reshape_1d(const ncnn_mat_t mat, int w, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(((const Mat*)mat)->reshape(w, alloc_of(a))));
}
ncnn_mat_t ncnn_mat_reshape_2d(const ncnn_mat_t mat, int w, int h, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(((const Mat*)mat)->reshape(w, h, alloc_of(a))));
}
ncnn_mat_t ncnn_mat_reshape_3d(const ncnn_mat_t mat, int w, int h, int c, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(((const Mat*)mat)->reshape(w, h, c, alloc_of(a))));
}
ncnn_mat_t ncnn_mat_reshape_4d(const ncnn_mat_t mat, int w, int h, int d, int c, ncnn_allocator_t a)
{
    return (ncnn_mat_t)(new Mat(((const Mat*)mat)->reshape(w, h, d, c, alloc_of(a))));
}
int ncnn_mat_get_dims(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->dims;
}
int ncnn_mat_get_w(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->w;
}
int ncnn_mat_get_h(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->h;
}
int ncnn_mat_get_d(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->d;
}
int ncnn_mat_get_c(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->c;
}
int ncnn_mat_get_n(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->n;
}
size_t ncnn_mat_get_elemsize(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->elemsize;
}
int ncnn_mat_get_elempack(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->elempack;
}
size_t ncnn_mat_get_cstep(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->cstep;
}
size_t ncnn_mat_get_nstep(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->nstep;
}
void* ncnn_mat_get_data(const ncnn_mat_t mat)
{
    return ((const Mat*)mat)->data;
}
void* ncnn_mat_get_batch_data(const ncnn_mat_t mat, int b)
{
    const Mat* m = (const Mat*)mat;
    return (unsigned char*)m->data + m->nstep * b * m->elemsize;
}
ncnn_mat_t ncnn_mat_batch_range(const ncnn_mat_t mat, int b, int batches)
{
    const Mat* m = (const Mat*)mat;
    if (!m || b < 0 || batches <= 0 || b + batches > (m->n > 0 ? m->n : 1)) return 0;
    return (ncnn_mat_t)(new Mat(m->batch_range(b, batches)));
}
void* ncnn_mat_get_channel_data(const ncnn_mat_t mat, int c)
{
    const Mat* m = (const Mat*)mat;
    return (unsigned char*)m->data + m->cstep * c * m->elemsize;
}

// ------------------------------------------------------------------ paramdict
ncnn_paramdict_t ncnn_paramdict_create(void)
{
    return (ncnn_paramdict_t)(new ParamDict);
}
void ncnn_paramdict_destroy(ncnn_paramdict_t pd)
{
    delete (ParamDict*)pd;
}
int ncnn_paramdict_get_type(const ncnn_paramdict_t pd, int id)
{
    return ((const ParamDict*)pd)->type(id);
}
int ncnn_paramdict_get_int(const ncnn_paramdict_t pd, int id, int def)
{
    return ((const ParamDict*)pd)->get(id, def);
}
float ncnn_paramdict_get_float(const ncnn_paramdict_t pd, int id, float def)
{
    return ((const ParamDict*)pd)->get(id, def);
}
ncnn_mat_t ncnn_paramdict_get_array(const ncnn_paramdict_t pd, int id, const ncnn_mat_t def)
{
    return (ncnn_mat_t)(new Mat(((const ParamDict*)pd)->get(id, *(const Mat*)def)));
}
void ncnn_paramdict_set_int(ncnn_paramdict_t pd, int id, int i)
{
    ((ParamDict*)pd)->set(id, i);
}
void ncnn_paramdict_set_float(ncnn_paramdict_t pd, int id, float f)
{
    ((ParamDict*)pd)->set(id, f);
}
void ncnn_paramdict_set_array(ncnn_paramdict_t pd, int id, const ncnn_mat_t v)
{
    ((ParamDict*)pd)->set(id, *(const Mat*)v);
}

// ------------------------------------------------------------------ datareader
namespace {
class DataReader_c_api : public DataReader
{
public:
    explicit DataReader_c_api(ncnn_datareader_t _dr)
        : dr(_dr)
    {
    }
    virtual int scan(const char* format, void* p) const
    {
        return dr->scan ? dr->scan(dr, format, p) : 0;
    }
    virtual size_t read(void* buf, size_t size) const
    {
        return dr->read ? dr->read(dr, buf, size) : 0;
    }
    ncnn_datareader_t dr;
};
struct StdioOrMemory
{
    DataReader* impl;
};
} // namespace

static int dr_scan_none(ncnn_datareader_t, const char*, void*)
{
    return 0;
}
static size_t dr_read_none(ncnn_datareader_t, void*, size_t)
{
    return 0;
}
static int dr_scan_impl(ncnn_datareader_t dr, const char* format, void* p)
{
    return ((DataReader*)dr->pthis)->scan(format, p);
}
static size_t dr_read_impl(ncnn_datareader_t dr, void* buf, size_t size)
{
    return ((DataReader*)dr->pthis)->read(buf, size);
}

ncnn_datareader_t ncnn_datareader_create(void)
{
    ncnn_datareader_t dr = (ncnn_datareader_t)malloc(sizeof(struct __ncnn_datareader_t));
    dr->pthis = 0;
    dr->scan = dr_scan_none;
    dr->read = dr_read_none;
    return dr;
}
ncnn_datareader_t ncnn_datareader_create_from_stdio(FILE* fp)
{
    ncnn_datareader_t dr = (ncnn_datareader_t)malloc(sizeof(struct __ncnn_datareader_t));
    dr->pthis = new DataReaderFromStdio(fp);
    dr->scan = dr_scan_impl;
    dr->read = dr_read_impl;
    return dr;
}
ncnn_datareader_t ncnn_datareader_create_from_memory(const unsigned char** mem)
{
    ncnn_datareader_t dr = (ncnn_datareader_t)malloc(sizeof(struct __ncnn_datareader_t));
    dr->pthis = new DataReaderFromMemory(*mem);
    dr->scan = dr_scan_impl;
    dr->read = dr_read_impl;
    return dr;
}
void ncnn_datareader_destroy(ncnn_datareader_t dr)
{
    if (!dr) return;
    delete (DataReader*)dr->pthis;
    free(dr);
}

// ------------------------------------------------------------------ modelbin
namespace {
class ModelBinOwned : public ModelBin
{
public:
    ModelBinOwned()
        : reader(0), from_reader(0), from_array(0)
    {
    }
    ~ModelBinOwned()
    {
        delete from_reader;
        delete from_array;
        delete reader;
    }
    virtual Mat load(int w, int type) const
    {
        return from_reader ? from_reader->load(w, type) : from_array->load(w, type);
    }
    DataReader_c_api* reader;
    ModelBinFromDataReader* from_reader;
    std::vector<Mat> mats;
    ModelBinFromMatArray* from_array;
};
// a ModelBin whose loads call back into a C table (custom modelbins handed to layers)
class ModelBin_c_api : public ModelBin
{
public:
    explicit ModelBin_c_api(ncnn_modelbin_t _mb)
        : mb(_mb)
    {
    }
    virtual Mat load(int w, int type) const
    {
        ncnn_mat_t m = mb->load_1d(mb, w, type);
        Mat r = *(Mat*)m;
        ncnn_mat_destroy(m);
        return r;
    }
    ncnn_modelbin_t mb;
};
} // namespace

static ncnn_mat_t mb_load_1d(const ncnn_modelbin_t mb, int w, int type)
{
    return (ncnn_mat_t)(new Mat(((const ModelBin*)mb->pthis)->load(w, type)));
}
static ncnn_mat_t mb_load_2d(const ncnn_modelbin_t mb, int w, int h, int type)
{
    return (ncnn_mat_t)(new Mat(((const ModelBin*)mb->pthis)->load(w, h, type)));
}
static ncnn_mat_t mb_load_3d(const ncnn_modelbin_t mb, int w, int h, int c, int type)
{
    return (ncnn_mat_t)(new Mat(((const ModelBin*)mb->pthis)->load(w, h, c, type)));
}

static ncnn_modelbin_t wrap_modelbin(ModelBin* impl)
{
    ncnn_modelbin_t mb = (ncnn_modelbin_t)malloc(sizeof(struct __ncnn_modelbin_t));
    mb->pthis = impl;
    mb->load_1d = mb_load_1d;
    mb->load_2d = mb_load_2d;
    mb->load_3d = mb_load_3d;
    return mb;
}

ncnn_modelbin_t ncnn_modelbin_create_from_datareader(const ncnn_datareader_t dr)
{
    ModelBinOwned* o = new ModelBinOwned;
    o->reader = new DataReader_c_api(dr);
    o->from_reader = new ModelBinFromDataReader(*o->reader);
    return wrap_modelbin(o);
}

ncnn_modelbin_t ncnn_modelbin_create_from_mat_array(const ncnn_mat_t* weights, int n)
{
    ModelBinOwned* o = new ModelBinOwned;
    o->mats.resize(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) o->mats[i] = *(const Mat*)weights[i];
    o->from_array = new ModelBinFromMatArray(&o->mats[0]);
    return wrap_modelbin(o);
}

void ncnn_modelbin_destroy(ncnn_modelbin_t mb)
{
    if (!mb) return;
    delete (ModelBin*)mb->pthis;
    free(mb);
}

// ------------------------------------------------------------------ layer
static int layer_load_param(ncnn_layer_t layer, const ncnn_paramdict_t pd)
{
    return ((Layer*)layer->pthis)->load_param(*(const ParamDict*)pd);
}
static int layer_load_model(ncnn_layer_t layer, const ncnn_modelbin_t mb)
{
    return ((Layer*)layer->pthis)->load_model(*(const ModelBin*)mb->pthis);
}
static int layer_create_pipeline(ncnn_layer_t layer, const ncnn_option_t opt)
{
    return ((Layer*)layer->pthis)->create_pipeline(*(const Option*)opt);
}
static int layer_destroy_pipeline(ncnn_layer_t layer, const ncnn_option_t opt)
{
    return ((Layer*)layer->pthis)->destroy_pipeline(*(const Option*)opt);
}
static int layer_forward_1(const ncnn_layer_t layer, const ncnn_mat_t bottom_blob, ncnn_mat_t* top_blob, const ncnn_option_t opt)
{
    Mat top;
    int ret = ((const Layer*)layer->pthis)->forward(*(const Mat*)bottom_blob, top, *(const Option*)opt);
    *top_blob = (ncnn_mat_t)(new Mat(top));
    return ret;
}
static int layer_forward_n(const ncnn_layer_t layer, const ncnn_mat_t* bottom_blobs, int n, ncnn_mat_t* top_blobs, int n2, const ncnn_option_t opt)
{
    std::vector<Mat> b(n), t(n2);
    for (int i = 0; i < n; i++) b[i] = *(const Mat*)bottom_blobs[i];
    int ret = ((const Layer*)layer->pthis)->forward(b, t, *(const Option*)opt);
    for (int i = 0; i < n2; i++) top_blobs[i] = (ncnn_mat_t)(new Mat(i < (int)t.size() ? t[i] : Mat()));
    return ret;
}
static int layer_forward_inplace_1(const ncnn_layer_t layer, ncnn_mat_t bottom_top_blob, const ncnn_option_t opt)
{
    return ((const Layer*)layer->pthis)->forward_inplace(*(Mat*)bottom_top_blob, *(const Option*)opt);
}
static int layer_forward_inplace_n(const ncnn_layer_t layer, ncnn_mat_t* bottom_top_blobs, int n, const ncnn_option_t opt)
{
    std::vector<Mat> b(n);
    for (int i = 0; i < n; i++) b[i] = *(Mat*)bottom_top_blobs[i];
    int ret = ((const Layer*)layer->pthis)->forward_inplace(b, *(const Option*)opt);
    for (int i = 0; i < n; i++) *(Mat*)bottom_top_blobs[i] = b[i];
    return ret;
}

static ncnn_layer_t wrap_layer(Layer* impl)
{
    if (!impl) return 0;
    ncnn_layer_t layer = (ncnn_layer_t)malloc(sizeof(struct __ncnn_layer_t));
    layer->pthis = impl;
    layer->load_param = layer_load_param;
    layer->load_model = layer_load_model;
    layer->create_pipeline = layer_create_pipeline;
    layer->destroy_pipeline = layer_destroy_pipeline;
    layer->forward_1 = layer_forward_1;
    layer->forward_n = layer_forward_n;
    layer->forward_inplace_1 = layer_forward_inplace_1;
    layer->forward_inplace_n = layer_forward_inplace_n;
    return layer;
}

namespace {
// a Layer whose behaviour lives in a C function-pointer table (src/c_api.cpp Layer_c_api): host Mats only, so inside a
// CUDA graph it costs a download + upload, like a CPU layer inside the reference's Vulkan graph (src/net.cpp:229-247)
class Layer_c_api : public Layer
{
public:
    explicit Layer_c_api(ncnn_layer_t _layer)
        : layer(_layer)
    {
    }
    virtual int load_param(const ParamDict& pd)
    {
        return layer->load_param(layer, (ncnn_paramdict_t)&pd);
    }
    virtual int load_model(const ModelBin& mb)
    {
        struct __ncnn_modelbin_t mb0;
        mb0.pthis = (void*)&mb;
        mb0.load_1d = mb_load_1d;
        mb0.load_2d = mb_load_2d;
        mb0.load_3d = mb_load_3d;
        return layer->load_model(layer, &mb0);
    }
    virtual int create_pipeline(const Option& opt)
    {
        return layer->create_pipeline(layer, (ncnn_option_t)&opt);
    }
    virtual int destroy_pipeline(const Option& opt)
    {
        return layer->destroy_pipeline(layer, (ncnn_option_t)&opt);
    }
    int host_forward(std::vector<Mat>& b, std::vector<Mat>& t, const Option& opt) const
    {
        if (one_blob_only && support_inplace)
        {
            int ret = layer->forward_inplace_1(layer, (ncnn_mat_t)&b[0], (ncnn_option_t)&opt);
            t.assign(1, b[0]);
            return ret;
        }
        if (one_blob_only)
        {
            ncnn_mat_t top = 0;
            int ret = layer->forward_1(layer, (ncnn_mat_t)&b[0], &top, (ncnn_option_t)&opt);
            t.resize(1);
            if (top)
            {
                t[0] = *(Mat*)top;
                ncnn_mat_destroy(top);
            }
            return ret;
        }
        std::vector<ncnn_mat_t> bp(b.size());
        for (size_t i = 0; i < b.size(); i++) bp[i] = (ncnn_mat_t)&b[i];
        if (support_inplace)
        {
            int ret = layer->forward_inplace_n(layer, &bp[0], (int)bp.size(), (ncnn_option_t)&opt);
            t = b;
            return ret;
        }
        std::vector<ncnn_mat_t> tp(t.size(), (ncnn_mat_t)0);
        int ret = layer->forward_n(layer, &bp[0], (int)bp.size(), tp.empty() ? 0 : &tp[0], (int)tp.size(), (ncnn_option_t)&opt);
        for (size_t i = 0; i < tp.size(); i++)
        {
            if (tp[i])
            {
                t[i] = *(Mat*)tp[i];
                ncnn_mat_destroy(tp[i]);
            }
        }
        return ret;
    }
    virtual int forward(const std::vector<Mat>& bottom_blobs, std::vector<Mat>& top_blobs, const Option& opt) const
    {
        std::vector<Mat> b = bottom_blobs;
        return host_forward(b, top_blobs, opt);
    }
    virtual int forward(const Mat& bottom_blob, Mat& top_blob, const Option& opt) const
    {
        std::vector<Mat> b(1, bottom_blob), t(1);
        int ret = host_forward(b, t, opt);
        top_blob = t[0];
        return ret;
    }
    virtual int forward_inplace(std::vector<Mat>& bottom_top_blobs, const Option& opt) const
    {
        std::vector<Mat> t(bottom_top_blobs.size());
        int ret = host_forward(bottom_top_blobs, t, opt);
        bottom_top_blobs = t;
        return ret;
    }
    virtual int forward_inplace(Mat& bottom_top_blob, const Option& opt) const
    {
        std::vector<Mat> b(1, bottom_top_blob), t(1);
        int ret = host_forward(b, t, opt);
        bottom_top_blob = t[0];
        return ret;
    }
    // device overloads: download, run on the host per sample (a custom layer knows nothing about batches), upload
    virtual int forward(const std::vector<CudaMat>& bottom_blobs, std::vector<CudaMat>& top_blobs, CudaCompute& cmd, const Option& opt) const
    {
        std::vector<Mat> hb(bottom_blobs.size());
        for (size_t i = 0; i < hb.size(); i++)
        {
            int ret = cmd.record_download(bottom_blobs[i], hb[i], opt);
            if (ret != 0) return ret;
        }
        int ret = cmd.submit_and_wait();
        if (ret != 0) return ret;
        const int n = hb.empty() ? 1 : (hb[0].n < 1 ? 1 : hb[0].n);
        std::vector<Mat> ht(top_blobs.size());
        for (int s = 0; s < n; s++)
        {
            std::vector<Mat> sb(hb.size()), st(top_blobs.size());
            for (size_t i = 0; i < hb.size(); i++) sb[i] = n > 1 ? hb[i].batch(s).clone() : hb[i];
            ret = host_forward(sb, st, opt);
            if (ret != 0) return ret;
            for (size_t i = 0; i < st.size(); i++)
            {
                if (n == 1)
                {
                    ht[i] = st[i];
                    continue;
                }
                if (s == 0) ht[i].create_like(st[i], n, opt.blob_allocator);
                if (ht[i].empty()) return -100;
                memcpy((unsigned char*)ht[i].data + ht[i].nstep * s * ht[i].elemsize, st[i].data, st[i].total() * st[i].elemsize);
            }
        }
        for (size_t i = 0; i < ht.size(); i++)
        {
            ret = cmd.record_upload(ht[i], top_blobs[i], opt);
            if (ret != 0) return ret;
        }
        return cmd.submit_and_wait();
    }
    virtual int forward(const CudaMat& bottom_blob, CudaMat& top_blob, CudaCompute& cmd, const Option& opt) const
    {
        std::vector<CudaMat> b(1, bottom_blob), t(1);
        int ret = forward(b, t, cmd, opt);
        top_blob = t[0];
        return ret;
    }
    virtual int forward_inplace(std::vector<CudaMat>& bottom_top_blobs, CudaCompute& cmd, const Option& opt) const
    {
        std::vector<CudaMat> t(bottom_top_blobs.size());
        int ret = forward(bottom_top_blobs, t, cmd, opt);
        bottom_top_blobs = t;
        return ret;
    }
    virtual int forward_inplace(CudaMat& bottom_top_blob, CudaCompute& cmd, const Option& opt) const
    {
        std::vector<CudaMat> b(1, bottom_top_blob), t(1);
        int ret = forward(b, t, cmd, opt);
        bottom_top_blob = t[0];
        return ret;
    }
    ncnn_layer_t layer;
};

struct custom_layer_factory_t
{
    ncnn_layer_creator_t creator;
    ncnn_layer_destroyer_t destroyer;
    void* userdata;
    custom_layer_factory_t* next;
};
} // namespace

static int base_load_param(ncnn_layer_t, const ncnn_paramdict_t)
{
    return 0;
}
static int base_load_model(ncnn_layer_t, const ncnn_modelbin_t)
{
    return 0;
}
static int base_pipeline(ncnn_layer_t, const ncnn_option_t)
{
    return 0;
}
static int base_forward_1(const ncnn_layer_t, const ncnn_mat_t, ncnn_mat_t*, const ncnn_option_t)
{
    return -1;
}
static int base_forward_n(const ncnn_layer_t, const ncnn_mat_t*, int, ncnn_mat_t*, int, const ncnn_option_t)
{
    return -1;
}
static int base_forward_inplace_1(const ncnn_layer_t, ncnn_mat_t, const ncnn_option_t)
{
    return -1;
}
static int base_forward_inplace_n(const ncnn_layer_t, ncnn_mat_t*, int, const ncnn_option_t)
{
    return -1;
}

ncnn_layer_t ncnn_layer_create(void)
{
    // an empty table for a custom operator to fill in; pthis is the Layer the Net will own
    ncnn_layer_t layer = (ncnn_layer_t)malloc(sizeof(struct __ncnn_layer_t));
    layer->load_param = base_load_param;
    layer->load_model = base_load_model;
    layer->create_pipeline = base_pipeline;
    layer->destroy_pipeline = base_pipeline;
    layer->forward_1 = base_forward_1;
    layer->forward_n = base_forward_n;
    layer->forward_inplace_1 = base_forward_inplace_1;
    layer->forward_inplace_n = base_forward_inplace_n;
    layer->pthis = new Layer_c_api(layer);
    return layer;
}
ncnn_layer_t ncnn_layer_create_by_typeindex(int typeindex)
{
    return wrap_layer(create_layer(typeindex));
}
ncnn_layer_t ncnn_layer_create_by_type(const char* type)
{
    return wrap_layer(create_layer(type));
}
int ncnn_layer_type_to_index(const char* type)
{
    return layer_to_index(type);
}
void ncnn_layer_destroy(ncnn_layer_t layer)
{
    if (!layer) return;
    delete (Layer*)layer->pthis;
    free(layer);
}
const char* ncnn_layer_get_name(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->name.c_str();
}
int ncnn_layer_get_typeindex(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->typeindex;
}
const char* ncnn_layer_get_type(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->type.c_str();
}
int ncnn_layer_get_one_blob_only(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->one_blob_only;
}
int ncnn_layer_get_support_inplace(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->support_inplace;
}
int ncnn_layer_get_support_vulkan(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->support_vulkan;
}
int ncnn_layer_get_support_packing(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->support_packing;
}
int ncnn_layer_get_support_bf16_storage(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->support_bf16_storage;
}
int ncnn_layer_get_support_fp16_storage(const ncnn_layer_t layer)
{
    return ((const Layer*)layer->pthis)->support_fp16_storage;
}
void ncnn_layer_set_one_blob_only(ncnn_layer_t layer, int enable)
{
    ((Layer*)layer->pthis)->one_blob_only = enable != 0;
}
void ncnn_layer_set_support_inplace(ncnn_layer_t layer, int enable)
{
    ((Layer*)layer->pthis)->support_inplace = enable != 0;
}
int ncnn_layer_get_support_vulkan_packing(const ncnn_layer_t)
{
    return 0;
}
int ncnn_layer_get_support_any_packing(const ncnn_layer_t)
{
    return 0;
}
int ncnn_layer_get_support_vulkan_any_packing(const ncnn_layer_t)
{
    return 0;
}
void ncnn_layer_set_support_vulkan(ncnn_layer_t, int)
{
}
void ncnn_layer_set_support_packing(ncnn_layer_t, int)
{
}
void ncnn_layer_set_support_bf16_storage(ncnn_layer_t layer, int enable)
{
    ((Layer*)layer->pthis)->support_bf16_storage = enable != 0;
}
void ncnn_layer_set_support_fp16_storage(ncnn_layer_t layer, int enable)
{
    ((Layer*)layer->pthis)->support_fp16_storage = enable != 0;
}
void ncnn_layer_set_support_vulkan_packing(ncnn_layer_t, int)
{
}
void ncnn_layer_set_support_any_packing(ncnn_layer_t, int)
{
}
void ncnn_layer_set_support_vulkan_any_packing(ncnn_layer_t, int)
{
}
static void shape_of(const std::vector<Mat>& shapes, int i, int* dims, int* w, int* h, int* c)
{
    *dims = *w = *h = *c = 0;
    if (i < 0 || i >= (int)shapes.size()) return;
    const Mat& shape = shapes[i];
    *dims = shape.dims;
    *w = shape.w;
    *h = shape.h;
    *c = shape.c;
}
void ncnn_blob_get_bottom_shape(const ncnn_layer_t layer, int i, int* dims, int* w, int* h, int* c)
{
    shape_of(((const Layer*)layer->pthis)->bottom_shapes, i, dims, w, h, c);
}
void ncnn_blob_get_top_shape(const ncnn_layer_t layer, int i, int* dims, int* w, int* h, int* c)
{
    shape_of(((const Layer*)layer->pthis)->top_shapes, i, dims, w, h, c);
}
int ncnn_layer_get_bottom_count(const ncnn_layer_t layer)
{
    return (int)((const Layer*)layer->pthis)->bottoms.size();
}
int ncnn_layer_get_bottom(const ncnn_layer_t layer, int i)
{
    return ((const Layer*)layer->pthis)->bottoms[i];
}
int ncnn_layer_get_top_count(const ncnn_layer_t layer)
{
    return (int)((const Layer*)layer->pthis)->tops.size();
}
int ncnn_layer_get_top(const ncnn_layer_t layer, int i)
{
    return ((const Layer*)layer->pthis)->tops[i];
}

// ------------------------------------------------------------------ net
ncnn_net_t ncnn_net_create(void)
{
    ncnn_net_t net = (ncnn_net_t)malloc(sizeof(struct __ncnn_net_t));
    net->pthis = new Net;
    net->custom_layer_factory = 0;
    return net;
}

void ncnn_net_destroy(ncnn_net_t net)
{
    if (!net) return;
    delete (Net*)net->pthis;
    custom_layer_factory_t* f = (custom_layer_factory_t*)net->custom_layer_factory;
    while (f)
    {
        custom_layer_factory_t* nx = f->next;
        free(f);
        f = nx;
    }
    free(net);
}

ncnn_option_t ncnn_net_get_option(ncnn_net_t net)
{
    return (ncnn_option_t)(&((Net*)net->pthis)->opt);
}
void ncnn_net_set_option(ncnn_net_t net, ncnn_option_t opt)
{
    ((Net*)net->pthis)->opt = *(Option*)opt;
}
void ncnn_net_set_cuda_device(ncnn_net_t net, int device_index)
{
    ((Net*)net->pthis)->set_cuda_device(device_index);
}
int ncnn_net_get_fused_layer_count(const ncnn_net_t net)
{
    return ((const Net*)net->pthis)->fused_layer_count();
}
int ncnn_get_cuda_device_count(void)
{
    return get_cuda_device_count();
}

static Layer* c_api_layer_creator(void* userdata)
{
    custom_layer_factory_t* f = (custom_layer_factory_t*)userdata;
    ncnn_layer_t layer0 = f->creator(f->userdata);
    if (!layer0) return 0;
    // the table the user returned drives a Layer_c_api (ncnn_layer_create already made one behind pthis)
    return (Layer*)layer0->pthis;
}

static void c_api_layer_destroyer(Layer* layer, void* userdata)
{
    custom_layer_factory_t* f = (custom_layer_factory_t*)userdata;
    ncnn_layer_t layer0 = ((Layer_c_api*)layer)->layer;
    if (f->destroyer)
        f->destroyer(layer0, f->userdata);
    else
        ncnn_layer_destroy(layer0);
}

void ncnn_net_register_custom_layer_by_type(ncnn_net_t net, const char* type, ncnn_layer_creator_t creator, ncnn_layer_destroyer_t destroyer, void* userdata)
{
    custom_layer_factory_t* f = (custom_layer_factory_t*)malloc(sizeof(custom_layer_factory_t));
    f->creator = creator;
    f->destroyer = destroyer;
    f->userdata = userdata;
    f->next = (custom_layer_factory_t*)net->custom_layer_factory;
    net->custom_layer_factory = f;
    ((Net*)net->pthis)->register_custom_layer(type, c_api_layer_creator, c_api_layer_destroyer, f);
}

void ncnn_net_register_custom_layer_by_typeindex(ncnn_net_t net, int typeindex, ncnn_layer_creator_t creator, ncnn_layer_destroyer_t destroyer, void* userdata)
{
    const char* type = layer_index_to_type(typeindex);
    if (type) ncnn_net_register_custom_layer_by_type(net, type, creator, destroyer, userdata);
}

int ncnn_net_load_param(ncnn_net_t net, const char* path)
{
    return ((Net*)net->pthis)->load_param(path);
}
int ncnn_net_load_param_bin(ncnn_net_t net, const char* path)
{
    return ((Net*)net->pthis)->load_param_bin(path);
}
int ncnn_net_load_model(ncnn_net_t net, const char* path)
{
    return ((Net*)net->pthis)->load_model(path);
}
int ncnn_net_load_param_memory(ncnn_net_t net, const char* mem)
{
    return ((Net*)net->pthis)->load_param_mem(mem);
}
size_t ncnn_net_load_param_bin_memory(ncnn_net_t net, const unsigned char* mem)
{
    return ((Net*)net->pthis)->load_param_bin_mem(mem);
}
size_t ncnn_net_load_model_memory(ncnn_net_t net, const unsigned char* mem)
{
    return ((Net*)net->pthis)->load_model(mem);
}
int ncnn_net_load_param_datareader(ncnn_net_t net, const ncnn_datareader_t dr)
{
    DataReader_c_api r(dr);
    return ((Net*)net->pthis)->load_param(r);
}
int ncnn_net_load_param_bin_datareader(ncnn_net_t net, const ncnn_datareader_t dr)
{
    DataReader_c_api r(dr);
    return ((Net*)net->pthis)->load_param_bin(r);
}
int ncnn_net_load_model_datareader(ncnn_net_t net, const ncnn_datareader_t dr)
{
    DataReader_c_api r(dr);
    return ((Net*)net->pthis)->load_model(r);
}
void ncnn_net_clear(ncnn_net_t net)
{
    ((Net*)net->pthis)->clear();
}
int ncnn_net_get_input_count(const ncnn_net_t net)
{
    return (int)((const Net*)net->pthis)->input_indexes().size();
}
int ncnn_net_get_output_count(const ncnn_net_t net)
{
    return (int)((const Net*)net->pthis)->output_indexes().size();
}
const char* ncnn_net_get_input_name(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->input_names()[i];
}
const char* ncnn_net_get_output_name(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->output_names()[i];
}
int ncnn_net_get_input_index(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->input_indexes()[i];
}
int ncnn_net_get_output_index(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->output_indexes()[i];
}

// ------------------------------------------------------------------ extractor
ncnn_extractor_t ncnn_extractor_create(ncnn_net_t net)
{
    return (ncnn_extractor_t)(new Extractor(((Net*)net->pthis)->create_extractor()));
}
void ncnn_extractor_destroy(ncnn_extractor_t ex)
{
    delete (Extractor*)ex;
}
void ncnn_extractor_set_option(ncnn_extractor_t ex, const ncnn_option_t opt)
{
    Extractor* e = (Extractor*)ex;
    const Option* o = (const Option*)opt;
    e->set_light_mode(o->lightmode);
    e->set_blob_allocator(o->blob_allocator);
    e->set_workspace_allocator(o->workspace_allocator);
}
int ncnn_extractor_input(ncnn_extractor_t ex, const char* name, const ncnn_mat_t mat)
{
    return ((Extractor*)ex)->input(name, *(const Mat*)mat);
}
int ncnn_extractor_input_pixels(ncnn_extractor_t ex, const char* name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride,
                                const float* mean_vals, const float* norm_vals)
{
    return ((Extractor*)ex)->input_pixels(name, pixels, type, w, h, stride, n, nstride, mean_vals, norm_vals);
}
// ---- host-side Mat helpers (src/mat.cpp copy_make_border / copy_cut_border / flatten run the Padding / Crop / Flatten layers on
// the host; here they are plain loops over fp32 Mats)
static int border_index(int i, int n, int type)
{
    if (i >= 0 && i < n) return i;
    if (type == 1) return i < 0 ? 0 : n - 1;             // replicate
    if (type == 2) return i < 0 ? -i : 2 * (n - 1) - i;  // reflect without repeating the edge (padding.cpp:185-260)
    return -1;                                           // constant
}
static void make_border(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int front, int behind, int type, float v, Allocator* alloc)
{
    if (src.empty() || src.elemsize != 4u || src.dims < 1 || src.dims > 3 || type < 0 || type > 2 || top < 0 || bottom < 0 || left < 0 || right < 0 || front < 0 || behind < 0)
    {
        dst.release();
        return;
    }
    if (src.dims == 1) top = bottom = 0;
    if (src.dims != 3) front = behind = 0;
    if (top == 0 && bottom == 0 && left == 0 && right == 0 && front == 0 && behind == 0)
    {
        dst = src;
        return;
    }
    if (type == 2 && (left >= src.w || right >= src.w || (src.dims >= 2 && (top >= src.h || bottom >= src.h)) || (src.dims == 3 && (front >= src.c || behind >= src.c))))
    {
        dst.release();
        return;
    }
    Mat out;
    const int ow = src.w + left + right, oh = src.h + top + bottom, oc = src.c + front + behind;
    if (src.dims == 1) out.create(ow, (size_t)4u, alloc);
    if (src.dims == 2) out.create(ow, oh, (size_t)4u, alloc);
    if (src.dims == 3) out.create(ow, oh, oc, (size_t)4u, alloc);
    if (out.empty())
    {
        dst.release();
        return;
    }
    const int chs = src.dims == 3 ? oc : 1;
    for (int q = 0; q < chs; q++)
    {
        // channel padding (src/layer/padding.cpp:333-372): a constant plane, or the replicated / reflected source channel
        const int sq = src.dims == 3 ? border_index(q - front, src.c, type) : 0;
        float* dp = src.dims == 3 ? (float*)out.channel(q) : (float*)out.data;
        const int rows = src.dims == 1 ? 1 : oh;
        if (sq < 0)
        {
            for (size_t i = 0; i < (size_t)rows * ow; i++) dp[i] = v;
            continue;
        }
        const float* sp = src.dims == 3 ? (const float*)src.channel(sq) : (const float*)src.data;
        for (int y = 0; y < rows; y++)
        {
            const int sy = src.dims == 1 ? 0 : border_index(y - top, src.h, type);
            for (int x = 0; x < ow; x++)
            {
                const int sx = border_index(x - left, src.w, type);
                dp[(size_t)y * ow + x] = (sx < 0 || sy < 0) ? v : sp[(size_t)sy * src.w + sx];
            }
        }
    }
    dst = out;
}
void ncnn_copy_make_border(const ncnn_mat_t _src, ncnn_mat_t _dst, int top, int bottom, int left, int right, int type, float v, const ncnn_option_t opt)
{
    make_border(*(const Mat*)_src, *(Mat*)_dst, top, bottom, left, right, 0, 0, type, v, opt ? ((const Option*)opt)->blob_allocator : 0);
}
void ncnn_copy_make_border_3d(const ncnn_mat_t _src, ncnn_mat_t _dst, int top, int bottom, int left, int right, int front, int behind, int type, float v,
                              const ncnn_option_t opt)
{
    make_border(*(const Mat*)_src, *(Mat*)_dst, top, bottom, left, right, front, behind, type, v, opt ? ((const Option*)opt)->blob_allocator : 0);
}
static void cut_border(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int front, int behind, Allocator* alloc)
{
    if (src.empty() || src.elemsize != 4u || src.dims < 1 || src.dims > 3 || top < 0 || bottom < 0 || left < 0 || right < 0 || front < 0 || behind < 0)
    {
        dst.release();
        return;
    }
    if (src.dims == 1) top = bottom = 0;
    if (src.dims != 3) front = behind = 0;
    const int ow = src.w - left - right, oh = src.h - top - bottom, oc = src.c - front - behind;
    if (ow <= 0 || oh <= 0 || oc <= 0)
    {
        dst.release();
        return;
    }
    if (ow == src.w && oh == src.h && oc == src.c)
    {
        dst = src;
        return;
    }
    Mat out;
    if (src.dims == 1) out.create(ow, (size_t)4u, alloc);
    if (src.dims == 2) out.create(ow, oh, (size_t)4u, alloc);
    if (src.dims == 3) out.create(ow, oh, oc, (size_t)4u, alloc);
    if (out.empty())
    {
        dst.release();
        return;
    }
    const int chs = src.dims == 3 ? oc : 1;
    for (int q = 0; q < chs; q++)
    {
        const float* sp = src.dims == 3 ? (const float*)src.channel(q + front) : (const float*)src.data;
        float* dp = src.dims == 3 ? (float*)out.channel(q) : (float*)out.data;
        for (int y = 0; y < oh; y++) memcpy(dp + (size_t)y * ow, sp + (size_t)(y + top) * src.w + left, (size_t)ow * sizeof(float));
    }
    dst = out;
}
void ncnn_copy_cut_border(const ncnn_mat_t _src, ncnn_mat_t _dst, int top, int bottom, int left, int right, const ncnn_option_t opt)
{
    cut_border(*(const Mat*)_src, *(Mat*)_dst, top, bottom, left, right, 0, 0, opt ? ((const Option*)opt)->blob_allocator : 0);
}
void ncnn_flatten(const ncnn_mat_t _src, ncnn_mat_t* _dst, const ncnn_option_t opt)
{
    const Mat& src = *(const Mat*)_src;
    Allocator* alloc = opt ? ((const Option*)opt)->blob_allocator : 0;
    Mat* out = new Mat;
    if (!src.empty() && src.elemsize == 4u)
    {
        const size_t plane = (size_t)src.w * src.h * src.d;
        const int chs = src.dims >= 3 ? src.c : 1;
        out->create((int)(plane * chs), (size_t)4u, alloc);
        if (!out->empty())
            for (int q = 0; q < chs; q++)
                memcpy((float*)out->data + plane * q, src.dims >= 3 ? (const float*)src.channel(q) : (const float*)src.data, plane * sizeof(float));
    }
    *_dst = (ncnn_mat_t)out;
}
int ncnn_extractor_input_pixels_resize(ncnn_extractor_t ex, const char* name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride,
                                       int target_w, int target_h, const float* mean_vals, const float* norm_vals)
{
    return ((Extractor*)ex)->input_pixels_resize(name, pixels, type, w, h, stride, n, nstride, target_w, target_h, mean_vals, norm_vals);
}
int ncnn_extractor_extract_yolov8_proposals(ncnn_extractor_t ex, const char* name, const int* strides, int num_strides, int in_w, int in_h, float prob_threshold,
                                            ncnn_mat_t* proposals)
{
    Mat m;
    int ret = ((Extractor*)ex)->extract_yolov8_proposals(name, strides, num_strides, in_w, in_h, prob_threshold, m);
    *proposals = ret == 0 ? (ncnn_mat_t)(new Mat(m)) : 0;
    return ret;
}
int ncnn_extractor_extract(ncnn_extractor_t ex, const char* name, ncnn_mat_t* mat)
{
    Mat m;
    int ret = ((Extractor*)ex)->extract(name, m);
    *mat = (ncnn_mat_t)(new Mat(m));
    return ret;
}
int ncnn_extractor_input_index(ncnn_extractor_t ex, int index, const ncnn_mat_t mat)
{
    return ((Extractor*)ex)->input(index, *(const Mat*)mat);
}
int ncnn_extractor_extract_index(ncnn_extractor_t ex, int index, ncnn_mat_t* mat)
{
    Mat m;
    int ret = ((Extractor*)ex)->extract(index, m);
    *mat = (ncnn_mat_t)(new Mat(m));
    return ret;
}
size_t ncnn_extractor_get_last_h2d_bytes(const ncnn_extractor_t ex)
{
    return ((const Extractor*)ex)->last_h2d_bytes();
}
size_t ncnn_extractor_get_last_d2h_bytes(const ncnn_extractor_t ex)
{
    return ((const Extractor*)ex)->last_d2h_bytes();
}

// ------------------------------------------------------------------ CUDA additions
namespace {
struct ComputeHolder
{
    CudaContext* ctx;
    CudaCompute* cmd;
};
} // namespace

ncnn_cuda_compute_t ncnn_cuda_compute_create(int device_index)
{
    CudaContext* ctx = acquire_cuda_context(device_index);
    if (!ctx) return 0;
    ComputeHolder* h = new ComputeHolder;
    h->ctx = ctx;
    h->cmd = new CudaCompute(ctx);
    return (ncnn_cuda_compute_t)h;
}
void ncnn_cuda_compute_destroy(ncnn_cuda_compute_t cmd)
{
    if (!cmd) return;
    ComputeHolder* h = (ComputeHolder*)cmd;
    delete h->cmd;
    reclaim_cuda_context(h->ctx);
    delete h;
}
void* ncnn_cuda_compute_get_stream(ncnn_cuda_compute_t cmd)
{
    return ((ComputeHolder*)cmd)->cmd->stream();
}
int ncnn_cuda_compute_record_upload(ncnn_cuda_compute_t cmd, const ncnn_mat_t src, ncnn_cuda_mat_t* dst, const ncnn_option_t opt)
{
    CudaMat* m = new CudaMat;
    int ret = ((ComputeHolder*)cmd)->cmd->record_upload(*(const Mat*)src, *m, *(const Option*)opt);
    *dst = (ncnn_cuda_mat_t)m;
    return ret;
}
int ncnn_cuda_compute_record_download(ncnn_cuda_compute_t cmd, const ncnn_cuda_mat_t src, ncnn_mat_t* dst, const ncnn_option_t opt)
{
    Mat* m = new Mat;
    int ret = ((ComputeHolder*)cmd)->cmd->record_download(*(const CudaMat*)src, *m, *(const Option*)opt);
    *dst = (ncnn_mat_t)m;
    return ret;
}
int ncnn_cuda_compute_submit_and_wait(ncnn_cuda_compute_t cmd)
{
    return ((ComputeHolder*)cmd)->cmd->submit_and_wait();
}
void ncnn_cuda_compute_set_profiling(ncnn_cuda_compute_t cmd, int enable)
{
    ((ComputeHolder*)cmd)->cmd->set_profiling(enable != 0);
}
int ncnn_cuda_compute_get_profile_count(ncnn_cuda_compute_t cmd)
{
    return (int)((ComputeHolder*)cmd)->cmd->timings().size();
}
int ncnn_cuda_compute_get_profile(ncnn_cuda_compute_t cmd, int i, int* layer_index, float* ms, int shape[6])
{
    const std::vector<CudaCompute::LayerTiming>& t = ((ComputeHolder*)cmd)->cmd->timings();
    if (i < 0 || i >= (int)t.size()) return -1;
    *layer_index = t[i].layer_index;
    *ms = t[i].ms;
    shape[0] = t[i].dims;
    shape[1] = t[i].w;
    shape[2] = t[i].h;
    shape[3] = t[i].d;
    shape[4] = t[i].c;
    shape[5] = t[i].n;
    return 0;
}
void ncnn_cuda_compute_clear_profile(ncnn_cuda_compute_t cmd)
{
    ((ComputeHolder*)cmd)->cmd->clear_timings();
}
int ncnn_net_get_layer_count(const ncnn_net_t net)
{
    return (int)((const Net*)net->pthis)->layers().size();
}
const char* ncnn_net_get_layer_type(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->layers()[i]->type.c_str();
}
const char* ncnn_net_get_layer_name(const ncnn_net_t net, int i)
{
    return ((const Net*)net->pthis)->layers()[i]->name.c_str();
}
void ncnn_cuda_mat_destroy(ncnn_cuda_mat_t mat)
{
    delete (CudaMat*)mat;
}
int ncnn_cuda_mat_get_dims(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->dims;
}
int ncnn_cuda_mat_get_w(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->w;
}
int ncnn_cuda_mat_get_h(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->h;
}
int ncnn_cuda_mat_get_c(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->c;
}
int ncnn_cuda_mat_get_n(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->n;
}
int ncnn_cuda_mat_get_elemtype(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->elemtype;
}
void* ncnn_cuda_mat_get_data(const ncnn_cuda_mat_t mat)
{
    return ((const CudaMat*)mat)->data;
}
int ncnn_extractor_input_cuda(ncnn_extractor_t ex, const char* name, const ncnn_cuda_mat_t mat)
{
    return ((Extractor*)ex)->input(name, *(const CudaMat*)mat);
}
int ncnn_extractor_extract_cuda(ncnn_extractor_t ex, const char* name, ncnn_cuda_mat_t* mat, ncnn_cuda_compute_t cmd)
{
    CudaMat* m = new CudaMat;
    int ret = ((Extractor*)ex)->extract(name, *m, *((ComputeHolder*)cmd)->cmd);
    *mat = (ncnn_cuda_mat_t)m;
    return ret;
}

} // extern "C"
