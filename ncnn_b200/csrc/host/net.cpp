// net.cpp -- see net.h.  Control flow mirrors the reference's src/net.cpp (line references per function).
#include "net.h"

#include <ctype.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "layer/cuda_layers.h"
#include "modelbin.h"
#include "paramdict.h"

namespace ncnn {

// a read-only mapping of a model file (the reference's MappedFile, src/net.cpp:55-121)
struct MappedModelFile
{
    MappedModelFile()
        : ptr(0), size(0)
    {
    }
    ~MappedModelFile()
    {
        close();
    }
    int open(const char* path);
    void close();
    void* ptr;
    size_t size;
};

struct custom_layer_registry_entry
{
    std::string name;
    layer_creator_func creator;
    layer_destroyer_func destroyer;
    void* userdata;
};

class NetPrivate
{
public:
    NetPrivate()
        : device_index(-1), fused_layers(0), planned_concats(0)
    {
    }
    std::vector<Blob> blobs;
    std::vector<Layer*> layers;
    std::vector<int> layer_custom_index; // -1 built-in
    std::vector<int> input_blob_indexes;
    std::vector<int> output_blob_indexes;
    std::vector<const char*> input_blob_names;
    std::vector<const char*> output_blob_names;
    std::vector<custom_layer_registry_entry> custom_layer_registry;
    int device_index;
    int fused_layers;
    MappedModelFile mapped_model;

    // ---- Concat in place (load-time plan): a channel-axis Concat of 3-D blobs whose inputs' channel counts are
    // known from the graph gets ONE buffer per forward walk; the layer that produces an input (through Split shares and Slice
    // views) allocates its top blob as a channel-range VIEW of that buffer, so the Concat itself has nothing left to copy.
    struct ConcatPlan
    {
        ConcatPlan()
            : planned(false), total_c(0)
        {
        }
        bool planned;
        int total_c;
        std::vector<int> offset; // channel offset of every bottom in the buffer
    };
    std::vector<ConcatPlan> concat_plan; // per layer
    std::vector<int> placed_concat;      // per blob: Concat layer whose buffer the blob's producer writes into (-1: none)
    std::vector<int> placed_offset;      // per blob: channel offset there
    std::vector<int> static_channels;    // per blob: channel count when it is a 3-D blob whose channels the graph fixes, else -1
    int planned_concats;

    void plan_concat_placement();
    void update_input_output_indexes();
    void update_input_output_names();
    int fuse_graph(const Option& opt);
    int forward_layer(int layer_index, std::vector<Mat>& blob_mats, std::vector<CudaMat>& blob_mats_gpu, CudaCompute& cmd, const Option& opt) const;
    int do_forward_layer(const Layer* layer, std::vector<CudaMat>& blob_mats_gpu, CudaCompute& cmd, const Option& opt_in) const;
};

int MappedModelFile::open(const char* path)
{
    close();
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0)
    {
        ::close(fd);
        return -1;
    }
    void* p = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd); // the mapping keeps the file
    if (p == MAP_FAILED) return -1;
    ptr = p;
    size = (size_t)st.st_size;
    return 0;
}

void MappedModelFile::close()
{
    if (ptr) munmap(ptr, size);
    ptr = 0;
    size = 0;
}

Net::Net()
    : d(new NetPrivate)
{
}

Net::~Net()
{
    clear();
    delete d;
}

void Net::set_cuda_device(int device_index)
{
    d->device_index = device_index;
    opt.cuda_device_index = device_index;
}

int Net::cuda_device() const
{
    return d->device_index;
}

int Net::register_custom_layer(const char* type, layer_creator_func creator, layer_destroyer_func destroyer, void* userdata)
{
    for (size_t i = 0; i < d->custom_layer_registry.size(); i++)
    {
        if (d->custom_layer_registry[i].name == type)
        {
            NCNN_LOGE("overwrite existing custom layer type %s", type);
            d->custom_layer_registry[i].creator = creator;
            d->custom_layer_registry[i].destroyer = destroyer;
            d->custom_layer_registry[i].userdata = userdata;
            return 0;
        }
    }
    if (layer_to_index(type) != -1) NCNN_LOGE("overwrite built-in layer type %s", type);
    custom_layer_registry_entry e;
    e.name = type;
    e.creator = creator;
    e.destroyer = destroyer;
    e.userdata = userdata;
    d->custom_layer_registry.push_back(e);
    return 0;
}

Layer* Net::create_layer_by_type(const char* type, int* custom_index)
{
    *custom_index = -1;
    // registered types win over built-ins (src/net.cpp:1399-1413)
    for (size_t i = 0; i < d->custom_layer_registry.size(); i++)
    {
        if (d->custom_layer_registry[i].name == type && d->custom_layer_registry[i].creator)
        {
            Layer* layer = d->custom_layer_registry[i].creator(d->custom_layer_registry[i].userdata);
            if (layer) *custom_index = (int)i;
            return layer;
        }
    }
    return create_layer_cuda(type);
}

void NetPrivate::update_input_output_indexes()
{
    // inputs = tops of Input layers, outputs = blobs with a producer and no consumer (src/net.cpp:1145-1169)
    input_blob_indexes.clear();
    output_blob_indexes.clear();
    for (size_t i = 0; i < layers.size(); i++)
    {
        if (layers[i] && layers[i]->type == "Input" && !layers[i]->tops.empty()) input_blob_indexes.push_back(layers[i]->tops[0]);
    }
    for (size_t i = 0; i < blobs.size(); i++)
    {
        if (blobs[i].producer != -1 && blobs[i].consumer == -1) output_blob_indexes.push_back((int)i);
    }
}

void NetPrivate::update_input_output_names()
{
    input_blob_names.clear();
    output_blob_names.clear();
    for (size_t i = 0; i < input_blob_indexes.size(); i++) input_blob_names.push_back(blobs[input_blob_indexes[i]].name.c_str());
    for (size_t i = 0; i < output_blob_indexes.size(); i++) output_blob_names.push_back(blobs[output_blob_indexes[i]].name.c_str());
}

static bool next_line(const std::string& text, size_t& pos, std::string& line)
{
    while (pos < text.size())
    {
        size_t e = text.find('\n', pos);
        if (e == std::string::npos) e = text.size();
        line.assign(text, pos, e - pos);
        pos = e + 1;
        // skip blank lines
        size_t k = 0;
        while (k < line.size() && isspace((unsigned char)line[k])) k++;
        if (k < line.size()) return true;
    }
    return false;
}

static const char* next_token(const char* p, const char* end, std::string& tok)
{
    while (p < end && isspace((unsigned char)*p)) p++;
    const char* b = p;
    while (p < end && !isspace((unsigned char)*p)) p++;
    tok.assign(b, p);
    return p;
}

static Mat shape_hint_mat(int dims, int w, int h, int dd, int c)
{
    Mat m;
    m.dims = dims;
    m.w = w;
    m.h = h;
    m.d = dd;
    m.c = c;
    m.elemsize = 4u;
    m.elempack = 1;
    return m;
}

// src/net.cpp:1305-1665
int Net::load_param_text(const std::string& text)
{
    clear();
    size_t pos = 0;
    std::string line;
    if (!next_line(text, pos, line)) return -1;
    int magic = atoi(line.c_str());
    if (magic != 7767517)
    {
        NCNN_LOGE("param is too old, please regenerate");
        return -1;
    }
    if (!next_line(text, pos, line)) return -1;
    int layer_count = 0, blob_count = 0;
    if (sscanf(line.c_str(), "%d %d", &layer_count, &blob_count) != 2 || layer_count <= 0 || blob_count <= 0)
    {
        NCNN_LOGE("invalid layer_count or blob_count");
        return -1;
    }
    if (!opt.use_cuda_compute)
    {
        NCNN_LOGE("opt.use_cuda_compute is off: this runtime has no CPU compute path");
        return -1;
    }
    d->layers.resize((size_t)layer_count, 0);
    d->layer_custom_index.resize((size_t)layer_count, -1);
    d->blobs.resize((size_t)blob_count);

    ParamDict pd;
    int blob_index = 0;
    for (int i = 0; i < layer_count; i++)
    {
        if (!next_line(text, pos, line))
        {
            NCNN_LOGE("parse layer %d failed: unexpected end of param", i);
            clear();
            return -1;
        }
        const char* p = line.c_str();
        const char* end = p + line.size();
        std::string layer_type, layer_name, tok;
        p = next_token(p, end, layer_type);
        p = next_token(p, end, layer_name);
        p = next_token(p, end, tok);
        int bottom_count = atoi(tok.c_str());
        p = next_token(p, end, tok);
        int top_count = atoi(tok.c_str());
        if (layer_type.empty() || layer_name.empty() || bottom_count < 0 || top_count < 0)
        {
            NCNN_LOGE("parse layer %d failed", i);
            clear();
            return -1;
        }
        int custom_index = -1;
        Layer* layer = create_layer_by_type(layer_type.c_str(), &custom_index);
        if (!layer)
        {
            NCNN_LOGE("layer %s not exists or registered (the CUDA backend has no CPU fallback)", layer_type.c_str());
            clear();
            return -1;
        }
        layer->type = layer_type;
        layer->name = layer_name;
        layer->bottoms.resize(bottom_count);
        for (int j = 0; j < bottom_count; j++)
        {
            p = next_token(p, end, tok);
            int bottom_blob_index = find_blob_index_by_name(tok.c_str());
            if (bottom_blob_index == -1)
            {
                if (blob_index >= blob_count)
                {
                    NCNN_LOGE("blob count exceeds the header's %d", blob_count);
                    delete layer;
                    clear();
                    return -1;
                }
                bottom_blob_index = blob_index;
                d->blobs[blob_index].name = tok;
                blob_index++;
            }
            d->blobs[bottom_blob_index].consumer = i;
            layer->bottoms[j] = bottom_blob_index;
        }
        layer->tops.resize(top_count);
        for (int j = 0; j < top_count; j++)
        {
            p = next_token(p, end, tok);
            if (blob_index >= blob_count || tok.empty())
            {
                NCNN_LOGE("blob count exceeds the header's %d", blob_count);
                delete layer;
                clear();
                return -1;
            }
            d->blobs[blob_index].name = tok;
            d->blobs[blob_index].producer = i;
            layer->tops[j] = blob_index;
            blob_index++;
        }
        if (pd.load_param_text(p, end) != 0)
        {
            NCNN_LOGE("ParamDict load_param %d %s failed", i, layer_name.c_str());
            delete layer;
            clear();
            return -1;
        }
        // top shape hints, id 30 (src/net.cpp:1487-1519)
        Mat shape_hints = pd.get(30, Mat());
        // an int array of top_count records of (dims, extents...): anything else (a float array, too few values) is ignored
        const int hint_type = pd.type(30);
        const int psh_step_checked = (!shape_hints.empty() && top_count > 0) ? shape_hints.w / top_count : 0;
        if ((hint_type == 5 || hint_type == 4) && psh_step_checked >= 2)
        {
            const int psh_step = psh_step_checked;
            const int* psh = (const int*)shape_hints.data;
            for (int j = 0; j < top_count; j++)
            {
                Blob& blob = d->blobs[layer->tops[j]];
                int dims = psh[0];
                if (dims < 1 || dims > 4 || dims + 1 > psh_step)
                {
                    psh += psh_step;
                    continue;
                }
                if (dims == 1) blob.shape = shape_hint_mat(1, psh[1], 1, 1, 1);
                if (dims == 2) blob.shape = shape_hint_mat(2, psh[1], psh[2], 1, 1);
                if (dims == 3) blob.shape = psh_step == 5 ? shape_hint_mat(3, psh[1], psh[2], 1, psh[4]) : shape_hint_mat(3, psh[1], psh[2], 1, psh[3]);
                if (dims == 4) blob.shape = shape_hint_mat(4, psh[1], psh[2], psh[3], psh[4]);
                psh += psh_step;
            }
        }
        layer->featmask = pd.get(31, 0);
        layer->top_count_hint = top_count;
        int lr = layer->load_param(pd);
        if (lr != 0)
        {
            NCNN_LOGE("layer load_param %d %s failed", i, layer_name.c_str());
            delete layer;
            clear();
            return -1;
        }
        layer->bottom_shapes.resize(bottom_count);
        for (int j = 0; j < bottom_count; j++) layer->bottom_shapes[j] = d->blobs[layer->bottoms[j]].shape;
        layer->top_shapes.resize(top_count);
        for (int j = 0; j < top_count; j++) layer->top_shapes[j] = d->blobs[layer->tops[j]].shape;
        d->layers[i] = layer;
        d->layer_custom_index[i] = custom_index;
    }
    d->update_input_output_indexes();
    d->update_input_output_names();
    return 0;
}

// Text params arrive through DataReader::scan, exactly the formats the reference pulls them with (src/net.cpp:1305-1400
// SCAN_VALUE, src/paramdict.cpp:263-480): a memory reader has no length to drain with read(), and a custom reader may
// implement scan() only.  The tokens are re-assembled into the line form load_param_text parses.
static bool scan_param_value(const DataReader& dr, int id, std::string& out)
{
    char tmp[32];
    if (id <= -23300)
    {
        // old style array: -233xx=len,v0,v1,...
        int len = 0;
        if (dr.scan("%d", &len) != 1 || len < 0) return false;
        sprintf(tmp, "%d", len);
        out += tmp;
        for (int j = 0; j < len; j++)
        {
            char v[16];
            if (dr.scan(",%15[^,\n ]", v) != 1) return false;
            out += ',';
            out += v;
        }
        return true;
    }
    char v[16];
    if (dr.scan("%15[^,\n ]", v) != 1) return false;
    out += v;
    const bool is_string = v[0] == '"' || isalpha((unsigned char)v[0]);
    if (is_string)
    {
        // the rest of a string longer than the first 15 characters
        char rest[256];
        size_t n = strlen(v);
        if (v[0] == '"')
        {
            if (n < 2 || v[n - 1] != '"')
            {
                if (dr.scan("%255[^\"\n]\"", rest) == 1) out += rest;
                out += '"';
            }
        }
        else if (dr.scan("%255[^\n ]", rest) == 1)
            out += rest;
        return true;
    }
    // new style array: v0,v1,...
    char comma[4];
    while (dr.scan("%1[,]", comma) == 1)
    {
        if (dr.scan("%15[^,\n ]", v) != 1) return false;
        out += ',';
        out += v;
    }
    return true;
}

int Net::load_param(const DataReader& dr)
{
    int magic = 0, layer_count = 0, blob_count = 0;
    if (dr.scan("%d", &magic) != 1 || dr.scan("%d", &layer_count) != 1 || dr.scan("%d", &blob_count) != 1)
    {
        NCNN_LOGE("parse magic / layer_count / blob_count failed");
        return -1;
    }
    if (layer_count <= 0 || blob_count <= 0 || layer_count > (1 << 24))
    {
        NCNN_LOGE("invalid layer_count or blob_count");
        return -1;
    }
    char tmp[64];
    sprintf(tmp, "%d\n%d %d\n", magic, layer_count, blob_count);
    std::string text(tmp);
    for (int i = 0; i < layer_count; i++)
    {
        char layer_type[256], layer_name[256];
        int bottom_count = 0, top_count = 0;
        if (dr.scan("%255s", layer_type) != 1 || dr.scan("%255s", layer_name) != 1 || dr.scan("%d", &bottom_count) != 1 || dr.scan("%d", &top_count) != 1
                || bottom_count < 0 || top_count < 0)
        {
            NCNN_LOGE("parse layer %d failed", i);
            return -1;
        }
        text += layer_type;
        text += ' ';
        text += layer_name;
        sprintf(tmp, " %d %d", bottom_count, top_count);
        text += tmp;
        for (int j = 0; j < bottom_count + top_count; j++)
        {
            char blob_name[256];
            if (dr.scan("%255s", blob_name) != 1)
            {
                NCNN_LOGE("parse blob name of layer %d failed", i);
                return -1;
            }
            text += ' ';
            text += blob_name;
        }
        int id = 0;
        while (dr.scan("%d=", &id) == 1)
        {
            sprintf(tmp, " %d=", id);
            text += tmp;
            if (!scan_param_value(dr, id, text))
            {
                NCNN_LOGE("ParamDict read value failed (layer %d id %d)", i, id);
                return -1;
            }
        }
        text += '\n';
    }
    return load_param_text(text);
}

int Net::load_param(FILE* fp)
{
    std::string text;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
    return load_param_text(text);
}

int Net::load_param(const char* protopath)
{
    FILE* fp = fopen(protopath, "rb");
    if (!fp)
    {
        NCNN_LOGE("fopen %s failed", protopath);
        return -1;
    }
    int ret = load_param(fp);
    fclose(fp);
    return ret;
}

int Net::load_param_mem(const char* mem)
{
    return load_param_text(std::string(mem));
}

// .param.bin (src/net.cpp:1667-2019): int32 magic, layer_count, blob_count; per layer typeindex, bottom_count,
// top_count, blob indexes, then ParamDict binary records
int Net::load_param_bin(const DataReader& dr)
{
    clear();
#define READ_VALUE(buf)                                   \
    if (dr.read(&buf, sizeof(buf)) != sizeof(buf))        \
    {                                                     \
        NCNN_LOGE("read " #buf " failed");                \
        clear();                                          \
        return -1;                                        \
    }
    int magic = 0;
    READ_VALUE(magic)
    if (magic != 7767517)
    {
        NCNN_LOGE("param is too old, please regenerate");
        return -1;
    }
    int layer_count = 0, blob_count = 0;
    READ_VALUE(layer_count)
    READ_VALUE(blob_count)
    if (layer_count <= 0 || blob_count <= 0)
    {
        NCNN_LOGE("invalid layer_count or blob_count");
        return -1;
    }
    d->layers.resize((size_t)layer_count, 0);
    d->layer_custom_index.resize((size_t)layer_count, -1);
    d->blobs.resize((size_t)blob_count);
    ParamDict pd;
    for (int i = 0; i < layer_count; i++)
    {
        int typeindex = 0, bottom_count = 0, top_count = 0;
        READ_VALUE(typeindex)
        READ_VALUE(bottom_count)
        READ_VALUE(top_count)
        // typeindex with LayerType::CustomBit (1 << 8, src/layer_type.h) addresses the custom-layer registry by position
        // (src/net.cpp:1715-1726: create_custom_layer(typeindex & ~CustomBit))
        const int kCustomBit = 1 << 8;
        const char* type = 0;
        int custom_index = -1;
        Layer* layer = 0;
        std::string custom_type_name;
        if (typeindex & kCustomBit)
        {
            const int ci = typeindex & ~kCustomBit;
            if (ci >= 0 && ci < (int)d->custom_layer_registry.size() && d->custom_layer_registry[ci].creator)
            {
                layer = d->custom_layer_registry[ci].creator(d->custom_layer_registry[ci].userdata);
                if (layer) custom_index = ci;
                custom_type_name = d->custom_layer_registry[ci].name;
                type = custom_type_name.c_str();
            }
        }
        else
        {
            type = layer_index_to_type(typeindex);
            layer = type ? create_layer_by_type(type, &custom_index) : 0;
        }
        if (!layer)
        {
            NCNN_LOGE("layer %d not exists or registered (the CUDA backend has no CPU fallback)", typeindex);
            clear();
            return -1;
        }
        layer->type = type;
        layer->typeindex = typeindex;
        char namebuf[32];
        sprintf(namebuf, "layer%d", i);
        layer->name = namebuf;
        layer->bottoms.resize(bottom_count);
        for (int j = 0; j < bottom_count; j++)
        {
            int bottom_blob_index = 0;
            READ_VALUE(bottom_blob_index)
            if (bottom_blob_index < 0 || bottom_blob_index >= blob_count)
            {
                delete layer;
                clear();
                return -1;
            }
            d->blobs[bottom_blob_index].consumer = i;
            layer->bottoms[j] = bottom_blob_index;
        }
        layer->tops.resize(top_count);
        for (int j = 0; j < top_count; j++)
        {
            int top_blob_index = 0;
            READ_VALUE(top_blob_index)
            if (top_blob_index < 0 || top_blob_index >= blob_count)
            {
                delete layer;
                clear();
                return -1;
            }
            d->blobs[top_blob_index].producer = i;
            char bn[32];
            sprintf(bn, "blob%d", top_blob_index);
            d->blobs[top_blob_index].name = bn;
            layer->tops[j] = top_blob_index;
        }
        if (pd.load_param_bin(dr) != 0)
        {
            NCNN_LOGE("ParamDict load_param_bin %d failed", i);
            delete layer;
            clear();
            return -1;
        }
        layer->featmask = pd.get(31, 0);
        layer->top_count_hint = top_count;
        if (layer->load_param(pd) != 0)
        {
            NCNN_LOGE("layer load_param %d failed", i);
            delete layer;
            clear();
            return -1;
        }
        layer->bottom_shapes.resize(bottom_count);
        layer->top_shapes.resize(top_count);
        d->layers[i] = layer;
        d->layer_custom_index[i] = custom_index;
    }
#undef READ_VALUE
    d->update_input_output_indexes();
    d->update_input_output_names();
    return 0;
}

int Net::load_param_bin(const char* protopath)
{
    FILE* fp = fopen(protopath, "rb");
    if (!fp)
    {
        NCNN_LOGE("fopen %s failed", protopath);
        return -1;
    }
    DataReaderFromStdio dr(fp);
    int ret = load_param_bin(dr);
    fclose(fp);
    return ret;
}

// src/net.cpp:2021-2155
int Net::load_model(const DataReader& dr)
{
    if (d->layers.empty())
    {
        NCNN_LOGE("network graph not ready");
        return -1;
    }
    if (d->device_index < 0) d->device_index = opt.cuda_device_index >= 0 ? opt.cuda_device_index : ncnn_cuda_get_device();
    if (d->device_index < 0 || ncnn_cuda_set_device(d->device_index) != 0)
    {
        NCNN_LOGE("no CUDA device available: %s", ncnn_cuda_last_error());
        return -1;
    }
    opt.cuda_device_index = d->device_index;

    int ret = 0;
    ModelBinFromDataReader mb(dr);
    for (size_t i = 0; i < d->layers.size(); i++)
    {
        Layer* layer = d->layers[i];
        if (!layer)
        {
            NCNN_LOGE("load_model error at layer %d, parameter file has inconsistent content.", (int)i);
            ret = -1;
            break;
        }
        int lret = layer->load_model(mb);
        if (lret != 0)
        {
            NCNN_LOGE("layer load_model %d %s failed", (int)i, layer->name.c_str());
            ret = -1;
            break;
        }
    }
    if (ret != 0) return ret;

    if (opt.use_cuda_graph_fusion)
    {
        d->fuse_graph(opt);
        d->plan_concat_placement();
    }

    for (size_t i = 0; i < d->layers.size(); i++)
    {
        Layer* layer = d->layers[i];
        int cret = layer->create_pipeline(opt);
        if (cret != 0)
        {
            NCNN_LOGE("layer create_pipeline %d %s failed", (int)i, layer->name.c_str());
            ret = -1;
            break;
        }
    }
    ncnn_cuda_device_sync();
    return ret;
}

int Net::load_model(FILE* fp)
{
    DataReaderFromStdio dr(fp);
    return load_model(dr);
}

// opt.use_mapped_model_loading (src/net.cpp:2263-2301): the .bin is mmap'ed read-only and parsed in place -- raw fp32 weight
// records become views of the mapping (ModelBinFromDataReader::load via DataReader::reference), nothing is copied through the
// page cache into a second host buffer before create_pipeline re-packs the weights for the device.  The mapping lives until
// Net::clear().  Falls back to stdio when the file cannot be mapped.
int Net::load_model(const char* modelpath)
{
    if (opt.use_mapped_model_loading)
    {
        if (d->mapped_model.open(modelpath) == 0)
        {
            const unsigned char* mem = (const unsigned char*)d->mapped_model.ptr;
            size_t consumed = 0;
            int ret;
            {
                const unsigned char* cur = mem;
                DataReaderFromMemory dr(cur);
                ret = load_model(dr);
                consumed = cur - mem;
            }
            if (ret != 0) return ret;
            if (consumed != d->mapped_model.size)
            {
                NCNN_LOGE("mapped_file consumed %zu != %zu", consumed, d->mapped_model.size);
                d->mapped_model.close();
                return -1;
            }
            return 0;
        }
        // fallback to regular file loading
    }
    FILE* fp = fopen(modelpath, "rb");
    if (!fp)
    {
        NCNN_LOGE("fopen %s failed", modelpath);
        return -1;
    }
    int ret = load_model(fp);
    fclose(fp);
    return ret;
}

size_t Net::load_param_bin_mem(const unsigned char* _mem)
{
    const unsigned char* mem = _mem;
    DataReaderFromMemory dr(mem);
    if (load_param_bin(dr) != 0) return 0;
    return mem - _mem;
}

size_t Net::load_model(const unsigned char* _mem)
{
    const unsigned char* mem = _mem;
    DataReaderFromMemory dr(mem);
    load_model(dr);
    return mem - _mem;
}

void Net::clear()
{
    Option o = opt;
    for (size_t i = 0; i < d->layers.size(); i++)
    {
        Layer* layer = d->layers[i];
        if (!layer) continue;
        layer->destroy_pipeline(o);
        int ci = i < d->layer_custom_index.size() ? d->layer_custom_index[i] : -1;
        if (ci >= 0 && d->custom_layer_registry[ci].destroyer)
            d->custom_layer_registry[ci].destroyer(layer, d->custom_layer_registry[ci].userdata);
        else
            delete layer;
    }
    d->layers.clear();
    d->layer_custom_index.clear();
    d->blobs.clear();
    d->mapped_model.close(); // (after the layers: their weight Mats may be views of the mapping)
    d->input_blob_indexes.clear();
    d->output_blob_indexes.clear();
    d->input_blob_names.clear();
    d->output_blob_names.clear();
    d->fused_layers = 0;
}

Extractor Net::create_extractor() const
{
    return Extractor(this, d->blobs.size());
}

const std::vector<int>& Net::input_indexes() const
{
    return d->input_blob_indexes;
}
const std::vector<int>& Net::output_indexes() const
{
    return d->output_blob_indexes;
}
const std::vector<const char*>& Net::input_names() const
{
    return d->input_blob_names;
}
const std::vector<const char*>& Net::output_names() const
{
    return d->output_blob_names;
}
const std::vector<Blob>& Net::blobs() const
{
    return d->blobs;
}
const std::vector<Layer*>& Net::layers() const
{
    return d->layers;
}
int Net::fused_layer_count() const
{
    return d->fused_layers;
}

int Net::find_blob_index_by_name(const char* name) const
{
    for (size_t i = 0; i < d->blobs.size(); i++)
        if (d->blobs[i].name == name) return (int)i;
    return -1;
}

int Net::find_layer_index_by_name(const char* name) const
{
    for (size_t i = 0; i < d->layers.size(); i++)
        if (d->layers[i] && d->layers[i]->name == name) return (int)i;
    return -1;
}

// ------------------------------------------------------------------ load-time graph fusion
// What tools/ncnnoptimize.cpp does offline (fuse_convolution_activation :1268-1419 and friends), done here on the
// loaded graph so that un-optimised .param files get the same kernels.  A folded layer is taken out of the walk by
// re-wiring: the producer takes over the folded layer's top blob and the folded layer is replaced by a disconnected Noop.
//   Convolution / ConvolutionDepthWise / InnerProduct (activation_type 0) -> ReLU | Clip | Sigmoid | Mish | HardSwish   => epilogue activation
//   Convolution (activation_type 0) -> Swish                                       => epilogue activation (code 7)
//   Convolution (activation_type 0) -> Eltwise(SUM, 2 inputs, no coeffs) [-> ReLU]  => residual add (+ReLU) in the epilogue
//   Eltwise -> ReLU                                                               => fused_relu
namespace {
class RetiredLayer : public Layer
{
public:
    RetiredLayer()
    {
        one_blob_only = false;
        support_inplace = false;
    }
    virtual int forward(const std::vector<CudaMat>&, std::vector<CudaMat>&, CudaCompute&, const Option&) const
    {
        return 0;
    }
};
} // namespace

int NetPrivate::fuse_graph(const Option&)
{
    const int L = (int)layers.size();
    // a blob must not be a net output to be folded away
    auto sole_consumer = [&](int blob) -> int { return blobs[blob].consumer; };
    auto retire = [&](int li, bool keep_object = false) {
        // turn layer li into a disconnected no-op (keep_object: the layer object lives on inside the layer it was folded into)
        Layer* old = layers[li];
        RetiredLayer* n = new RetiredLayer;
        n->type = "Noop";
        n->name = old->name;
        int ci = layer_custom_index[li];
        if (keep_object)
            ;
        else if (ci >= 0 && custom_layer_registry[ci].destroyer)
            custom_layer_registry[ci].destroyer(old, custom_layer_registry[ci].userdata);
        else
            delete old;
        layers[li] = n;
        layer_custom_index[li] = -1;
        fused_layers++;
    };
    for (int i = 0; i < L; i++)
    {
        if (layer_custom_index[i] >= 0) continue;
        Layer* l = layers[i];
        // ---- X -> activation
        int* act_slot = 0;
        Mat* act_params = 0;
        bool is_conv = false;
        if (l->type == "Convolution")
        {
            Convolution* c = (Convolution*)l;
            act_slot = &c->activation_type;
            act_params = &c->activation_params;
            is_conv = true;
        }
        else if (l->type == "ConvolutionDepthWise")
        {
            ConvolutionDepthWise* c = (ConvolutionDepthWise*)l;
            act_slot = &c->activation_type;
            act_params = &c->activation_params;
        }
        else if (l->type == "InnerProduct")
        {
            InnerProduct* c = (InnerProduct*)l;
            act_slot = &c->activation_type;
            act_params = &c->activation_params;
        }
        if (act_slot && *act_slot == 0 && l->tops.size() == 1)
        {
            int top = l->tops[0];
            int j = sole_consumer(top);
            if (j > i && layer_custom_index[j] < 0 && layers[j]->bottoms.size() == 1 && layers[j]->tops.size() == 1)
            {
                Layer* a = layers[j];
                int code = -1;
                float p0 = 0.f, p1 = 0.f;
                int np = 0;
                if (a->type == "ReLU")
                {
                    float slope = ((ReLU*)a)->p0;
                    if (slope == 0.f)
                        code = 1;
                    else
                    {
                        code = 2;
                        p0 = slope;
                        np = 1;
                    }
                }
                else if (a->type == "Clip")
                {
                    code = 3;
                    p0 = ((Clip*)a)->p0;
                    p1 = ((Clip*)a)->p1;
                    np = 2;
                }
                else if (a->type == "Sigmoid")
                    code = 4;
                else if (a->type == "Mish")
                    code = 5;
                else if (a->type == "HardSwish")
                {
                    code = 6;
                    p0 = ((HardSwish*)a)->p0;
                    p1 = ((HardSwish*)a)->p1;
                    np = 2;
                }
                else if (a->type == "Swish" && is_conv)
                    code = 7; // backend-private epilogue code (common.cuh apply_activation)
                if (code > 0)
                {
                    *act_slot = code;
                    if (np > 0)
                    {
                        Mat ap(np);
                        ((float*)ap.data)[0] = p0;
                        if (np > 1) ((float*)ap.data)[1] = p1;
                        *act_params = ap;
                    }
                    // the producer now writes the activation's top blob
                    int newtop = a->tops[0];
                    l->tops[0] = newtop;
                    blobs[newtop].producer = i;
                    blobs[top].producer = -1;
                    blobs[top].consumer = -1;
                    blobs[top].folded_into = i;
                    a->bottoms.clear();
                    a->tops.clear();
                    retire(j);
                    continue;
                }
            }
        }
    }

    // ---- stem Convolution (<= 4 input channels, stride 2, none / ReLU) -> max Pooling 3x3 s2: one kernel, the full-resolution map
    // never reaches HBM (include/ncnn_cuda.h ncnn_cuda_conv2d_forward_maxpool3x3s2; other geometries run the two layers in turn)
    for (int i = 0; i < L; i++)
    {
        if (layer_custom_index[i] >= 0 || layers[i]->type != "Convolution") continue;
        Convolution* c = (Convolution*)layers[i];
        if (c->tops.size() != 1 || c->bottoms.size() != 1 || c->fused_residual || c->shortcut || c->fused_pool) continue;
        if (c->activation_type != 0 && c->activation_type != 1) continue;
        const int inch = c->weight_data_size / (c->num_output * c->kernel_w * c->kernel_h);
        if (inch > 4 || c->stride_w != 2 || c->stride_h != 2 || c->num_output > 64 || c->kernel_w * c->kernel_h < 2) continue;
        int top = c->tops[0];
        int j = sole_consumer(top);
        if (j <= i || layer_custom_index[j] >= 0 || layers[j]->type != "Pooling") continue;
        Pooling* pl = (Pooling*)layers[j];
        if (pl->pooling_type != 0 || pl->global_pooling || pl->adaptive_pooling || pl->kernel_w != 3 || pl->kernel_h != 3 || pl->stride_w != 2 || pl->stride_h != 2)
            continue;
        if (pl->bottoms.size() != 1 || pl->tops.size() != 1) continue;
        int newtop = pl->tops[0];
        c->fused_pool = pl;
        c->tops[0] = newtop;
        blobs[newtop].producer = i;
        blobs[top].producer = -1;
        blobs[top].consumer = -1;
        blobs[top].folded_into = i;
        pl->bottoms.clear();
        pl->tops.clear();
        retire(j, true);
    }

    // ---- Convolution -> Eltwise(SUM) [-> ReLU]
    for (int j = 0; j < L; j++)
    {
        if (layer_custom_index[j] >= 0) continue;
        Layer* e = layers[j];
        if (e->type != "Eltwise") continue;
        Eltwise* el = (Eltwise*)e;
        // Eltwise -> ReLU first
        int etop = e->tops.size() == 1 ? e->tops[0] : -1;
        int relu_layer = -1;
        if (etop >= 0)
        {
            int k = blobs[etop].consumer;
            if (k > j && layer_custom_index[k] < 0 && layers[k]->type == "ReLU" && ((ReLU*)layers[k])->p0 == 0.f && layers[k]->tops.size() == 1) relu_layer = k;
        }
        bool folded_into_conv = false;
        if (el->op_type == 1 && el->coeffs.empty() && e->bottoms.size() == 2 && etop >= 0)
        {
            // pick the operand produced LAST by a plain Convolution whose only consumer is this Eltwise
            int best = -1, best_slot = -1;
            for (int s = 0; s < 2; s++)
            {
                int b = e->bottoms[s];
                int p = blobs[b].producer;
                if (p < 0 || p >= j || layer_custom_index[p] >= 0) continue;
                if (layers[p]->type != "Convolution") continue;
                Convolution* c = (Convolution*)layers[p];
                if (c->activation_type != 0 || c->fused_residual || c->tops.size() != 1 || blobs[b].consumer != j) continue;
                if (p > best)
                {
                    best = p;
                    best_slot = s;
                }
            }
            if (best >= 0)
            {
                int other_blob = e->bottoms[1 - best_slot];
                int other_prod = blobs[other_blob].producer;
                // the residual must exist before the conv runs: its producer has to come earlier in layer order
                if (other_prod < best)
                {
                    Convolution* c = (Convolution*)layers[best];
                    int conv_top = c->tops[0];
                    // the other operand is itself a bare 1x1 projection (ResNet's downsample branch) and this layer is a plain 1x1:
                    // fold it in as extra K of ONE GEMM -- its output blob is never written or re-read
                    Convolution* sc = 0;
                    if (other_prod >= 0 && layer_custom_index[other_prod] < 0 && layers[other_prod]->type == "Convolution")
                    {
                        Convolution* k = (Convolution*)layers[other_prod];
                        auto bare1x1 = [](const Convolution* q) {
                            return q->kernel_w == 1 && q->kernel_h == 1 && q->dilation_w == 1 && q->dilation_h == 1 && q->pad_left == 0 && q->pad_right == 0 &&
                                   q->pad_top == 0 && q->pad_bottom == 0 && q->activation_type == 0 && !q->fused_residual && !q->shortcut;
                        };
                        if (bare1x1(k) && bare1x1(c) && c->stride_w == 1 && c->stride_h == 1 && k->num_output == c->num_output && k->tops.size() == 1 &&
                                k->bottoms.size() == 1 && c->bottoms.size() == 1 && blobs[other_blob].consumer == j)
                            sc = k;
                    }
                    c->one_blob_only = false;
                    if (sc)
                    {
                        int sc_bottom = sc->bottoms[0];
                        c->shortcut = sc;
                        c->bottoms.push_back(sc_bottom);
                        blobs[sc_bottom].consumer = best;
                        blobs[other_blob].producer = -1;
                        blobs[other_blob].consumer = -1;
                        blobs[other_blob].folded_into = best;
                        sc->bottoms.clear();
                        sc->tops.clear();
                        retire(other_prod, true);
                    }
                    else
                    {
                        c->fused_residual = true;
                        c->bottoms.push_back(other_blob);
                        blobs[other_blob].consumer = best;
                    }
                    int final_top = etop;
                    c->fused_post_activation = -1;
                    if (relu_layer >= 0)
                    {
                        c->fused_post_activation = 1;
                        final_top = layers[relu_layer]->tops[0];
                        layers[relu_layer]->bottoms.clear();
                        layers[relu_layer]->tops.clear();
                        blobs[etop].producer = -1;
                        blobs[etop].consumer = -1;
                        blobs[etop].folded_into = best;
                        retire(relu_layer);
                    }
                    c->tops[0] = final_top;
                    blobs[final_top].producer = best;
                    blobs[conv_top].producer = -1;
                    blobs[conv_top].consumer = -1;
                    blobs[conv_top].folded_into = best;
                    e->bottoms.clear();
                    e->tops.clear();
                    retire(j);
                    folded_into_conv = true;
                }
            }
        }
        if (!folded_into_conv && relu_layer >= 0)
        {
            el->fused_relu = true;
            int newtop = layers[relu_layer]->tops[0];
            e->tops[0] = newtop;
            blobs[newtop].producer = j;
            blobs[etop].producer = -1;
            blobs[etop].consumer = -1;
            blobs[etop].folded_into = j;
            layers[relu_layer]->bottoms.clear();
            layers[relu_layer]->tops.clear();
            retire(relu_layer);
        }
    }
    update_input_output_indexes();
    update_input_output_names();
    return 0;
}

// ------------------------------------------------------------------ Concat in place: load-time plan
static bool is_channel_axis_3d(int axis)
{
    return axis == 0 || axis == -3;
}

// per-top channel counts of a channel-axis Slice of a 3-D blob with `extent` channels (slice.cpp:40-72); false when not computable
static bool slice_channel_counts(const Slice* sl, int extent, size_t ntops, std::vector<int>& out)
{
    out.assign(ntops, 0);
    const int* slices_ptr = (const int*)sl->slices.data;
    const int* indices_ptr = (const int*)sl->indices.data;
    int q = 0;
    for (size_t i = 0; i < ntops; i++)
    {
        int slice;
        if (indices_ptr)
        {
            if (i == ntops - 1)
                slice = extent - q;
            else
            {
                if ((int)i >= sl->indices.w) return false;
                int indice = indices_ptr[i];
                int positive_indice = indice < 0 ? extent + indice : indice;
                slice = positive_indice - q;
            }
        }
        else
        {
            if (!slices_ptr || (int)i >= sl->slices.w) return false;
            slice = slices_ptr[i];
            if (slice == -233) slice = (int)((extent - q) / (ntops - i));
        }
        if (slice <= 0 || q + slice > extent) return false;
        out[i] = slice;
        q += slice;
    }
    return true;
}

void NetPrivate::plan_concat_placement()
{
    const size_t L = layers.size();
    concat_plan.assign(L, ConcatPlan());
    placed_concat.assign(blobs.size(), -1);
    placed_offset.assign(blobs.size(), 0);
    static_channels.assign(blobs.size(), -1);
    planned_concats = 0;
    std::vector<int>& chan = static_channels;

    // 1. channel counts the graph fixes (3-D blobs only: everything here starts from a 3-D Input hint or a convolution)
    static const char* same_as_bottom[] = {"ReLU", "Sigmoid", "Swish", "TanH", "Mish", "Clip", "HardSwish", "HardSigmoid", "GELU", "Dropout", "BatchNorm", "Scale",
                                           "LRN", "Noop", "Pooling", "Interp", "Eltwise", "ShuffleChannel", 0};
    for (size_t li = 0; li < L; li++)
    {
        const Layer* layer = layers[li];
        if (!layer || layer->tops.empty() || layer_custom_index[li] >= 0) continue;
        const std::string& t = layer->type;
        const int b0 = layer->bottoms.empty() ? -1 : chan[layer->bottoms[0]];
        if (t == "Input")
        {
            const Input* in = (const Input*)layer;
            if (in->w > 0 && in->h > 0 && in->c > 0 && in->d == 0) chan[layer->tops[0]] = in->c;
            continue;
        }
        if (t == "Convolution" || t == "ConvolutionDepthWise" || t == "Deconvolution" || t == "DeconvolutionDepthWise")
        {
            if (b0 < 0) continue;
            int num_output = -1;
            if (t == "Convolution") num_output = ((const Convolution*)layer)->num_output;
            if (t == "ConvolutionDepthWise") num_output = ((const ConvolutionDepthWise*)layer)->num_output;
            if (t == "Deconvolution" || t == "DeconvolutionDepthWise") num_output = ((const Deconvolution*)layer)->num_output;
            chan[layer->tops[0]] = num_output;
            continue;
        }
        bool same = false;
        for (int k = 0; same_as_bottom[k]; k++) same = same || t == same_as_bottom[k];
        if (same)
        {
            if (t == "Eltwise")
                for (size_t j = 1; j < layer->bottoms.size(); j++)
                    if (chan[layer->bottoms[j]] != b0) same = false;
            if (t == "Pooling" && ((const Pooling*)layer)->global_pooling) same = false; // (stays 3-D with the same channels, but never feeds a concat buffer of its size)
            if (same && layer->tops.size() == 1) chan[layer->tops[0]] = b0;
            continue;
        }
        if (t == "BinaryOp")
        {
            const BinaryOp* bo = (const BinaryOp*)layer;
            if (bo->with_scalar || layer->bottoms.size() == 1)
                chan[layer->tops[0]] = b0;
            else if (layer->bottoms.size() == 2 && b0 >= 0 && chan[layer->bottoms[1]] == b0)
                chan[layer->tops[0]] = b0;
            continue;
        }
        if (t == "Split")
        {
            for (size_t j = 0; j < layer->tops.size(); j++) chan[layer->tops[j]] = b0;
            continue;
        }
        if (t == "Slice")
        {
            const Slice* sl = (const Slice*)layer;
            std::vector<int> counts;
            if (b0 >= 0 && is_channel_axis_3d(sl->axis) && slice_channel_counts(sl, b0, layer->tops.size(), counts))
                for (size_t j = 0; j < layer->tops.size(); j++) chan[layer->tops[j]] = counts[j];
            continue;
        }
        if (t == "Concat")
        {
            const Concat* cc = (const Concat*)layer;
            int sum = 0;
            bool ok = is_channel_axis_3d(cc->axis);
            for (size_t j = 0; j < layer->bottoms.size() && ok; j++)
            {
                if (chan[layer->bottoms[j]] < 0) ok = false;
                sum += chan[layer->bottoms[j]];
            }
            if (ok) chan[layer->tops[0]] = sum;
            continue;
        }
    }

    // 2. a blob seen through Split shares and channel-axis Slice views: (root blob, first channel inside it)
    auto resolve = [&](int b, int& root, int& start) {
        root = b;
        start = 0;
        for (;;)
        {
            const int p = blobs[root].producer;
            if (p < 0 || layer_custom_index[p] >= 0) return;
            const Layer* pl = layers[p];
            if (pl->type == "Split")
            {
                root = pl->bottoms[0];
                continue;
            }
            if (pl->type == "Slice" && is_channel_axis_3d(((const Slice*)pl)->axis) && chan[pl->bottoms[0]] >= 0)
            {
                int off = 0;
                bool found = false;
                for (size_t j = 0; j < pl->tops.size(); j++)
                {
                    if (pl->tops[j] == root)
                    {
                        found = true;
                        break;
                    }
                    if (chan[pl->tops[j]] < 0) return;
                    off += chan[pl->tops[j]];
                }
                if (!found) return;
                start += off;
                root = pl->bottoms[0];
                continue;
            }
            return;
        }
    };
    static const char* placeable[] = {"Convolution", "ConvolutionDepthWise", "Deconvolution", "DeconvolutionDepthWise", "Pooling", "Interp", "Eltwise", "BinaryOp",
                                      "ReLU", "Sigmoid", "Swish", "TanH", "Mish", "Clip", "HardSwish", "HardSigmoid", "GELU", 0};
    const int kVec = 8; // channel offsets / counts in whole 16-byte vectors for 16-bit blobs (and 32 bytes for fp32 ones)
    for (size_t k = 0; k < L; k++)
    {
        const Layer* layer = layers[k];
        if (!layer || layer->tops.size() != 1 || layer_custom_index[k] >= 0 || layer->type != "Concat") continue;
        if (!is_channel_axis_3d(((const Concat*)layer)->axis) || layer->bottoms.size() < 2) continue;
        ConcatPlan plan;
        bool ok = true;
        int off = 0;
        for (size_t j = 0; j < layer->bottoms.size() && ok; j++)
        {
            const int c = chan[layer->bottoms[j]];
            if (c <= 0 || c % kVec != 0) ok = false;
            plan.offset.push_back(off);
            off += c;
        }
        if (!ok) continue;
        plan.total_c = off;
        // groups of consecutive bottoms that are adjacent channel ranges of one root covering it completely
        int marked = 0;
        for (size_t i = 0; i < layer->bottoms.size();)
        {
            int root, start;
            resolve(layer->bottoms[i], root, start);
            int covered = chan[layer->bottoms[i]];
            size_t j = i + 1;
            if (start == 0)
            {
                for (; j < layer->bottoms.size(); j++)
                {
                    int r2, s2;
                    resolve(layer->bottoms[j], r2, s2);
                    if (r2 != root || s2 != covered) break;
                    covered += chan[layer->bottoms[j]];
                }
            }
            const int prod = blobs[root].producer;
            bool can = start == 0 && covered == chan[root] && prod >= 0 && layer_custom_index[prod] < 0 && placed_concat[root] < 0;
            if (can)
            {
                bool allowed = false;
                for (int q = 0; placeable[q]; q++) allowed = allowed || layers[prod]->type == placeable[q];
                if (layers[prod]->type == "BinaryOp" && ((const BinaryOp*)layers[prod])->with_scalar) allowed = false;
                if (layers[prod]->tops.size() != 1) allowed = false;
                can = allowed;
            }
            if (can)
            {
                placed_concat[root] = (int)k;
                placed_offset[root] = plan.offset[i];
                marked++;
            }
            i = can ? j : i + 1;
        }
        if (marked > 0)
        {
            plan.planned = true;
            concat_plan[k] = plan;
            planned_concats++;
        }
    }
}

// Hands the blob a producer creates for its top a channel-range view of its Concat's buffer (allocating that buffer at the first
// request of a walk: its pixel grid and batch are only known then).  Lives on the executor's stack for one layer call.
class CudaPlacementAllocator : public CudaAllocator
{
public:
    CudaPlacementAllocator(CudaAllocator* real_allocator, CudaMat* _slot, int _offset, int _total_c, int _expect_c)
        : CudaAllocator(real_allocator->device_index), real_(real_allocator), slot(_slot), offset(_offset), total_c(_total_c), expect_c(_expect_c), used(false)
    {
    }
    virtual void* fastMalloc(size_t size)
    {
        return real_->fastMalloc(size);
    }
    virtual void fastFree(void* ptr)
    {
        real_->fastFree(ptr);
    }
    virtual CudaAllocator* real()
    {
        return real_;
    }
    virtual bool place(CudaMat& m)
    {
        if (used || m.dims != 3 || m.c != expect_c) return false;
        const int vec = m.elemtype == NCNN_CUDA_F32 ? 4 : 8;
        if (expect_c % vec != 0 || offset % vec != 0 || total_c % vec != 0) return false;
        if (slot->empty())
        {
            slot->create(m.w, m.h, total_c, m.elemtype, m.n, real_);
            if (slot->empty()) return false;
        }
        else if (slot->w != m.w || slot->h != m.h || slot->n != m.n || slot->elemtype != m.elemtype || slot->c != total_c)
            return false; // (an earlier input of this walk had another pixel grid: this one is copied by the Concat instead)
        CudaMat v = slot->channel_range(offset, expect_c);
        if (v.empty()) return false;
        m = v;
        used = true;
        return true;
    }

private:
    CudaAllocator* real_;
    CudaMat* slot;
    int offset, total_c, expect_c;
    bool used;
};

// ------------------------------------------------------------------ executor
// src/net.cpp:192-356 (Vulkan forward_layer): depth-first, lazy
int NetPrivate::forward_layer(int layer_index, std::vector<Mat>& blob_mats, std::vector<CudaMat>& blob_mats_gpu, CudaCompute& cmd, const Option& opt) const
{
    const Layer* layer = layers[layer_index];
    for (size_t i = 0; i < layer->bottoms.size(); i++)
    {
        int bottom_blob_index = layer->bottoms[i];
        if (!blob_mats_gpu[bottom_blob_index].empty()) continue;
        if (!blob_mats[bottom_blob_index].empty())
        {
            // host -> device boundary: upload once (src/net.cpp:215-228)
            int ret = cmd.record_upload(blob_mats[bottom_blob_index], blob_mats_gpu[bottom_blob_index], opt);
            if (ret != 0) return ret;
            if (opt.lightmode) blob_mats[bottom_blob_index].release();
            continue;
        }
        int producer = blobs[bottom_blob_index].producer;
        if (producer < 0)
        {
            NCNN_LOGE("blob %s has no producer and was not given as input", blobs[bottom_blob_index].name.c_str());
            return -1;
        }
        int ret = forward_layer(producer, blob_mats, blob_mats_gpu, cmd, opt);
        if (ret != 0) return ret;
    }
    if (cmd.profiling()) cmd.profile_begin(layer_index);
    int ret = do_forward_layer(layer, blob_mats_gpu, cmd, opt);
    if (cmd.profiling()) cmd.profile_end(layer->tops.empty() ? CudaMat() : blob_mats_gpu[layer->tops[0]]);
    if (ret != 0) NCNN_LOGE("layer %s (%s) forward failed: %d %s", layer->name.c_str(), layer->type.c_str(), ret, ncnn_cuda_last_error());
    return ret;
}

// src/net.cpp:886-1140 (Vulkan do_forward_layer): lightmode recycling, in-place clone rule; no per-sample batch loop
int NetPrivate::do_forward_layer(const Layer* layer, std::vector<CudaMat>& blob_mats_gpu, CudaCompute& cmd, const Option& opt_in) const
{
    // ---- Concat in place (plan_concat_placement): the walk keeps one buffer per planned Concat in the slots behind the blobs
    // (blob_mats_gpu[blobs.size() + layer index]); a layer whose top is a planned input creates that top through a placement
    // allocator, and the Concat itself only copies the inputs that did not end up in place
    const bool have_slots = planned_concats > 0 && blob_mats_gpu.size() >= blobs.size() + layers.size();
    Option opt = opt_in;
    CudaPlacementAllocator* placer = 0;
    char placer_storage[sizeof(CudaPlacementAllocator)];
    if (have_slots && layer->tops.size() == 1 && placed_concat[layer->tops[0]] >= 0)
    {
        const int t = layer->tops[0];
        const int k = placed_concat[t];
        placer = new (placer_storage) CudaPlacementAllocator(cmd.blob_allocator(opt_in), &blob_mats_gpu[blobs.size() + k], placed_offset[t], concat_plan[k].total_c,
                                                             static_channels[t]);
        opt.blob_cuda_allocator = placer;
    }
    struct PlacerGuard
    {
        CudaPlacementAllocator* p;
        ~PlacerGuard()
        {
            if (p) p->~CudaPlacementAllocator();
        }
    } placer_guard = {placer};

    if (have_slots && layer->type == "Concat" && layer->tops.size() == 1)
    {
        int k = -1;
        for (size_t i = 0; i < layers.size(); i++)
            if (layers[i] == layer)
            {
                k = (int)i;
                break;
            }
        CudaMat& slot = k >= 0 ? blob_mats_gpu[blobs.size() + k] : blob_mats_gpu[0];
        if (k >= 0 && concat_plan[k].planned && !slot.empty())
        {
            const ConcatPlan& plan = concat_plan[k];
            bool usable = true;
            int batch = slot.n;
            for (size_t i = 0; i < layer->bottoms.size() && usable; i++)
            {
                const CudaMat& b = blob_mats_gpu[layer->bottoms[i]];
                usable = b.dims == 3 && b.w == slot.w && b.h == slot.h && b.elemtype == slot.elemtype && b.c == static_channels[layer->bottoms[i]] && (b.n == batch || b.n <= 1);
            }
            if (usable)
            {
                ncnn_cuda_tensor t = slot.view();
                for (size_t i = 0; i < layer->bottoms.size(); i++)
                {
                    const CudaMat& b = blob_mats_gpu[layer->bottoms[i]];
                    const void* expect = (const unsigned char*)slot.data + (size_t)plan.offset[i] * slot.elemsize();
                    if (b.data == expect && b.base == slot.base && b.cpitch == slot.cpitch && b.nstep == slot.nstep) continue; // already in place
                    ncnn_cuda_tensor bv = b.view();
                    int ret = ncnn_cuda_copy_into_axis(&bv, &t, 0, plan.offset[i], cmd.stream());
                    if (ret != 0) return ret;
                }
                blob_mats_gpu[layer->tops[0]] = slot;
                slot.release();
                if (opt.lightmode)
                    for (size_t i = 0; i < layer->bottoms.size(); i++) blob_mats_gpu[layer->bottoms[i]].release();
                return 0;
            }
            slot.release(); // inputs of another shape than planned: the ordinary copying Concat below
        }
    }

    if (layer->one_blob_only)
    {
        int bottom_blob_index = layer->bottoms[0];
        int top_blob_index = layer->tops[0];
        CudaMat& bottom_blob_ref = blob_mats_gpu[bottom_blob_index];
        CudaMat bottom_blob;
        if (opt.lightmode && layer->support_inplace)
        {
            // deep copy for inplace forward if data is shared (src/net.cpp:635-644)
            if (bottom_blob_ref.refcount && *bottom_blob_ref.refcount != 1)
            {
                int ret = cmd.record_clone(bottom_blob_ref, bottom_blob, opt);
                if (ret != 0) return ret;
            }
        }
        if (bottom_blob.dims == 0) bottom_blob = bottom_blob_ref;
        int ret;
        if (opt.lightmode && layer->support_inplace)
        {
            CudaMat& bottom_top_blob = bottom_blob;
            ret = layer->forward_inplace(bottom_top_blob, cmd, opt);
            if (ret != 0) return ret;
            blob_mats_gpu[top_blob_index] = bottom_top_blob;
        }
        else
        {
            CudaMat top_blob;
            ret = layer->forward(bottom_blob, top_blob, cmd, opt);
            if (ret != 0) return ret;
            blob_mats_gpu[top_blob_index] = top_blob;
        }
        if (opt.lightmode) blob_mats_gpu[bottom_blob_index].release();
        return 0;
    }

    std::vector<CudaMat> bottom_blobs(layer->bottoms.size());
    for (size_t i = 0; i < layer->bottoms.size(); i++)
    {
        int bottom_blob_index = layer->bottoms[i];
        CudaMat& ref = blob_mats_gpu[bottom_blob_index];
        if (opt.lightmode && layer->support_inplace && ref.refcount && *ref.refcount != 1)
        {
            int ret = cmd.record_clone(ref, bottom_blobs[i], opt);
            if (ret != 0) return ret;
        }
        if (bottom_blobs[i].dims == 0) bottom_blobs[i] = ref;
    }
    int ret;
    if (opt.lightmode && layer->support_inplace)
    {
        ret = layer->forward_inplace(bottom_blobs, cmd, opt);
        if (ret != 0) return ret;
        for (size_t i = 0; i < layer->tops.size(); i++) blob_mats_gpu[layer->tops[i]] = bottom_blobs[i];
    }
    else
    {
        std::vector<CudaMat> top_blobs(layer->tops.size());
        ret = layer->forward(bottom_blobs, top_blobs, cmd, opt);
        if (ret != 0) return ret;
        for (size_t i = 0; i < layer->tops.size(); i++) blob_mats_gpu[layer->tops[i]] = top_blobs[i];
    }
    if (opt.lightmode)
    {
        for (size_t i = 0; i < layer->bottoms.size(); i++) blob_mats_gpu[layer->bottoms[i]].release();
    }
    return 0;
}

// ------------------------------------------------------------------ Extractor
class ExtractorPrivate
{
public:
    explicit ExtractorPrivate(const Net* _net)
        : net(_net), h2d(0), d2h(0), ctx(0)
    {
    }
    // One stream + device pool for the Extractor's whole life (acquired at the first host extract, handed back by clear() / the
    // destructor): the device blobs an Extractor keeps between extract() calls belong to this pool and were enqueued on this
    // stream, so a later extract() must not run on another stream while those blocks are recycled.
    CudaContext* context()
    {
        if (!ctx) ctx = acquire_cuda_context(net->opt.cuda_device_index);
        return ctx;
    }
    void release_context()
    {
        if (ctx) reclaim_cuda_context(ctx);
        ctx = 0;
    }
    const Net* net;
    std::vector<Mat> blob_mats;
    std::vector<CudaMat> blob_mats_gpu;
    Option opt;
    size_t h2d, d2h;
    CudaContext* ctx;
    // inputs given as raw 8-bit pixels, converted on the device when the walk starts
    struct PixelInput
    {
        int blob_index;
        const unsigned char* pixels;
        int type, w, h, stride, n;
        int target_w, target_h; // 0: no resize
        size_t nstride;
        bool has_mean, has_norm;
        float mean_vals[4], norm_vals[4];
    };
    std::vector<PixelInput> pixel_inputs;
};

Extractor::Extractor(const Net* _net, size_t blob_count)
    : d(new ExtractorPrivate(_net))
{
    d->blob_mats.resize(blob_count);
    // device side: the blobs, then one slot per layer for the buffers of planned in-place Concats (NetPrivate::do_forward_layer)
    d->blob_mats_gpu.resize(blob_count + _net->layers().size());
    d->opt = _net->opt;
}

Extractor::~Extractor()
{
    clear();
    delete d;
}

Extractor::Extractor(const Extractor& rhs)
    : d(new ExtractorPrivate(rhs.d->net))
{
    d->blob_mats = rhs.d->blob_mats;
    d->blob_mats_gpu = rhs.d->blob_mats_gpu;
    d->opt = rhs.d->opt;
}

Extractor& Extractor::operator=(const Extractor& rhs)
{
    if (this == &rhs) return *this;
    d->net = rhs.d->net;
    d->blob_mats = rhs.d->blob_mats;
    d->blob_mats_gpu = rhs.d->blob_mats_gpu;
    d->opt = rhs.d->opt;
    return *this;
}

void Extractor::clear()
{
    d->blob_mats.clear();
    d->blob_mats_gpu.clear(); // (device blocks go back to the pool of the pinned context before the context itself does)
    d->release_context();
}

void Extractor::set_light_mode(bool enable)
{
    d->opt.lightmode = enable;
}
void Extractor::set_blob_allocator(Allocator* allocator)
{
    d->opt.blob_allocator = allocator;
}
void Extractor::set_workspace_allocator(Allocator* allocator)
{
    d->opt.workspace_allocator = allocator;
}
void Extractor::set_blob_cuda_allocator(CudaAllocator* allocator)
{
    d->opt.blob_cuda_allocator = allocator;
}

int Extractor::input(const char* blob_name, const Mat& in)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1)
    {
        NCNN_LOGE("Try");
        const std::vector<const char*>& names = d->net->input_names();
        for (size_t i = 0; i < names.size(); i++) NCNN_LOGE("    ex.input(\"%s\", in%d);", names[i], (int)i);
        return -1;
    }
    return input(blob_index, in);
}

// a blob that load-time fusion folded into a layer does not exist at run time (was silently ignored / "no producer")
static int reject_folded_blob(const Net* net, int blob_index, const char* what)
{
    const Blob& b = net->blobs()[blob_index];
    if (b.folded_into < 0) return 0;
    NCNN_LOGE("%s: blob %s was folded into layer %s by load-time graph fusion and is never materialised; load the net with "
              "opt.use_cuda_graph_fusion = false (ncnn_option_set_use_cuda_graph_fusion(opt, 0)) to %s it",
              what, b.name.c_str(), net->layers()[b.folded_into]->name.c_str(), what);
    return -1;
}

int Extractor::input(int blob_index, const Mat& in)
{
    if (blob_index < 0 || blob_index >= (int)d->blob_mats.size()) return -1;
    if (reject_folded_blob(d->net, blob_index, "input")) return -1;
    d->blob_mats[blob_index] = in;
    d->blob_mats_gpu[blob_index].release();
    return 0;
}

int Extractor::input(const char* blob_name, const CudaMat& in)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1) return -1;
    return input(blob_index, in);
}

int Extractor::input(int blob_index, const CudaMat& in)
{
    if (blob_index < 0 || blob_index >= (int)d->blob_mats.size()) return -1;
    if (reject_folded_blob(d->net, blob_index, "input")) return -1;
    d->blob_mats_gpu[blob_index] = in;
    d->blob_mats[blob_index].release();
    return 0;
}

int Extractor::input_pixels(const char* blob_name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, const float* mean_vals,
                            const float* norm_vals)
{
    return input_pixels_resize(blob_name, pixels, type, w, h, stride, n, nstride, 0, 0, mean_vals, norm_vals);
}

int Extractor::input_pixels_resize(const char* blob_name, const unsigned char* pixels, int type, int w, int h, int stride, int n, size_t nstride, int target_w, int target_h,
                                   const float* mean_vals, const float* norm_vals)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1 || !pixels) return -1;
    if (reject_folded_blob(d->net, blob_index, "input")) return -1;
    ExtractorPrivate::PixelInput pi;
    pi.blob_index = blob_index;
    pi.pixels = pixels;
    pi.type = type;
    pi.w = w;
    pi.h = h;
    pi.stride = stride;
    pi.target_w = target_w;
    pi.target_h = target_h;
    pi.n = n < 1 ? 1 : n;
    pi.nstride = nstride;
    pi.has_mean = mean_vals != 0;
    pi.has_norm = norm_vals != 0;
    for (int i = 0; i < 4; i++)
    {
        pi.mean_vals[i] = 0.f;
        pi.norm_vals[i] = 1.f;
    }
    const int from = type & 0xffff;
    const int channels = from == 3 ? 1 : (from == 4 || from == 5 ? 4 : 3);
    for (int i = 0; i < channels; i++)
    {
        if (mean_vals) pi.mean_vals[i] = mean_vals[i];
        if (norm_vals) pi.norm_vals[i] = norm_vals[i];
    }
    d->blob_mats[blob_index].release();
    d->blob_mats_gpu[blob_index].release();
    for (size_t i = 0; i < d->pixel_inputs.size(); i++)
        if (d->pixel_inputs[i].blob_index == blob_index)
        {
            d->pixel_inputs[i] = pi;
            return 0;
        }
    d->pixel_inputs.push_back(pi);
    return 0;
}

int Extractor::extract(const char* blob_name, Mat& feat, int type)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1)
    {
        NCNN_LOGE("Try");
        const std::vector<const char*>& names = d->net->output_names();
        for (size_t i = 0; i < names.size(); i++) NCNN_LOGE("    ex.extract(\"%s\", out%d);", names[i], (int)i);
        return -1;
    }
    return extract(blob_index, feat, type);
}

// src/net.cpp:2855-3025, Vulkan branch :2885-2933
int Extractor::extract(int blob_index, Mat& feat, int /*type*/)
{
    if (blob_index < 0 || blob_index >= (int)d->blob_mats.size()) return -1;
    if (reject_folded_blob(d->net, blob_index, "extract")) return -1;
    if (!d->blob_mats[blob_index].empty())
    {
        feat = d->blob_mats[blob_index];
        return 0;
    }
    CudaContext* ctx = d->context();
    if (!ctx)
    {
        NCNN_LOGE("no CUDA device available: %s", ncnn_cuda_last_error());
        return -1;
    }
    int ret;
    {
        CudaCompute cmd(ctx);
        CudaMat feat_gpu;
        ret = extract(blob_index, feat_gpu, cmd);
        if (ret == 0)
        {
            ret = cmd.record_download(feat_gpu, d->blob_mats[blob_index], d->opt);
        }
        int sret = cmd.submit_and_wait(); // the single device sync of an extract (src/net.cpp:2912-2916)
        if (ret == 0) ret = sret;
        d->h2d = cmd.h2d_bytes;
        d->d2h = cmd.d2h_bytes;
    }
    if (ret != 0) return ret;
    feat = d->blob_mats[blob_index];
    return 0;
}

int Extractor::extract(const char* blob_name, CudaMat& feat, CudaCompute& cmd)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1) return -1;
    return extract(blob_index, feat, cmd);
}

// src/net.cpp:3083-3116
int Extractor::extract(int blob_index, CudaMat& feat, CudaCompute& cmd)
{
    if (blob_index < 0 || blob_index >= (int)d->blob_mats.size()) return -1;
    if (reject_folded_blob(d->net, blob_index, "extract")) return -1;
    int ret = 0;
    // pixel inputs first: upload the raw bytes, convert + normalise on the device
    for (size_t i = 0; i < d->pixel_inputs.size(); i++)
    {
        const ExtractorPrivate::PixelInput& pi = d->pixel_inputs[i];
        ret = cmd.record_upload_pixels(pi.pixels, pi.type, pi.w, pi.h, pi.stride, pi.n, pi.nstride, pi.has_mean ? pi.mean_vals : 0, pi.has_norm ? pi.norm_vals : 0,
                                       d->blob_mats_gpu[pi.blob_index], d->opt, pi.target_w, pi.target_h);
        if (ret != 0) return ret;
    }
    d->pixel_inputs.clear();
    if (d->blob_mats_gpu[blob_index].empty())
    {
        if (!d->blob_mats[blob_index].empty())
        {
            ret = cmd.record_upload(d->blob_mats[blob_index], d->blob_mats_gpu[blob_index], d->opt);
        }
        else
        {
            int layer_index = d->net->blobs()[blob_index].producer;
            if (layer_index < 0)
            {
                NCNN_LOGE("blob %s has no producer", d->net->blobs()[blob_index].name.c_str());
                return -1;
            }
            ret = d->net->d->forward_layer(layer_index, d->blob_mats, d->blob_mats_gpu, cmd, d->opt);
        }
    }
    if (ret != 0) return ret;
    feat = d->blob_mats_gpu[blob_index];
    return 0;
}

int Extractor::extract_yolov8_proposals(const char* blob_name, const int* strides, int num_strides, int in_w, int in_h, float prob_threshold, Mat& proposals)
{
    int blob_index = d->net->find_blob_index_by_name(blob_name);
    if (blob_index == -1) return -1;
    CudaContext* ctx = d->context();
    if (!ctx)
    {
        NCNN_LOGE("no CUDA device available: %s", ncnn_cuda_last_error());
        return -1;
    }
    int ret;
    {
        CudaCompute cmd(ctx);
        CudaMat pred, decoded;
        ret = extract(blob_index, pred, cmd);
        if (ret == 0 && pred.dims != 2) ret = -1;
        if (ret == 0)
        {
            decoded.create(6, pred.h, NCNN_CUDA_F32, pred.n, cmd.blob_allocator(d->opt));
            if (decoded.empty()) ret = -100;
        }
        if (ret == 0)
        {
            ncnn_cuda_tensor p = pred.view(), q = decoded.view();
            ret = ncnn_cuda_yolov8_decode(&p, strides, num_strides, in_w, in_h, prob_threshold, &q, cmd.stream());
        }
        if (ret == 0) ret = cmd.record_download(decoded, proposals, d->opt);
        int sret = cmd.submit_and_wait();
        if (ret == 0) ret = sret;
        d->h2d = cmd.h2d_bytes;
        d->d2h = cmd.d2h_bytes;
    }
    return ret;
}

size_t Extractor::last_h2d_bytes() const
{
    return d->h2d;
}
size_t Extractor::last_d2h_bytes() const
{
    return d->d2h;
}

} // namespace ncnn
