#include "modelbin.h"

#include <stdint.h>
#include <vector>

#include "datareader.h"

namespace ncnn {

ModelBin::ModelBin()
{
}
ModelBin::~ModelBin()
{
}

Mat ModelBin::load(int w, int h, int type) const
{
    Mat m = load(w * h, type);
    if (m.empty()) return m;
    return m.reshape(w, h);
}

Mat ModelBin::load(int w, int h, int c, int type) const
{
    Mat m = load(w * h * c, type);
    if (m.empty()) return m;
    return m.reshape(w, h, c);
}

Mat ModelBin::load(int w, int h, int d, int c, int type) const
{
    Mat m = load(w * h * d * c, type);
    if (m.empty()) return m;
    return m.reshape(w, h, d, c);
}

static float half_to_float(unsigned short v)
{
    // IEEE binary16 -> binary32
    uint32_t sign = (uint32_t)(v & 0x8000) << 16;
    uint32_t exp = (v >> 10) & 0x1f;
    uint32_t man = v & 0x3ff;
    uint32_t out;
    if (exp == 0)
    {
        if (man == 0)
            out = sign;
        else
        {
            // subnormal: normalise
            int e = -1;
            do
            {
                e++;
                man <<= 1;
            } while ((man & 0x400) == 0);
            out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    }
    else if (exp == 31)
        out = sign | 0x7f800000u | (man << 13);
    else
        out = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f;
    memcpy(&f, &out, 4);
    return f;
}

ModelBinFromDataReader::ModelBinFromDataReader(const DataReader& dr)
    : dr_(dr)
{
}
ModelBinFromDataReader::~ModelBinFromDataReader()
{
}

Mat ModelBinFromDataReader::load(int w, int type) const
{
    Mat m;
    if (type == 0)
    {
        union
        {
            unsigned char f[4];
            unsigned int tag;
        } flag;
        size_t nread = dr_.read(&flag, sizeof(flag));
        if (nread != sizeof(flag))
        {
            NCNN_LOGE("ModelBin read flag_struct failed %zd", nread);
            return Mat();
        }
        unsigned int sum = (unsigned int)flag.f[0] + flag.f[1] + flag.f[2] + flag.f[3];
        if (flag.tag == 0x01306B47)
        {
            size_t bytes = alignSize(w * sizeof(unsigned short), 4);
            std::vector<unsigned short> h16(bytes / 2 + 1);
            nread = dr_.read(&h16[0], bytes);
            if (nread != bytes)
            {
                NCNN_LOGE("ModelBin read float16_weights failed %zd", nread);
                return Mat();
            }
            m.create(w);
            if (m.empty()) return m;
            float* p = m;
            for (int i = 0; i < w; i++) p[i] = half_to_float(h16[i]);
            return m;
        }
        if (flag.tag == 0x000D4B38)
        {
            NCNN_LOGE("ModelBin: int8 weights are outside the CUDA backend's scope (fp32/fp16 models only)");
            return Mat();
        }
        if (flag.tag == 0x0002C056 || sum == 0)
        {
            // a reader over memory that outlives the weights (a mapped model file, src/net.cpp:2263-2301) lends the bytes
            // instead of copying them (src/modelbin.cpp:180-196): the Mat is an external-data view, refcount NULL
            const void* refbuf = 0;
            if (dr_.reference(w * sizeof(float), &refbuf) == w * sizeof(float) && refbuf && ((size_t)refbuf & 3) == 0)
                return Mat(w, (void*)refbuf, (size_t)4u);
            m.create(w);
            if (m.empty()) return m;
            nread = dr_.read(m.data, w * sizeof(float));
            if (nread != w * sizeof(float))
            {
                NCNN_LOGE("ModelBin read weight_data failed %zd", nread);
                return Mat();
            }
            return m;
        }
        // 256-entry codebook + u8 indices (modelbin.cpp:216-258)
        float table[256];
        nread = dr_.read(table, sizeof(table));
        if (nread != sizeof(table))
        {
            NCNN_LOGE("ModelBin read quantization_value failed %zd", nread);
            return Mat();
        }
        size_t bytes = alignSize(w * sizeof(unsigned char), 4);
        std::vector<unsigned char> idx(bytes);
        nread = dr_.read(&idx[0], bytes);
        if (nread != bytes)
        {
            NCNN_LOGE("ModelBin read index_array failed %zd", nread);
            return Mat();
        }
        m.create(w);
        if (m.empty()) return m;
        float* p = m;
        for (int i = 0; i < w; i++) p[i] = table[idx[i]];
        return m;
    }
    if (type == 1)
    {
        m.create(w);
        if (m.empty()) return m;
        size_t nread = dr_.read(m.data, w * sizeof(float));
        if (nread != w * sizeof(float))
        {
            NCNN_LOGE("ModelBin read weight_data failed %zd", nread);
            return Mat();
        }
        return m;
    }
    NCNN_LOGE("ModelBin load type %d not implemented", type);
    return Mat();
}

ModelBinFromMatArray::ModelBinFromMatArray(const Mat* weights)
    : weights_(weights)
{
}
ModelBinFromMatArray::~ModelBinFromMatArray()
{
}

Mat ModelBinFromMatArray::load(int /*w*/, int /*type*/) const
{
    if (!weights_) return Mat();
    Mat m = weights_[0];
    weights_++;
    return m;
}

} // namespace ncnn
