// modelbin.h -- weight reader (reference: src/modelbin.h, src/modelbin.cpp:82-338).
// type 0 loads start with a 4-byte tag: 0 raw fp32, 0x01306B47 fp16, 0x000D4B38 int8, 0x0002C056 raw fp32,
// any other non-zero byte sum = 256-float codebook + u8 indices; type 1 loads are raw fp32 with no tag.
#ifndef NCNN_B200_MODELBIN_H
#define NCNN_B200_MODELBIN_H

#include "mat.h"

namespace ncnn {

class DataReader;

class NCNN_EXPORT ModelBin
{
public:
    ModelBin();
    virtual ~ModelBin();
    virtual Mat load(int w, int type) const = 0;
    virtual Mat load(int w, int h, int type) const;
    virtual Mat load(int w, int h, int c, int type) const;
    virtual Mat load(int w, int h, int d, int c, int type) const;
};

class NCNN_EXPORT ModelBinFromDataReader : public ModelBin
{
public:
    explicit ModelBinFromDataReader(const DataReader& dr);
    virtual ~ModelBinFromDataReader();
    virtual Mat load(int w, int type) const;

private:
    const DataReader& dr_;
};

class NCNN_EXPORT ModelBinFromMatArray : public ModelBin
{
public:
    // weights is an array of already-loaded Mats consumed in order
    explicit ModelBinFromMatArray(const Mat* weights);
    virtual ~ModelBinFromMatArray();
    virtual Mat load(int w, int type) const;

private:
    mutable const Mat* weights_;
};

} // namespace ncnn

#endif // NCNN_B200_MODELBIN_H
