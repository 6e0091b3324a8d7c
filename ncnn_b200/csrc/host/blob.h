// blob.h -- a named edge of the graph (reference: src/blob.h:12-28)
#ifndef NCNN_B200_BLOB_H
#define NCNN_B200_BLOB_H

#include <string>

#include "mat.h"

namespace ncnn {

class NCNN_EXPORT Blob
{
public:
    Blob()
        : producer(-1), consumer(-1)
    {
    }
    std::string name;
    int producer; // layer index which produces this blob
    int consumer; // layer index which consumes this blob
    Mat shape;    // shape hint (param id 30)
};

} // namespace ncnn

#endif // NCNN_B200_BLOB_H
