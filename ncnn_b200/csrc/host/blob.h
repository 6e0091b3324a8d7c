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
        : producer(-1), consumer(-1), folded_into(-1)
    {
    }
    std::string name;
    int producer; // layer index which produces this blob
    int consumer; // layer index which consumes this blob
    Mat shape;    // shape hint (param id 30)
    // load-time graph fusion (opt.use_cuda_graph_fusion): this blob was an intermediate folded away into layer `folded_into`
    // (a Convolution's pre-activation output, an Eltwise sum ...): it is never materialised, so it can be neither extracted
    // nor fed -- Extractor::input / extract on it fail with a message instead of silently doing nothing
    int folded_into;
};

} // namespace ncnn

#endif // NCNN_B200_BLOB_H
