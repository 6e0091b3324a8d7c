// paramdict.h -- 32-slot typed parameter dictionary of a layer line (reference: src/paramdict.h, parser
// src/paramdict.cpp:263-488 text, :491-600 binary).  Value kinds: int, float, int/float array, string.
#ifndef NCNN_B200_PARAMDICT_H
#define NCNN_B200_PARAMDICT_H

#include <string>

#include "mat.h"

#define NCNN_MAX_PARAM_COUNT 32

namespace ncnn {

class DataReader;

class NCNN_EXPORT ParamDict
{
public:
    ParamDict();
    // 0 null, 2 int, 3 float, 5 int array, 6 float array, 7 string (1 and 4: untyped scalar/array from .param.bin)
    int type(int id) const;
    int get(int id, int def) const;
    float get(int id, float def) const;
    Mat get(int id, const Mat& def) const;
    std::string get(int id, const std::string& def) const;
    void set(int id, int i);
    void set(int id, float f);
    void set(int id, const Mat& v);
    void set(int id, const std::string& s);
    void clear();

    // the `k=v ...` tail of one .param text line; [begin,end) may span further tokens
    int load_param_text(const char* begin, const char* end);
    // .param.bin record stream up to the -233 terminator
    int load_param_bin(const DataReader& dr);

private:
    struct Slot
    {
        int type;
        union
        {
            int i;
            float f;
        };
        Mat v;
        std::string s;
    };
    Slot params_[NCNN_MAX_PARAM_COUNT];
};

} // namespace ncnn

#endif // NCNN_B200_PARAMDICT_H
