#include "paramdict.h"

#include <ctype.h>
#include <stdlib.h>
#include <vector>

#include "datareader.h"

namespace ncnn {

ParamDict::ParamDict()
{
    clear();
}

void ParamDict::clear()
{
    for (int i = 0; i < NCNN_MAX_PARAM_COUNT; i++)
    {
        params_[i].type = 0;
        params_[i].i = 0;
        params_[i].v = Mat();
        params_[i].s.clear();
    }
}

int ParamDict::type(int id) const
{
    return params_[id].type;
}
int ParamDict::get(int id, int def) const
{
    return params_[id].type ? params_[id].i : def;
}
float ParamDict::get(int id, float def) const
{
    return params_[id].type ? params_[id].f : def;
}
Mat ParamDict::get(int id, const Mat& def) const
{
    return params_[id].type ? params_[id].v : def;
}
std::string ParamDict::get(int id, const std::string& def) const
{
    return params_[id].type ? params_[id].s : def;
}
void ParamDict::set(int id, int i)
{
    params_[id].type = 2;
    params_[id].i = i;
}
void ParamDict::set(int id, float f)
{
    params_[id].type = 3;
    params_[id].f = f;
}
void ParamDict::set(int id, const Mat& v)
{
    params_[id].type = 4;
    params_[id].v = v;
}
void ParamDict::set(int id, const std::string& s)
{
    params_[id].type = 7;
    params_[id].s = s;
}

static bool token_is_float(const std::string& t)
{
    for (size_t j = 0; j < t.size(); j++)
        if (t[j] == '.' || tolower(t[j]) == 'e') return true;
    return false;
}

static float token_to_float(const std::string& t)
{
    return (float)strtod(t.c_str(), 0);
}

int ParamDict::load_param_text(const char* p, const char* end)
{
    clear();
    while (p < end)
    {
        while (p < end && isspace((unsigned char)*p)) p++;
        if (p >= end) break;
        // id
        char* q = 0;
        long id = strtol(p, &q, 10);
        if (q == p || q >= end || *q != '=')
        {
            NCNN_LOGE("ParamDict: malformed key near '%.16s'", p);
            return -1;
        }
        p = q + 1;
        bool old_array = id <= -23300;
        if (old_array) id = -id - 23300;
        if (id < 0 || id >= NCNN_MAX_PARAM_COUNT)
        {
            NCNN_LOGE("id < NCNN_MAX_PARAM_COUNT failed (id=%ld, NCNN_MAX_PARAM_COUNT=%d)", id, NCNN_MAX_PARAM_COUNT);
            return -1;
        }
        // value token: up to whitespace, except quoted strings which run to the closing quote
        const char* vb = p;
        if (p < end && *p == '"')
        {
            p++;
            while (p < end && *p != '"' && *p != '\n') p++;
            std::string s(vb + 1, p);
            if (p < end && *p == '"') p++;
            params_[id].type = 7;
            params_[id].s = s;
            continue;
        }
        while (p < end && !isspace((unsigned char)*p)) p++;
        std::string tok(vb, p);
        if (tok.empty())
        {
            NCNN_LOGE("ParamDict read value failed");
            return -1;
        }
        if (!old_array && isalpha((unsigned char)tok[0]))
        {
            params_[id].type = 7;
            params_[id].s = tok;
            continue;
        }
        // split on commas
        std::vector<std::string> parts;
        {
            size_t b = 0;
            for (size_t j = 0; j <= tok.size(); j++)
                if (j == tok.size() || tok[j] == ',')
                {
                    parts.push_back(tok.substr(b, j - b));
                    b = j + 1;
                }
        }
        if (old_array)
        {
            int len = atoi(parts[0].c_str());
            if (len < 0 || (int)parts.size() != len + 1)
            {
                NCNN_LOGE("ParamDict read array element failed");
                return -1;
            }
            parts.erase(parts.begin());
        }
        if (old_array || parts.size() > 1)
        {
            int len = (int)parts.size();
            bool any_float = false;
            for (int j = 0; j < len; j++) any_float = any_float || token_is_float(parts[j]);
            Mat v(len);
            if (len > 0 && v.empty()) return -100;
            // the reference types the array by its LAST element (old style) / FIRST element (new style); mixed arrays
            // do not occur in valid files, so one decision for the whole array is equivalent
            for (int j = 0; j < len; j++)
            {
                if (any_float)
                    ((float*)v.data)[j] = token_to_float(parts[j]);
                else
                    ((int*)v.data)[j] = atoi(parts[j].c_str());
            }
            params_[id].type = any_float ? 6 : 5;
            params_[id].v = v;
            continue;
        }
        if (token_is_float(tok))
        {
            params_[id].type = 3;
            params_[id].f = token_to_float(tok);
        }
        else
        {
            params_[id].type = 2;
            params_[id].i = atoi(tok.c_str());
        }
    }
    return 0;
}

int ParamDict::load_param_bin(const DataReader& dr)
{
    clear();
    // binary 0: id int32, value int32/float32 ; array: id = -23300 - id, len int32, len * 4 bytes ; end: -233
    int id = 0;
    if (dr.read(&id, sizeof(int)) != sizeof(int)) return -1;
    while (id != -233)
    {
        bool is_array = id <= -23300;
        if (is_array) id = -id - 23300;
        if (id < 0 || id >= NCNN_MAX_PARAM_COUNT)
        {
            NCNN_LOGE("id < NCNN_MAX_PARAM_COUNT failed (id=%d, NCNN_MAX_PARAM_COUNT=%d)", id, NCNN_MAX_PARAM_COUNT);
            return -1;
        }
        if (is_array)
        {
            int len = 0;
            if (dr.read(&len, sizeof(int)) != sizeof(int) || len < 0) return -1;
            Mat v(len > 0 ? len : 1);
            v.w = len;
            if (len > 0 && dr.read(v.data, sizeof(float) * len) != sizeof(float) * len) return -1;
            params_[id].type = 4;
            params_[id].v = v;
        }
        else
        {
            if (dr.read(&params_[id].i, sizeof(int)) != sizeof(int)) return -1;
            params_[id].type = 1;
        }
        if (dr.read(&id, sizeof(int)) != sizeof(int)) return -1;
    }
    return 0;
}

} // namespace ncnn
