#include "datareader.h"

#include <string.h>

namespace ncnn {

DataReader::DataReader()
{
}
DataReader::~DataReader()
{
}
int DataReader::scan(const char*, void*) const
{
    return 0;
}
size_t DataReader::read(void*, size_t) const
{
    return 0;
}
size_t DataReader::reference(size_t, const void**) const
{
    return 0;
}

DataReaderFromStdio::DataReaderFromStdio(FILE* fp)
    : fp_(fp)
{
}
DataReaderFromStdio::~DataReaderFromStdio()
{
}
int DataReaderFromStdio::scan(const char* format, void* p) const
{
    return fscanf(fp_, format, p);
}
size_t DataReaderFromStdio::read(void* buf, size_t size) const
{
    return fread(buf, 1, size, fp_);
}

DataReaderFromMemory::DataReaderFromMemory(const unsigned char*& mem)
    : mem_(mem)
{
}
DataReaderFromMemory::~DataReaderFromMemory()
{
}
int DataReaderFromMemory::scan(const char* format, void* p) const
{
    size_t fmtlen = strlen(format);
    char* format_with_n = new char[fmtlen + 4];
    sprintf(format_with_n, "%s%%n", format);
    int nconsumed = 0;
    int nscan = sscanf((const char*)mem_, format_with_n, p, &nconsumed);
    mem_ += nconsumed;
    delete[] format_with_n;
    return nconsumed > 0 ? nscan : 0;
}
size_t DataReaderFromMemory::read(void* buf, size_t size) const
{
    memcpy(buf, mem_, size);
    mem_ += size;
    return size;
}
size_t DataReaderFromMemory::reference(size_t size, const void** buf) const
{
    *buf = mem_;
    mem_ += size;
    return size;
}

} // namespace ncnn
