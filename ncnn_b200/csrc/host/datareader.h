// datareader.h -- byte source abstraction for .param / .bin (reference: src/datareader.h:21-83)
#ifndef NCNN_B200_DATAREADER_H
#define NCNN_B200_DATAREADER_H

#include <stddef.h>
#include <stdio.h>

#include "platform.h"

namespace ncnn {

class NCNN_EXPORT DataReader
{
public:
    DataReader();
    virtual ~DataReader();
    // parse plain param text; return 1 if scan success
    virtual int scan(const char* format, void* p) const;
    // read binary param and model data; return bytes read
    virtual size_t read(void* buf, size_t size) const;
    // get model data reference; return bytes referenced
    virtual size_t reference(size_t size, const void** buf) const;
};

class NCNN_EXPORT DataReaderFromStdio : public DataReader
{
public:
    explicit DataReaderFromStdio(FILE* fp);
    virtual ~DataReaderFromStdio();
    virtual int scan(const char* format, void* p) const;
    virtual size_t read(void* buf, size_t size) const;

private:
    FILE* fp_;
};

class NCNN_EXPORT DataReaderFromMemory : public DataReader
{
public:
    explicit DataReaderFromMemory(const unsigned char*& mem);
    virtual ~DataReaderFromMemory();
    virtual int scan(const char* format, void* p) const;
    virtual size_t read(void* buf, size_t size) const;
    virtual size_t reference(size_t size, const void** buf) const;

private:
    const unsigned char*& mem_;
};

} // namespace ncnn

#endif // NCNN_B200_DATAREADER_H
