// platform.h -- build switches and logging of the host runtime (counterpart of the reference's generated
// src/platform.h.in:7-63, :463-476; only what the CUDA hot path needs).
#ifndef NCNN_B200_PLATFORM_H
#define NCNN_B200_PLATFORM_H

#include <stdio.h>

#define NCNN_STDIO 1
#define NCNN_STRING 1
#define NCNN_BATCH 1
#define NCNN_CUDA 1

#define NCNN_LOGE(...)                \
    do                                \
    {                                 \
        fprintf(stderr, ##__VA_ARGS__); \
        fprintf(stderr, "\n");        \
    } while (0)

#if defined(__GNUC__)
#define NCNN_EXPORT __attribute__((visibility("default")))
#else
#define NCNN_EXPORT
#endif

#endif // NCNN_B200_PLATFORM_H
