#ifndef NCNN_EXPORT_H
#define NCNN_EXPORT_H
#define NCNN_EXPORT __attribute__((visibility("default")))
#define NCNN_NO_EXPORT
#define NCNN_DEPRECATED
#endif
