// yolov8_example_driver.cpp -- TEST INFRASTRUCTURE (oracle): C entry points over the reference's OWN YOLOv8 post-processing,
// compiled from where it lies (/root/reference/examples/yolov8.cpp, included below -- no source is copied into the repository).
//
// The reference keeps generate_proposals / qsort_descent_inplace / nms_sorted_bboxes as file-static functions of its example
// program (examples/yolov8.cpp:67-273); there is no test or fixture for them upstream.  Including the example as a translation
// unit (its main() renamed, OpenCV replaced by the reference's own simpleocv.h through the example's USE_NCNN_SIMPLEOCV switch)
// makes those very functions callable, so that ncnn_cuda_yolov8_decode and the numpy restatement oracle/yolov8_decode.py are
// pinned against the reference's code instead of against a reading of it.
//
// Built by oracle/build_ref.py with -DUSE_NCNN_SIMPLEOCV -DNCNN_SIMPLEOCV=1 into oracle/_ref/libncnn_ref_<isa>.so.
#define main ncnn_example_yolov8_main
#ifndef NCNN_REFERENCE_YOLOV8_CPP
#define NCNN_REFERENCE_YOLOV8_CPP "/root/reference/examples/yolov8.cpp"
#endif
#include NCNN_REFERENCE_YOLOV8_CPP
#undef main

extern "C" {

// pred: rows x cols fp32 (cols = 64 + num_class, rows ordered stride by stride, y-major), one image.
// out: up to max_out records of 6 floats {x, y, width, height, prob, label} in generation order; returns the number of proposals
// the reference produced (may exceed max_out: then only the first max_out were written).
__attribute__((visibility("default"))) int ref_yolov8_generate_proposals(const float* pred, int rows, int cols, const int* strides, int num_strides, int in_w, int in_h,
                                                                         float prob_threshold, float* out, int max_out)
{
    ncnn::Mat p(cols, rows);
    for (int y = 0; y < rows; y++) memcpy(p.row(y), pred + (size_t)y * cols, sizeof(float) * cols);
    ncnn::Mat in_pad;
    in_pad.w = in_w;
    in_pad.h = in_h;
    std::vector<int> st(strides, strides + num_strides);
    std::vector<Object> objects;
    generate_proposals(p, st, in_pad, prob_threshold, objects);
    int n = 0;
    for (size_t i = 0; i < objects.size() && n < max_out; i++, n++)
    {
        float* o = out + (size_t)n * 6;
        o[0] = objects[i].rect.x;
        o[1] = objects[i].rect.y;
        o[2] = objects[i].rect.width;
        o[3] = objects[i].rect.height;
        o[4] = objects[i].prob;
        o[5] = (float)objects[i].label;
    }
    return (int)objects.size();
}

// the reference's sort + NMS (examples/yolov8.cpp:73-153) over `count` records {x, y, w, h, prob, label}: writes the indices (into the
// SORTED order) that survive to picked[] and the sorted records back to boxes; returns how many were picked.
__attribute__((visibility("default"))) int ref_yolov8_sort_nms(float* boxes, int count, float nms_threshold, int agnostic, int* picked, int max_picked)
{
    std::vector<Object> objects(count);
    for (int i = 0; i < count; i++)
    {
        objects[i].rect.x = boxes[i * 6 + 0];
        objects[i].rect.y = boxes[i * 6 + 1];
        objects[i].rect.width = boxes[i * 6 + 2];
        objects[i].rect.height = boxes[i * 6 + 3];
        objects[i].prob = boxes[i * 6 + 4];
        objects[i].label = (int)boxes[i * 6 + 5];
    }
    qsort_descent_inplace(objects);
    std::vector<int> keep;
    nms_sorted_bboxes(objects, keep, nms_threshold, agnostic != 0);
    for (int i = 0; i < count; i++)
    {
        boxes[i * 6 + 0] = objects[i].rect.x;
        boxes[i * 6 + 1] = objects[i].rect.y;
        boxes[i * 6 + 2] = objects[i].rect.width;
        boxes[i * 6 + 3] = objects[i].rect.height;
        boxes[i * 6 + 4] = objects[i].prob;
        boxes[i * 6 + 5] = (float)objects[i].label;
    }
    int n = 0;
    for (size_t i = 0; i < keep.size() && n < max_picked; i++, n++) picked[n] = keep[i];
    return (int)keep.size();
}

} // extern "C"
