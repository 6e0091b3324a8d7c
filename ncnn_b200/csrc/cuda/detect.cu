// detect.cu -- the decode half of YOLOv8 post-processing on the device: what generate_proposals of the
// reference's examples/yolov8.cpp:160-273 does on the host after downloading the whole prediction blob.
// The head's output is a 2-D blob per image: one row per anchor point, 4 x 16 box-distribution logits followed by the class
// logits (w = 64 + num_class, h = sum of the grids of all strides).  One warp owns one anchor row: lanes scan the class
// logits for the first maximum, the score is its sigmoid; rows at or above the threshold run the four 16-bin softmaxes
// (distribution focal loss) inside 16-lane halves with shuffles and turn the expected distances into a box.  The result is
// ONE 6-float record per anchor (x, y, w, h, prob, label; prob = 0 and label = -1 below the threshold) in anchor order --
// deterministic, no atomics -- so the device -> host copy is 24 bytes per anchor instead of the (64 + num_class) * 4 of the
// prediction blob (24x fewer bytes for COCO's 80 classes).  Sorting and NMS stay with the caller: they run over the few
// survivors.  HBM-bound: reads the blob once.
#include "common.cuh"

using namespace ncnn_cuda;

namespace {

#define NC_YOLO_MAX_LEVELS 8

struct YoloGeom
{
    int levels;
    int stride[NC_YOLO_MAX_LEVELS];
    int grid_w[NC_YOLO_MAX_LEVELS];
    int row_begin[NC_YOLO_MAX_LEVELS + 1]; // first anchor row of each level, and the total
    int num_class, n;
    int in_cpitch, out_cpitch;
    long long in_nstep, out_nstep;
    float prob_threshold;
};

template<typename T>
__global__ void __launch_bounds__(256) yolov8_decode_kernel(const T* __restrict__ pred, float* __restrict__ out, YoloGeom g)
{
    NC_PDL_PROLOGUE();
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int rows = g.row_begin[g.levels];
    const long long total = (long long)rows * g.n;
    for (long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < total; o += warps)
    {
        const int b = (int)(o / rows);
        const int r = (int)(o - (long long)b * rows);
        const T* row = pred + b * g.in_nstep + (long long)r * g.in_cpitch;
        // first maximum of the class logits (examples/yolov8.cpp:177-191: strict '>' keeps the lowest index among equals)
        float best = -FLT_MAX;
        int label = -1;
        for (int k = lane; k < g.num_class; k += 32)
        {
            const float s = to_f32(row[64 + k]);
            if (s > best)
            {
                best = s;
                label = k;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1)
        {
            const float ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int ol = __shfl_xor_sync(0xffffffffu, label, s);
            if (ol >= 0 && (ob > best || (ob == best && (label < 0 || ol < label))))
            {
                best = ob;
                label = ol;
            }
        }
        const float score = 1.0f / (1.0f + expf(-best)); // :155-158, :193
        float* rec = out + b * g.out_nstep + (long long)r * g.out_cpitch;
        if (!(score >= g.prob_threshold)) // warp-uniform
        {
            if (lane < 6) rec[lane] = lane == 5 ? -1.f : 0.f;
            continue;
        }
        // the four sides' 16-bin softmax (Softmax layer over each row of the 4 x 16 view, :198-219) and expected distance
        // (:221-232): lanes 0-15 / 16-31 hold sides 0 / 1, then sides 2 / 3
        float dist[2];
#pragma unroll
        for (int pass = 0; pass < 2; pass++)
        {
            const float v = to_f32(row[pass * 32 + lane]);
            float m = v;
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
            const float e = expf(v - m);
            float sum = e;
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
            float d = (float)(lane & 15) * (e / sum);
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
            dist[pass] = d;
        }
        const float left = __shfl_sync(0xffffffffu, dist[0], 0), top = __shfl_sync(0xffffffffu, dist[0], 16);
        const float right = __shfl_sync(0xffffffffu, dist[1], 0), bottom = __shfl_sync(0xffffffffu, dist[1], 16);
        if (lane == 0)
        {
            int lv = 0;
            while (lv + 1 < g.levels && r >= g.row_begin[lv + 1]) lv++;
            const int cell = r - g.row_begin[lv];
            const int gy = cell / g.grid_w[lv], gx = cell - gy * g.grid_w[lv];
            const float st = (float)g.stride[lv];
            const float cx = (gx + 0.5f) * st, cy = (gy + 0.5f) * st; // :234-240
            const float x0 = cx - left * st, y0 = cy - top * st, x1 = cx + right * st, y1 = cy + bottom * st;
            rec[0] = x0;
            rec[1] = y0;
            rec[2] = x1 - x0;
            rec[3] = y1 - y0;
            rec[4] = score;
            rec[5] = (float)label;
        }
    }
}

template<typename T>
static int run_decode(const YoloGeom& g, const ncnn_cuda_tensor* pred, const ncnn_cuda_tensor* out, cudaStream_t stream)
{
    const long long total = (long long)g.row_begin[g.levels] * g.n;
    NC_PDL_LAUNCH((yolov8_decode_kernel<T>), grid_for(total * 32, 256, 16), 256, 0, stream, (const T*)pred->data, (float*)out->data, g);
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace

extern "C" int ncnn_cuda_yolov8_decode(const ncnn_cuda_tensor* pred, const int* strides, int num_strides, int in_w, int in_h, float prob_threshold,
                                       const ncnn_cuda_tensor* proposals, void* stream)
{
    NC_REQUIRE(pred && proposals && strides && pred->dims == 2 && proposals->dims == 2, "yolov8_decode: 2-D blobs required");
    NC_REQUIRE(num_strides >= 1 && num_strides <= NC_YOLO_MAX_LEVELS, "yolov8_decode: 1..8 strides");
    NC_REQUIRE(pred->w > 64, "yolov8_decode: a row holds 4 x 16 box logits and at least one class");
    NC_REQUIRE(proposals->elemtype == NCNN_CUDA_F32 && proposals->w == 6 && proposals->h == pred->h, "yolov8_decode: proposals must be fp32, w = 6, one row per anchor");
    YoloGeom g;
    g.levels = num_strides;
    int rows = 0;
    for (int i = 0; i < num_strides; i++)
    {
        NC_REQUIRE(strides[i] > 0, "yolov8_decode: bad stride");
        g.stride[i] = strides[i];
        g.grid_w[i] = in_w / strides[i]; // examples/yolov8.cpp:165-166
        g.row_begin[i] = rows;
        rows += (in_w / strides[i]) * (in_h / strides[i]);
        NC_REQUIRE(g.grid_w[i] > 0, "yolov8_decode: stride larger than the input");
    }
    g.row_begin[num_strides] = rows;
    NC_REQUIRE(rows == pred->h, "yolov8_decode: the grids of the strides do not add up to the blob's rows");
    g.num_class = pred->w - 64;
    g.n = pred->n < 1 ? 1 : pred->n;
    NC_REQUIRE((proposals->n < 1 ? 1 : proposals->n) == g.n, "yolov8_decode: batch mismatch");
    g.in_cpitch = pred->cpitch;
    g.out_cpitch = proposals->cpitch;
    g.in_nstep = pred->nstep;
    g.out_nstep = proposals->nstep;
    g.prob_threshold = prob_threshold;
    if (rows == 0) return 0;
    switch (pred->elemtype)
    {
    case NCNN_CUDA_F32: return run_decode<float>(g, pred, proposals, as_stream(stream));
    case NCNN_CUDA_BF16: return run_decode<__nv_bfloat16>(g, pred, proposals, as_stream(stream));
    case NCNN_CUDA_F16: return run_decode<__half>(g, pred, proposals, as_stream(stream));
    }
    return -1;
}
