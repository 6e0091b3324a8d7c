"""oracle/yolov8_decode.py -- numpy restatement of the YOLOv8 post-processing of the reference's examples/yolov8.cpp.

TEST INFRASTRUCTURE ONLY (imported by tests/ as the checker of ncnn_cuda_yolov8_decode; never by the product).

This step lives in the reference's example program, not in its library, and the reference holds no test, fixture or golden
vector for it.  PINNED since round 2 against the reference's own code: oracle/build_ref.py compiles examples/yolov8.cpp where it
lies into oracle/_ref (oracle/yolov8_example_driver.cpp includes it as a translation unit, OpenCV replaced by the reference's
simpleocv.h through the example's USE_NCNN_SIMPLEOCV switch) and tests/test_oracle_golden.py::test_yolov8_decode_restatement_
matches_reference_example checks every function below against it on seeded inputs.  The Softmax the example runs over the 4 x 16
box logits is the library layer (src/layer/softmax.cpp: subtract the row maximum, exp, divide by the sum), restated here in fp32.

    generate_proposals  examples/yolov8.cpp:160-254 (one stride) and :256-273 (all strides, rows stride by stride)
    sigmoid             examples/yolov8.cpp:155-158
    nms_sorted_bboxes   examples/yolov8.cpp:118-153 (class-aware unless agnostic)
"""
import numpy as np


def generate_proposals(pred, strides, in_w, in_h, prob_threshold):
    """pred: (anchors, 64 + num_class) fp32 of ONE image -> (anchors, 6) fp32 rows {x, y, w, h, prob, label}; rows below the
    threshold are {0, 0, 0, 0, 0, -1} (the example simply does not push them; the dense form keeps anchor order)."""
    pred = np.asarray(pred, np.float32)
    out = np.zeros((pred.shape[0], 6), np.float32)
    out[:, 5] = -1
    reg = np.arange(16, dtype=np.float32)
    row = 0
    for stride in strides:
        gw, gh = in_w // stride, in_h // stride            # :165-166
        for y in range(gh):
            for x in range(gw):
                p = pred[row]
                scores = p[64:]
                label = int(np.argmax(scores))               # :177-191, first maximum (strict '>')
                score = np.float32(1.0) / (np.float32(1.0) + np.exp(-scores[label], dtype=np.float32))  # :193
                if score >= np.float32(prob_threshold):      # :196
                    box = p[:64].reshape(4, 16)              # :198
                    e = np.exp(box - box.max(axis=1, keepdims=True), dtype=np.float32)
                    sm = e / e.sum(axis=1, keepdims=True, dtype=np.float32)        # Softmax axis=1, :200-219
                    ltrb = (sm * reg).sum(axis=1, dtype=np.float32) * np.float32(stride)  # :221-232
                    cx, cy = np.float32((x + 0.5) * stride), np.float32((y + 0.5) * stride)  # :234-235
                    x0, y0, x1, y1 = cx - ltrb[0], cy - ltrb[1], cx + ltrb[2], cy + ltrb[3]  # :237-240
                    out[row] = (x0, y0, x1 - x0, y1 - y0, score, label)                      # :242-250
                row += 1
    assert row == pred.shape[0], (row, pred.shape)
    return out


def nms_sorted_bboxes(objs, nms_threshold, agnostic=False):
    """objs: (k, 6) rows sorted by prob descending -> indices kept (examples/yolov8.cpp:118-153)"""
    areas = objs[:, 2] * objs[:, 3]
    picked = []
    for i in range(objs.shape[0]):
        keep = True
        for j in picked:
            if not agnostic and objs[i, 5] != objs[j, 5]:
                continue
            iw = min(objs[i, 0] + objs[i, 2], objs[j, 0] + objs[j, 2]) - max(objs[i, 0], objs[j, 0])
            ih = min(objs[i, 1] + objs[i, 3], objs[j, 1] + objs[j, 3]) - max(objs[i, 1], objs[j, 1])
            inter = max(iw, 0.0) * max(ih, 0.0)               # cv::Rect_ & : empty intersection has zero area
            union = areas[i] + areas[j] - inter
            if inter / union > nms_threshold:
                keep = False
        if keep:
            picked.append(i)
    return picked
