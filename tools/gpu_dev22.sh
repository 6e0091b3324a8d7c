#!/bin/bash
python bench.py --workload yolov8s --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench yolo', d['ms_per_step'])"
python bench.py --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench resnet', d['ms_per_step'])"
python bench.py --workload mobilenet_v2 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench mbv2', d['ms_per_step'])"
