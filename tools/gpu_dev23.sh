#!/bin/bash
mkdir -p gpurun_out
for d in _old .; do
(cd $d; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launch_$$.csv python bench.py --workload yolov8s --steps 1 --warmup 1 --no-extra-legs --no-cpu-baseline > /dev/null 2>&1; cp /tmp/launch_$$.csv /root/repo/gpurun_out/launch_yolo_$(basename $(pwd)).csv)
done
ls -la gpurun_out/launch_yolo_*
