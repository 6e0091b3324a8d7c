#!/bin/bash
python tools/conv_layers.py --check --only "3x3" 2>&1 | tail -8
NCNN_B200_CONV_SHIFT=2 python tools/conv_layers.py --only "3x3" 2>&1 | tail -6
python tools/conv_layers.py --workload vgg16 --check 2>&1 | tail -16
NCNN_B200_CONV_SHIFT=2 python tools/conv_layers.py --workload vgg16 2>&1 | tail -16
