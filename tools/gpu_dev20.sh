#!/bin/bash
for cfg in "256 0" "32 8" "32 1" "16 8" "16 1" "64 1"; do
  set -- $cfg
  echo "=== batch $1 rotate $2"
  python tools/conv_layers.py --batch $1 --rotate $2 --only "s2|s3 1x1 128|s3 3x3|s4 1x1 256|s4 3x3" 2>&1 | grep -v "^#"
done
