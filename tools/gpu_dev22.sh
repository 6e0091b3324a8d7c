#!/bin/bash
python bench.py --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "^conv1\|res2a_branch2c\|res2b\|res3b\|res4b\|res5b\|pool5\|fc1000\|layers total"
