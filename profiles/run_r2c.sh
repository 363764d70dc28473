#!/bin/bash
# round 2, trip c: the whole GPU suite with full output (the previous trip's pytest died with a fatal-error dump whose head was cut off)
set -u
O=gpurun_out
mkdir -p $O
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2c_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest_gpu.txt
head -60 $O/r2c_pytest_gpu.txt
tail -5 $O/r2c_pytest_gpu.txt
