#!/bin/bash
mkdir -p gpurun_out
PPNP_B200_LIB=$PWD/ppnp_b200/variants/libppnp_b200_fewends.so timeout 40 python tools/fewends_check.py > gpurun_out/fewends.log 2>&1
echo "rc=$?" >> gpurun_out/fewends.log
tail -c 2500 gpurun_out/fewends.log
