#!/bin/bash
mkdir -p gpurun_out
AZ_ENGINE_LIB=$PWD/_ab/libaz_trace.so python scripts/wide_trace.py 2 > gpurun_out/wide_trace.txt 2>&1; tail -3 gpurun_out/wide_trace.txt
