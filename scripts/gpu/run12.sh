#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/ubench/mma_pair > gpurun_out/mma_pair.txt 2>&1; cat gpurun_out/mma_pair.txt
