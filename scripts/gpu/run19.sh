#!/bin/bash
# GPU session: what the epilogue of k_resnet_wide pays for (timing-only builds: results are wrong by construction)
# exp bit 0: no shuffles; bit 1: one third of the tensor-memory loads
for e in 1 2 3; do echo "exp $e"; AZ_ENGINE_LIB=$PWD/_ab/libaz_exp$e.so python scripts/profile_net_step.py 16384 resnet4x64:v4 2>&1 | tail -1; done
