#!/bin/bash
# GPU session: tensor core alone (no trunk epilogue), centre row's A operand from shared memory / from tensor memory
for e in 8_false 8_true; do echo "exp $e"; AZ_ENGINE_LIB=$PWD/_ab/libaz_exp$e.so python scripts/profile_net_step.py 16384 resnet4x64:v4 resnet8x64:v4 2>&1 | grep "E="; done
