#!/bin/bash
# GPU session: marginal cost of a residual block vs the fixed cost per batch (k_resnet_wide, variant 4; variant 2 beside it)
python scripts/profile_net_step.py 16384 resnet0x64:v4 resnet1x64:v4 resnet2x64:v4 resnet4x64:v4 resnet8x64:v4 resnet0x64:v2 resnet2x64:v2 resnet4x64:v2 2>&1 | grep "E="
echo exp1; AZ_ENGINE_LIB=$PWD/_ab/libaz_exp1.so python scripts/profile_net_step.py 16384 resnet0x64:v4 resnet2x64:v4 resnet4x64:v4 resnet8x64:v4 2>&1 | grep "E="
