#!/bin/bash
L=din-group-activity-recognition-benchmark_b200
cp $L/libdin_sm100.so /tmp/keep.so; cp $L/libdin_sm100_roles.so $L/libdin_sm100.so
export DIN_FUSED_DEBUG=1
python tools/probes/conv_layer_probe.py 27 357 637 32 32 3; python tools/probes/conv_layer_probe.py 27 357 637 32 64 3; python tools/probes/conv_layer_probe.py 107 180 320 64 64 3; python tools/probes/conv_layer_probe.py 107 180 320 64 64 3 --residual; python tools/probes/conv_layer_probe.py 16 360 640 64 128 3; python tools/probes/conv_layer_probe.py 16 360 640 128 128 3
cp /tmp/keep.so $L/libdin_sm100.so
