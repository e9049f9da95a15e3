#!/bin/bash
# the narrow convolution layers one by one (CUDA-event times), with the A/B knobs of this round's epilogue work
run() {
  python tools/probes/conv_layer_probe.py 27 357 637 32 32 3; python tools/probes/conv_layer_probe.py 27 357 637 32 64 3
  python tools/probes/conv_layer_probe.py 107 180 320 64 64 3; python tools/probes/conv_layer_probe.py 107 180 320 64 64 3 --residual
  python tools/probes/conv_layer_probe.py 16 360 640 64 128 3; python tools/probes/conv_layer_probe.py 16 360 640 128 128 3
}
echo "production path"; run
echo "DIN_CONV_BIAS_TC=0"; DIN_CONV_BIAS_TC=0 run
echo "DIN_CONV_DIRECT_STORE=0"; DIN_CONV_DIRECT_STORE=0 run
./build/tma_rate_probe 2>/dev/null || echo "(build/tma_rate_probe not built: see the header of tools/probes/tma_rate_probe.cu)"
