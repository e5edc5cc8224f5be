#!/bin/bash
# round 2, GPU call H: launch-shape A/B of the radix select passes.
mkdir -p gpurun_out
timeout 300 python tools/select_ab.py > gpurun_out/r2h_select_ab.txt 2>&1; cat gpurun_out/r2h_select_ab.txt | tail -14
