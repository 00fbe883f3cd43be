#!/bin/bash
mkdir -p gpurun_out/r5
timeout 600 python profiles/scripts/debug_attn_dropout.py > gpurun_out/r5/debug_attn_dropout.log 2>&1
cat gpurun_out/r5/debug_attn_dropout.log | cut -c1-330
