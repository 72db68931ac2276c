#!/bin/bash
# live per-phase timeline of the fused call (CANVAS_DEBUG marks, CUDA events on the library stream)
tag=${1:-tl}
mkdir -p gpurun_out
CANVAS_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_timeline.err > gpurun_out/${tag}_timeline.json
grep "^\[fused\]" gpurun_out/${tag}_timeline.err | tail -26
