#!/bin/bash
# refill threshold of the traversal loop re-tuned for the packed / 96-byte-node kernel
TUNE_VARIANTS=${V:-17,16} TUNE_THRESHOLDS=24,26,28,30,32 timeout 900 python tools/gpu_tune.py 2>&1 | grep -vE "^build|library" | tee gpurun_out/${OUT:-thr.log}
