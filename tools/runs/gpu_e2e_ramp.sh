#!/bin/bash
# host-array call: short first / last chunks (ADYPT_HOST_RAMP, default on) against uniform chunks
timeout 600 python -m pytest tests/test_gpu_traversal.py -m gpu -x -q 2>&1 | tail -1
for cfg in "1 524288" "0 524288" "1 1048576" "1 262144" "0 262144" "1 524288" "0 524288"; do set -- $cfg
  ADYPT_HOST_RAMP=$1 ADYPT_HOST_CHUNK=$2 timeout 300 python tools/e2e_chunk_probe.py 2>&1 | tail -1 | sed "s/^/ramp $1 /"
done | tee gpurun_out/e2e_ramp.log
