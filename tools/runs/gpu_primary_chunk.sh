#!/bin/bash
# class-sorted bounce 0: samples per queue-slot request (ADYPT_PRIMARY_CHUNK, default 4) and samples per item (ADYPT_PRIMARY_GROUP, default 16)
for cfg in "4 16" "8 16" "2 16" "4 8"; do set -- $cfg
  echo "== chunk $1 group $2"; ADYPT_PRIMARY_CHUNK=$1 ADYPT_PRIMARY_GROUP=$2 REPS=3 timeout 200 python tools/pt_time.py 2>&1 | grep -E "stage|C3"
done | tee gpurun_out/primary_sorted_chunk.log
