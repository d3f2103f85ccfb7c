#!/bin/bash
# A/B of two library builds under profiles/stress_tensor_subleq.py on ONE box (the race of DESIGN.md section 7 (4) showed on some boxes only).
# Usage (on a B200): bash profiles/ab_race.sh /path/to/old/libeaz_b200.so   -- the current build is the other arm
OLD=${1:?path of the library build to compare against}
nvidia-smi --query-gpu=serial --format=csv,noheader
for i in 1 2; do
  echo OLD; EAZ_LIB_PATH=$OLD timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
  echo NEW; timeout 200 python profiles/stress_tensor_subleq.py 250 2>&1 | tail -1
done
