#!/bin/bash
# same-box A/B of the lean step kernel: product build, the optional micro-optimisations (scripts/ab/libelg_{clip,sums,both}.so), round-1 library
for lib in "" scripts/ab/libelg_clip.so scripts/ab/libelg_sums.so scripts/ab/libelg_both.so scripts/ab/libelg_v5.so; do
  echo "--- ${lib:-product}"
  ELG_LIB_PATH=$lib timeout 300 python scripts/step_ab.py quick 2>&1 | grep "^fast"
done
