#!/bin/bash
# ncu evidence for the non-step kernels: launch list + one full capture each of depth camera, ray cast, SDF, clone.
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python scripts/secondary_driver.py > $OUT/${TAG}_secondary.json 2> $OUT/${TAG}_secondary.err; echo "driver rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:elg_ -c 400 --csv --log-file $OUT/${TAG}_secondary_launches.csv \
    python scripts/secondary_driver.py > /dev/null 2>&1
for k in ${KERNELS:-elg_depth_camera_kernel elg_raycast_kernel elg_sdf_kernel elg_clone_bulk_kernel elg_actuator_kernel elg_norm_cols_kernel elg_reset_kernel elg_torques4_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/${TAG}_$k python scripts/secondary_driver.py > /dev/null 2>&1
  echo "$k rc=$?"
done
ls -la $OUT | grep ${TAG}_
