#!/usr/bin/env bash
# step kernel table (CUPTI), then the ncu launch list of one training step.  gpurun --timeout 1500 -- bash tools/gpu_profile.sh
set -u
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/step_profile.txt 2> gpurun_out/step_profile.err
echo "profile_step rc=$?"; head -45 gpurun_out/step_profile.txt
timeout 300 python tools/profile_step.py --ncu > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py --ncu > gpurun_out/ncu.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/launches.csv
