#!/bin/bash
# Freeze the working tree into .gpu_snap/<tag>/ so that a queued gpurun call runs a consistent copy while development
# continues in the live tree (the driver snapshots /root/repo at an unknown moment after the call is queued).
#   tools/gpu_freeze.sh r2a   ->   gpurun -- 'cd .gpu_snap/r2a && ln -sfn $GRAFT_REPO_ROOT/gpurun_out gpurun_out && bash tools/gpu_r2.sh r2a'
set -e
TAG=${1:?tag}
cd "$(dirname "$0")/.."
rm -rf .gpu_snap/$TAG
mkdir -p .gpu_snap/$TAG
tar --exclude=./.git --exclude=./gpurun_out --exclude=./.gpu_snap --exclude='__pycache__' --exclude=./.pytest_cache \
    --exclude='*.o' --exclude='./oracle/_ref/obj' --exclude='./oracle/_ref/fast_gicp/obj' -cf - . | tar -xf - -C .gpu_snap/$TAG
touch .gpu_snap/$TAG
# keep only the two most recent snapshots (never the one just made)
ls -1dt .gpu_snap/*/ | grep -v "/$TAG/" | tail -n +2 | xargs -r rm -rf
du -sh .gpu_snap/$TAG
