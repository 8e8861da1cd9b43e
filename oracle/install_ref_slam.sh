#!/bin/bash
# oracle/install_ref_slam.sh — TEST INFRASTRUCTURE.  Installs the reference's SLAM scripts, UNMODIFIED, into
# oracle/_ref/gs_icp_slam/ so that tools/run_slam.py can run them where /root/reference does not exist (the GPU box):
#   gs_icp_slam.py gs_icp_slam_unlimit.py mp_Tracker.py mp_Tracker_unlimit.py mp_Mapper.py  scene/ utils/
#   gaussian_renderer/ arguments/ configs/
# Like the pip installs of oracle/build_ref_ext.sh this is an installation of the reference, not a copy into the
# repository: oracle/_ref/ is git-ignored (it travels with the gpurun snapshot only).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref/gs_icp_slam
if [ ! -f "$REF/mp_Mapper.py" ]; then
  echo "install_ref_slam: $REF not present, keeping $OUT (if any)"; exit 0
fi
rm -rf "$OUT"; mkdir -p "$OUT"
for f in gs_icp_slam.py gs_icp_slam_unlimit.py mp_Tracker.py mp_Tracker_unlimit.py mp_Mapper.py; do cp "$REF/$f" "$OUT/"; done
for d in scene utils gaussian_renderer arguments configs; do cp -r "$REF/$d" "$OUT/$d"; done
find "$OUT" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "install_ref_slam: $(find "$OUT" -name '*.py' | wc -l) python files -> $OUT"
