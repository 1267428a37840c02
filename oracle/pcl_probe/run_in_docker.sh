#!/usr/bin/env bash
# Pins the oracle against the REAL reference on a machine that has docker and network access (this repository's build
# image has neither ROS nor PCL):
#   1. pcl_probe: one PASS/FAIL line per third-party assumption A1..A18 (PCL only);
#   2. pcl_replay: the unmodified reference class over the seeded C1 fixture -> tests/golden/pcl_c1.json, which
#      tests/test_golden.py::test_*_reproduces_pcl_golden then hold the oracle and the CUDA path to.
# usage: oracle/pcl_probe/run_in_docker.sh <path to a checkout of prabinrath/dynamicslamtool> [frames=12]
set -euo pipefail
REF=$(realpath "${1:?path to the reference checkout}")
FRAMES=${2:-12}
REPO=$(realpath "$(dirname "$0")/../..")
docker run --rm -v "$REPO":/repo -v "$REF":/ref:ro ros:melodic-perception bash -ec '
  source /opt/ros/melodic/setup.bash
  mkdir -p /tmp/probe && cd /tmp/probe && cmake /repo/oracle/pcl_probe >/dev/null && make -j"$(nproc)" >/dev/null
  ./pcl_probe | tee /repo/oracle/pcl_probe/probe_result.txt || true
  mkdir -p /tmp/replay && cd /tmp/replay && cmake /repo/oracle/pcl_probe/replay -DMOR_REFERENCE_DIR=/ref >/dev/null && make -j"$(nproc)" >/dev/null
  roscore >/dev/null 2>&1 & sleep 4
  ./pcl_replay /repo/config/MOR_config.txt /repo/tests/golden/pcl_c1.json '"$FRAMES"' 1 1
  kill %1 || true
'
echo "wrote $REPO/tests/golden/pcl_c1.json and $REPO/oracle/pcl_probe/probe_result.txt; now run: python -m pytest tests/test_golden.py -q"
