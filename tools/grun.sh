#!/bin/bash
# Rebuild the in-tree library (stamp-checked), then run a command on the GPU box.
# usage: tools/grun.sh <timeout-seconds> '<command>'
set -e
cd /root/repo
python -m vmc_pde_b200.build >/dev/null
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
