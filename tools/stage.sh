#!/bin/bash
# Freeze a copy of the working tree under .stage/<name> so that a queued gpurun call runs a consistent
# snapshot while the main tree keeps changing.  usage: tools/stage.sh <name>
set -e
cd "$(dirname "$0")/.."
name=$1
rm -rf .stage/$name
mkdir -p .stage/$name
tar --exclude=./.git --exclude=./gpurun_out --exclude=./.stage --exclude=./.pytest_cache --exclude='__pycache__' \
    --exclude=./tools/exp/madx --exclude=./tools/exp/mulvar --exclude=./tools/exp/pipes --exclude=./tools/exp/prodvar --exclude=./tools/exp/rr \
    -cf - . | tar -xf - -C .stage/$name
echo staged .stage/$name $(du -sh .stage/$name | cut -f1)
