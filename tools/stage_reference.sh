#!/bin/bash
# Carry the UNMODIFIED reference checkout to the GPU box for tests/test_gpu_dropin.py: baseline/_ref/ is
# git-ignored (the reference's sources never enter this repository's history) but travels with the gpurun
# snapshot.  Run in the build container, where /root/reference exists.
set -e
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
rm -rf baseline/_ref/reference
cp -r "${PPNP_REFERENCE:-/root/reference}" baseline/_ref/reference
rm -rf baseline/_ref/reference/.git
echo "staged $(find baseline/_ref/reference -type f | wc -l) files under baseline/_ref/reference"
