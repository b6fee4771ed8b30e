#!/bin/sh
# Stage the UNMODIFIED reference scripts where the GPU box can see them: baseline/_ref/ (git-ignored, not
# gpurun-ignored -- the location the build contract names for the unmodified reference).  /root/reference does not
# exist on the GPU box, and reference sources are never committed; tests/test_reference_scripts.py looks for
# $SYMPGPR_REFERENCE, then baseline/_ref/python, then /root/reference/python, and skips when none is there.
# The vendored DVODE sources (0.7 MB, training-data generator of script 03) are left out.
set -e
cd "$(dirname "$0")/.."
SRC=${1:-/root/reference/python}
rm -rf baseline/_ref/python
mkdir -p baseline/_ref
cp -r "$SRC" baseline/_ref/python
rm -rf baseline/_ref/python/03_henon_heiles/vode
find baseline/_ref -name "__pycache__" -type d -prune -exec rm -rf {} +
du -sh baseline/_ref
