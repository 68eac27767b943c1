#!/bin/bash
# Stage the reference's xfuser/compact python files under baseline/_ref (git-ignored, travels to the
# GPU box with gpurun) so tools/ref_gpu_compare.py can run the UNMODIFIED reference on the B200.
# Build-container only: /root/reference does not exist on the GPU box.  Nothing is committed.
set -e
REF=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref"
mkdir -p "$DST/xfuser/collector" "$DST/xfuser/compact"
cp "$REF/xfuser/prof.py" "$DST/xfuser/"
cp "$REF"/xfuser/compact/*.py "$DST/xfuser/compact/"
cp -r "$REF/xfuser/compact/patchpara" "$DST/xfuser/compact/"
cp "$REF"/xfuser/collector/*.py "$DST/xfuser/collector/"
echo "staged reference files under $DST"
# the reference's own unit tests of the plugin (run against THIS library by tools/run_reference_tests.py)
mkdir -p "$DST/tests/compact"
cp "$REF"/tests/compact/*.py "$DST/tests/compact/"
# regular packages, so that `from tests.compact... import` finds THIS tree and not some installed `tests`
touch "$DST/tests/__init__.py" "$DST/tests/compact/__init__.py"
echo "staged the reference's tests/compact under $DST/tests/compact"
