#!/bin/bash
# Build in-tree (the .so travels with the snapshot), check the library loads, then run a command on a B200 box.
#   tools/gpurun.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m introtocomputervision_b200.build > /dev/null
python -c "from introtocomputervision_b200 import _capi; _capi.lib()"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
