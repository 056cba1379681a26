#!/bin/bash
# build the .so first (the snapshot ships it), then run the given command on a B200 box
set -e
cd "$(dirname "$0")/.."
python mmoe-multimodal-rec_b200/build.py > /dev/null
python -c "import mmoe_multimodal_rec_b200 as p; p.lib()" 
exec /usr/local/graft/bin/gpurun "$@"
