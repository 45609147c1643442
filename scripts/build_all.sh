#!/bin/bash
# product library, the A/B variants used by scripts/gpu_call_*.sh, the C++ host drivers and the oracle
set -e
cd "$(dirname "$0")/.."
python le-vins_b200/build.py > /dev/null
(cd le-vins_b200 && for v in "$@"; do name=${v%%:*}; defs=${v#*:}; python build_variant.py $name $defs > /dev/null; done)
python le-vins_b200/host/build_host.py > /dev/null
python -c "import sys; sys.path.insert(0,'le-vins_b200'); from levins import synthgen; synthgen.build()"
python -c "import sys; sys.path.insert(0,'oracle'); import oracle_py; oracle_py.build()"
echo built
