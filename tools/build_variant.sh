#!/bin/bash
# tools/build_variant.sh NAME "<extra nvcc flags>": builds audio_tabs_b200/lib/variants/NAME.so from the
# current sources (tuning experiments; select at run time with B200SPEC_LIB=<path>).
set -e
NAME=$1; shift
cd "$(dirname "$0")/.."
mkdir -p audio_tabs_b200/lib/variants
B200SPEC_EXTRA_NVCC_FLAGS="$*" python -m audio_tabs_b200.build --force > /dev/null
cp audio_tabs_b200/lib/libb200spec.so audio_tabs_b200/lib/variants/$NAME.so
python -m audio_tabs_b200.build --force > /dev/null   # leave the default build in place
echo "built variant $NAME ($*)"
