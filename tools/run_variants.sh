#!/bin/bash
# times BASELINE configs[1] (tests/prof_chain.py) and runs the pair-form parity tests for each library variant
# built by tools/build_variant.py:  tools/run_variants.sh name1 name2 ...   (on the GPU box)
cd "$(dirname "$0")/.."
for v in "$@"; do
  lib=$PWD/mosmicrosim_b200/variants/libmmsgpu_$v.so
  echo "=== $v"
  (cd tests && MMS_LIB=$lib python prof_chain.py 100000 2 fast 4 | tail -3)
  MMS_LIB=$lib python -m pytest -q -x tests/test_gpu_chain.py -k "pair_form or both_kernel_forms or two_habitats" 2>&1 | tail -2
done
