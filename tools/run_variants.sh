#!/bin/bash
# usage: tools/run_variants.sh "<VAR=val VAR=val>" ... ; each argument is one environment for tools/decode_curve.py 64 72,265
for spec in "$@"; do
  env TAG="[$spec]" $spec python tools/decode_curve.py 64 72,265 2>&1 | grep "B="
done
