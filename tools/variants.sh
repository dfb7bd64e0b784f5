#!/bin/bash
# Builds experiment variants of libnflgpu into build/variants/<name>/libnflgpu.so (only log2(N)=ONLY instantiated).
# usage: tools/variants.sh ONLY name1 "flags1" name2 "flags2" ...
set -e
ONLY=$1; shift
cd "$(dirname "$0")/../nfllib_b200/csrc"
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p ../../build/variants/$name
  make -s -j8 OUT=../../build/variants/$name/libnflgpu.so BUILD=../../build/variants/$name/obj EXTRA="-DNFLGPU_ONLY_LOGN=$ONLY $flags" 2>&1 | grep -E "error|warning: v|spill" || true
  echo "built $name"
done
