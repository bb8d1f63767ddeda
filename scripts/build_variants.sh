#!/bin/bash
# Build A/B variants of libselas_cuda.so into selas_b200/lib_ab/<name>/ (objects in csrc/build_ab/<name>); pick one at run time with
# SELAS_CUDA_LIB=selas_b200/lib_ab/<name>/libselas_cuda.so.  usage: scripts/build_variants.sh name=-DFLAG[,-DFLAG...] ...
set -e
cd "$(dirname "$0")/.."
# ONLY=trace (or shade, ptex ...): start from the objects of the default build and recompile that translation unit alone
build() { name=$1; shift; mkdir -p selas_b200/lib_ab/$name; if [ -n "$ONLY" ]; then mkdir -p selas_b200/csrc/build_ab/$name; cp -p selas_b200/csrc/build/*.o selas_b200/csrc/build_ab/$name/; rm -f selas_b200/csrc/build_ab/$name/$ONLY.o selas_b200/lib_ab/$name/libselas_cuda.so; fi; make -s -C selas_b200/csrc -j8 OBJ=$PWD/selas_b200/csrc/build_ab/$name OUT=$PWD/selas_b200/lib_ab/$name EXTRA="$*" > /dev/null 2>&1 && echo "built $name ($*)"; }
for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    build "$name" ${flags//,/ }
done
