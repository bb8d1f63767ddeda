#!/bin/bash
# Build A/B variants of libselas_cuda.so into selas_b200/lib_ab/<name>/ (objects in csrc/build_ab/<name>)
set -e
cd "$(dirname "$0")/.."
build() { name=$1; shift; mkdir -p selas_b200/lib_ab/$name; make -s -C selas_b200/csrc -j8 OBJ=$PWD/selas_b200/csrc/build_ab/$name OUT=$PWD/selas_b200/lib_ab/$name EXTRA="$*" > /dev/null 2>&1 && echo "built $name ($*)"; }
build refill16 -DSELAS_REFILL_THRESHOLD=16
build refill24 -DSELAS_REFILL_THRESHOLD=24
build refill28 -DSELAS_REFILL_THRESHOLD=28
build stack24 -DSELAS_STACK_SIZE=24
build t64x16 -DSELAS_TRACE_THREADS=64 -DSELAS_TRACE_CTAS_PER_SM=16
build ctas12 -DSELAS_TRACE_CTAS_PER_SM=12
build ctas6 -DSELAS_TRACE_CTAS_PER_SM=6
