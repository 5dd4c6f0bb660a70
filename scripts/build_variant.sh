#!/bin/bash
# Development aid: build a library variant out of tree.  usage: build_variant.sh <name> "<extra nvcc flags>"
# -> build_variants/libdcol_<name>.so (+ ptxas logs under /tmp/bv_<name>)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N=$1; shift
D=/tmp/bv_$N
rm -rf $D && mkdir -p $D/pkg/csrc $D/include
cp $ROOT/differentiablecollisions.jl_b200/csrc/*.cu $ROOT/differentiablecollisions.jl_b200/csrc/*.cuh $ROOT/differentiablecollisions.jl_b200/csrc/Makefile $D/pkg/csrc/
cp $ROOT/include/*.h $D/include/
mkdir -p $ROOT/build_variants
make -C $D/pkg/csrc -j3 EXTRA="$*" OUT=$ROOT/build_variants/libdcol_$N.so > $D/make.log 2>&1 || { tail -20 $D/make.log; exit 1; }
grep -h "registers" $D/pkg/csrc/dcol_kernels_v0.ptxas.log | head -1
grep -h "spill" $D/pkg/csrc/dcol_kernels_v0.ptxas.log | awk '{s+=$5; l+=$9} END {print "v0 spill stores", s, "loads", l}'
