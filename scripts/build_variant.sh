#!/bin/bash
# Development helper: builds homan_b200/_variants/<name>.so with extra nvcc flags for raster.cu, e.g.
#   scripts/build_variant.sh roleA -DHM_BWD_ROLE_MASK=1
set -e
name=$1; shift
mkdir -p homan_b200/_variants /tmp/hmvar_$name
C="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in api mano geom contact; do cp homan_b200/csrc/_obj/$f.o /tmp/hmvar_$name/; done
nvcc $C -fmad=false "$@" -c homan_b200/csrc/raster.cu -o /tmp/hmvar_$name/raster.o 2>&1 | grep -v deprecated || true
cp homan_b200/csrc/_obj/sdf.o /tmp/hmvar_$name/
nvcc -shared -o homan_b200/_variants/$name.so /tmp/hmvar_$name/*.o -lcudart 2>&1 | grep -v deprecated || true
echo homan_b200/_variants/$name.so
