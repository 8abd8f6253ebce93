#!/bin/bash
# Tuning experiments on one kernel source: builds libtbknarpe variants that differ only in the -D flags of that file
# (trafficbotsv1.5_b200/_var/libtb_<name>.so; select one with TB_LIB=<path>).
#   profiles/kernel_variants.sh knarpe_attn_mma.cu name1 "-DFLAG=1 ..." name2 "..."
set -e
cd "$(dirname "$0")/.."
PKG=trafficbotsv1.5_b200
python -c "import __graft_entry__ as g; g.build()"
SRC=$1; shift
mkdir -p $PKG/_var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $flags -Xptxas=-v -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
    -c $PKG/csrc/$SRC -o $PKG/_var/var_$name.o 2>&1 | grep -E "registers|spill" | sort | uniq -c | sed "s/^/[$name] /"
  objs=$(ls $PKG/csrc/_obj/*.o | grep -v ${SRC%.cu}.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/_var/libtb_$name.so $objs $PKG/_var/var_$name.o -lcuda
done
