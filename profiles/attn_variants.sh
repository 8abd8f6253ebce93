#!/bin/bash
# Tuning experiments on the pair attention kernel: builds libtbknarpe variants that differ only in the -D flags of
# knarpe_attn_mma.cu (trafficbotsv1.5_b200/_var/libtb_<name>.so; select one with TB_LIB=<path>).
#   profiles/attn_variants.sh name1 "-DFLAG=1 ..." name2 "..."
set -e
cd "$(dirname "$0")/.."
PKG=trafficbotsv1.5_b200
python -c "import __graft_entry__ as g; g.build()"
mkdir -p $PKG/_var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc $flags -Xptxas=-v -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include \
    -c $PKG/csrc/knarpe_attn_mma.cu -o $PKG/_var/attn_$name.o 2>&1 | grep -E "registers|spill" | sort | uniq -c | sed "s/^/[$name] /"
  objs=$(ls $PKG/csrc/_obj/*.o | grep -v knarpe_attn_mma.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/_var/libtb_$name.so $objs $PKG/_var/attn_$name.o -lcuda
done
