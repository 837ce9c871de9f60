#!/bin/bash
# tools/build_variant.sh NAME "-DRFB_X=.. -DRFB_Y=.." [file.cu ...]: an alternative build of librfb200.so (kernel-geometry sweeps).
# Recompiles the named .cu files (default k_fused_group.cu) with the extra defines, links them with the default build's other
# objects into rayforce_b200/librfb200_NAME.so; select it at run time with RFB200_LIB=...
set -e
name=$1; cfg=$2; shift 2
files=${@:-k_fused_group.cu}
here=$(cd "$(dirname "$0")/../rayforce_b200/csrc" && pwd)
make -s -C "$here" >/dev/null
tmp=$here/build/var_$name; mkdir -p "$tmp"
objs=""
for f in $here/build/*.o; do
  b=$(basename "$f" .o); skip=0
  for v in $files; do [ "$b.cu" = "$v" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $f"
done
for v in $files; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $cfg \
     -c "$here/$v" -o "$tmp/${v%.cu}.o" 2> "$tmp/${v%.cu}.ptxas.log" || { cat "$tmp/${v%.cu}.ptxas.log"; exit 1; }
  objs="$objs $tmp/${v%.cu}.o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$here/../librfb200_$name.so" $objs -cudart static
echo "built librfb200_$name.so"
