#!/bin/bash
# Build an experimental variant of libb200vt.so with extra -D flags: tools/build_variant.sh <out.so> -DVT_FWD_EMU=3 ...
set -e
out=$1; shift
cd "$(dirname "$0")/../videotuna-dev_b200"
mkdir -p build/variant
objs=""
for f in csrc/*.cu; do
  o=build/variant/$(basename ${f%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out $objs -Xcompiler -fPIC -cudart static
echo built $out
