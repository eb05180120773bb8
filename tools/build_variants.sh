#!/bin/bash
# Development aid: rebuild libosudit.so and build variants of csrc/attn_stream.cu with extra -D flags into
# tools/libosudit_<name>.so (loaded through OSUDIT_LIB).  usage: tools/build_variants.sh name1 "-DA=1 -DB=2" name2 "-DC=3" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/osu-diffusion_b200/csrc
make -C "$C" -j8 2>&1 | grep -v "^nvcc" | grep -v "^make" || true
grep -A2 "attn_stream_kernel" "$C/build/attn_stream.ptxas.log" | tail -2
mkdir -p "$C/build_var"
rm -f "$ROOT"/tools/libosudit_v_*.so
for f in "$C"/*.cu; do b=$(basename "${f%.cu}"); [ "$b" = attn_stream ] || cp "$C/build/$b.o" "$C/build_var/"; done
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
       -c "$C/attn_stream.cu" -o "$C/build_var/attn_stream.o" 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/tools/libosudit_v_$name.so" "$C"/build_var/*.o -lcudart
  echo "built v_$name ($flags)"
done
