#!/usr/bin/env bash
# Builds nn-fac_b200/lib/libnnfac_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/csrc"; OBJ="$HERE/build"; LIB="$HERE/lib"
mkdir -p "$OBJ" "$LIB"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall)
pids=()
compile() {  # src obj extra-flags...
  local src="$1" obj="$2"; shift 2
  if [[ ! -f "$obj" || "$src" -nt "$obj" || -n "$(find "$SRC" "$HERE/../include" -name '*.cuh' -newer "$obj" -o -name '*.h' -newer "$obj" 2>/dev/null | head -1)" ]]; then
    "$NVCC" "${FLAGS[@]}" "$@" -c "$src" -o "$obj" &
    pids+=($!)
    if (( ${#pids[@]} >= ${JOBS:-8} )); then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
}
for f in "$SRC"/*.cu; do
  base="$(basename "$f" .cu)"
  [[ "$base" == "hals_sweep_inst" ]] && continue
  compile "$f" "$OBJ/$base.o"
done
for t in f32:float f64:double; do
  tag="${t%%:*}"; ty="${t##*:}"
  for rp in 16 32 64 128; do
    compile "$SRC/hals_sweep_inst.cu" "$OBJ/hals_sweep_${tag}_${rp}.o" -DSWEEP_T="$ty" -DSWEEP_TAG="$tag" -DSWEEP_RP="$rp"
  done
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -Wno-deprecated-gpu-targets -shared -o "$LIB/libnnfac_b200.so" "$OBJ"/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $LIB/libnnfac_b200.so"
