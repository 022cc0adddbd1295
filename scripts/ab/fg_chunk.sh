#!/bin/bash
# Free gas (C3): the level-by-level walk against the chunked one (NDPPGPU_FG_CHUNK=128) on the library variant $V
# (default: a variant named "cur" built by `ab_build.py cur=`).  Usage on the box: V=cur bash scripts/ab/fg_chunk.sh
V=${V:-cur}
for c in 0 128 0; do
  echo "chunk $c"
  if [ $c = 0 ]; then unset NDPPGPU_FG_CHUNK; else export NDPPGPU_FG_CHUNK=$c; fi
  VARIANTS="$V" python scripts/ab/fg_time.py | tail -2
done
