for c in 0 128 0; do echo "chunk $c"; if [ $c = 0 ]; then unset NDPPGPU_FG_CHUNK; else export NDPPGPU_FG_CHUNK=$c; fi; VARIANTS="rt" python scripts/ab/fg_time.py | tail -2; done
