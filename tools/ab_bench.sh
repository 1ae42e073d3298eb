#!/bin/bash
# Development: decode-only bench line of several library variants on one box.  usage: tools/ab_bench.sh tag variant...   ("default" = the in-tree library)
tag=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset ZSTDLITE_GPU_LIB; else export ZSTDLITE_GPU_LIB=variants/$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu --no-dict --no-large --no-compress --no-config5 > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(f"{sys.argv[1]:16s} value {d['value']:.1f} GB/s  e2e {d['e2e']['value']:.1f}  stages {' '.join(f'{k}={v:.3f}' for k, v in d['stages_ms'].items())}", flush=True)
PY
done
