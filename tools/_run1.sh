mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_decode.py -q -x -m gpu -k "pageable or config2 or kat" 2>&1 | tail -5) > gpurun_out/t1_tests.log
for v in default plain; do
  if [ $v = plain ]; then export ZL_COPY_PLAIN=1; else unset ZL_COPY_PLAIN; fi
  echo "== $v" >> gpurun_out/t1_pageable.log
  timeout 200 python tools/probe_pageable.py >> gpurun_out/t1_pageable.log 2>&1
done
unset ZL_COPY_PLAIN
for t in 4 12 16; do echo "== threads $t" >> gpurun_out/t1_pageable.log; ZL_COPY_THREADS=$t timeout 200 python tools/probe_pageable.py 2>&1 | grep -v "^ZSTD_compress2\|pinned" >> gpurun_out/t1_pageable.log; done
ZL_DEC_TRACE=1 timeout 200 python tools/probe_dict.py > gpurun_out/t1_dict.log 2>&1
nproc >> gpurun_out/t1_pageable.log
cat gpurun_out/t1_tests.log gpurun_out/t1_pageable.log; tail -12 gpurun_out/t1_dict.log
