mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_decode.py tests/test_gpu_compress.py -q -x -m gpu -k "pageable or built_on_the_device or many_small or scattered or dictionary" 2>&1 | tail -8) > gpurun_out/t2_tests.log
echo "== devbuild" > gpurun_out/t2_dict.log
ZL_DEC_TRACE=1 timeout 200 python tools/probe_dict.py 2>&1 | grep -v "^slice" >> gpurun_out/t2_dict.log
echo "== host descriptors" >> gpurun_out/t2_dict.log
ZL_DEC_NODEVBUILD=1 ZL_DEC_TRACE=1 timeout 200 python tools/probe_dict.py 2>&1 | grep -v "^slice" >> gpurun_out/t2_dict.log
timeout 200 python tools/probe_pageable.py > gpurun_out/t2_pageable.log 2>&1
cat gpurun_out/t2_tests.log gpurun_out/t2_dict.log gpurun_out/t2_pageable.log
