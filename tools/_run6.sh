mkdir -p gpurun_out
(time timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/fuzz_decode.py 2000 2>&1 | tail -6) > gpurun_out/t6_fuzz.log 2>&1
(time timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode.py -q -x -m gpu -k "built_on_the_device or pageable" 2>&1 | tail -6) > gpurun_out/t6_memcheck.log 2>&1
(time timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=zl_k_build_descs --error-exitcode 9 python -m pytest tests/test_gpu_decode.py -q -x -m gpu -k "built_on_the_device" 2>&1 | tail -6) > gpurun_out/t6_racecheck.log 2>&1
cat gpurun_out/t6_fuzz.log gpurun_out/t6_memcheck.log gpurun_out/t6_racecheck.log
