mkdir -p gpurun_out
(time timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/fuzz_decode.py 3000 2>&1 | tail -4) > gpurun_out/t12_fuzz.log 2>&1
(time timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/fuzz_decode.py 400 --large 2>&1 | tail -4) > gpurun_out/t12_fuzz_large.log 2>&1
(time timeout 250 python tools/fuzz_compress.py 300 2>&1 | tail -3) > gpurun_out/t12_fuzzc.log 2>&1
cat gpurun_out/t12_fuzz.log gpurun_out/t12_fuzz_large.log gpurun_out/t12_fuzzc.log
