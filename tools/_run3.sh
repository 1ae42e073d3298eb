mkdir -p gpurun_out
: > gpurun_out/t3_seqg.log
for v in default seqg2 seqg1; do
  echo "== $v" >> gpurun_out/t3_seqg.log
  if [ $v = default ]; then unset ZSTDLITE_GPU_LIB; else export ZSTDLITE_GPU_LIB=$PWD/variants/$v.so; fi
  PROBE_ORDER=shuffled timeout 200 python tools/probe_decode.py 16384 65536 5 mix 2>&1 | grep -v "^iter [0-2]" >> gpurun_out/t3_seqg.log
done
cat gpurun_out/t3_seqg.log
