mkdir -p gpurun_out
bash tools/ab_bench.sh t11g6 default > gpurun_out/t11_ab.log 2>&1
ZL_DEC_GRADE=7 bash tools/ab_bench.sh t11g7 default >> gpurun_out/t11_ab.log 2>&1
ZL_DEC_GRADE=8 bash tools/ab_bench.sh t11g8 default >> gpurun_out/t11_ab.log 2>&1
cat gpurun_out/t11_ab.log
