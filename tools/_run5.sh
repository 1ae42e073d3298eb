mkdir -p gpurun_out
K='regex:zl_k_execute|zl_k_literals|zl_k_sequences|zl_k_index'
timeout 400 ncu --set full --import-source on --clock-control none -k "$K" -s 8 -c 4 -f -o gpurun_out/r02c_decode python tools/prof_exec.py 16384 65536 mix 4 > gpurun_out/t5_ncu_decode.log 2>&1
ncu -i gpurun_out/r02c_decode.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_full_decode_kernels_raw.csv 2>> gpurun_out/t5_ncu_decode.log
ls -la gpurun_out/r02c_decode.ncu-rep >> gpurun_out/t5_ncu_decode.log
if [ $(stat -c %s gpurun_out/r02c_decode.ncu-rep) -gt 30000000 ]; then rm gpurun_out/r02c_decode.ncu-rep; fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_16384x64k.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-dict --no-large --no-compress --no-config5 > gpurun_out/t5_launches.log 2>&1
timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none -k regex:zl_k_build_descs -c 2 --csv --page raw --log-file gpurun_out/r02c_ncu_build_descs_raw.csv python tools/prof_other.py > gpurun_out/t5_other.log 2>&1
bash tools/ab_bench.sh t5 default > gpurun_out/t5_ab.log 2>&1
ZL_SEQ_CTAS_PER_SM=4 bash tools/ab_bench.sh t5c4 default >> gpurun_out/t5_ab.log 2>&1
ZL_LIT_CTAS_PER_SM=4 bash tools/ab_bench.sh t5l4 default >> gpurun_out/t5_ab.log 2>&1
tail -3 gpurun_out/t5_ncu_decode.log; tail -2 gpurun_out/t5_launches.log; tail -2 gpurun_out/t5_other.log; cat gpurun_out/t5_ab.log
