mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/t4_tests.log 2>&1
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/t4_smoke.log 2>&1
(time timeout 600 python bench.py > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err) 2> gpurun_out/t4_bench.time
cat gpurun_out/t4_tests.log gpurun_out/t4_smoke.log gpurun_out/t4_bench.time; tail -3 gpurun_out/t4_bench.err; head -c 1500 gpurun_out/t4_bench.json
