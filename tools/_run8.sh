mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/t8_tests.log 2>&1
(time timeout 600 python bench.py --impl reference > gpurun_out/t8_bench_ref.json 2> gpurun_out/t8_bench_ref.err) 2> gpurun_out/t8_ref.time
(time timeout 600 python bench.py > gpurun_out/t8_bench.json 2> gpurun_out/t8_bench.err) 2> gpurun_out/t8_bench.time
cat gpurun_out/t8_tests.log gpurun_out/t8_ref.time gpurun_out/t8_bench.time; tail -2 gpurun_out/t8_bench.err; head -c 600 gpurun_out/t8_bench_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/t8_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['pageable']['value'], d['stages_ms'])
print({k:v for k,v in d['large_frame'].items() if 'ms' in k}, {k:v for k,v in d['dict'].items() if 'GBps' in k}, {k:v for k,v in d['config5'].items() if 'GBps' in k})
print({k:(v.get('GBps') if isinstance(v,dict) else v) for k,v in d['compress'].items()})
PY
