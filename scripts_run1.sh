timeout -s KILL 900 python -m pytest tests/test_gpu_host_adapter.py tests/test_gpu_golden.py tests/test_gpu_search.py -q -m gpu 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err
python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['e2e']['value']); print({k:round(v) for k,v in d['extra'].items() if 'per_s' in k})"
