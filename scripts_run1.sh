timeout -s KILL 500 python -m pytest tests/test_gpu_edt.py -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err
python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(d['value'], d['e2e']['value'], d['roofline']['stage_ms'])"
