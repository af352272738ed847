timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --no-cpu --no-stream > gpurun_out/bench_r03g.json 2> gpurun_out/bench_r03g.err; echo rc=$?
python -c "
import json; j=json.load(open('gpurun_out/bench_r03g.json')); print(round(j['value']), j['ms_per_step'], j['e2e'])"
