set -x
date +%s
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
date +%s
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 | cut -c1-400
date +%s
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default_err.log; cut -c1-300 gpurun_out/bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'], d['roofline']['frac'])"
date +%s
python -c "import __graft_entry__ as g; g.smoke()"
