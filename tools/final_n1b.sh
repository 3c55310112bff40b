# round-2 final single-GPU validation (after the last kernel change): tests, smoke, default bench, LeNet evidence
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 400 python -m pytest tests -m gpu -x -q > $O/final2_pytest.log 2>&1; tail -2 $O/final2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $O/final2_bench_n1.json 2> $O/final2_bench_n1.err; tail -c 200 $O/final2_bench_n1.err
python - <<P
import json
d=json.loads(open("$O/final2_bench_n1.json").read().strip().splitlines()[-1])
print("N1", d["steps"], d["warmup"], round(d["ms_per_step"],4), d["parity"]["ok"], round(d["roofline"]["frac"],3), {k:(round(v["ms_per_step"],4), v["parity"]["ok"], round(v["roofline"]["frac"],3)) for k,v in d["workloads"].items()})
P
B200TF_KERNEL_TIMES=1 timeout 150 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-parity --workloads lenet > /dev/null 2> $O/final2_ktimes_lenet.txt
grep -A30 "kernel times" $O/final2_ktimes_lenet.txt | cut -c1-160 > $O/final2_kernel_times_lenet.txt
B200TF_CUDA_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/final2_launches_lenet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --workloads lenet > $O/final2_ncu_lenet.log 2>&1
wc -l $O/final2_launches_lenet.csv
