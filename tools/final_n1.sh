# round-2 final single-GPU validation + evidence (run under gpurun; writes gpurun_out/r02/final_*)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 400 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; tail -3 $O/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err; tail -c 300 $O/final_bench_n1.err
python - <<P
import json
d=json.loads(open("$O/final_bench_n1.json").read().strip().splitlines()[-1])
print("N1", d["steps"], d["warmup"], round(d["ms_per_step"],4), d["parity"]["ok"], round(d["roofline"]["frac"],3), d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, {k:(round(v["ms_per_step"],4), v["parity"]["ok"], round(v["roofline"]["frac"],3)) for k,v in d["workloads"].items()})
P
for w in mlp lenet mlp_bf16; do
  B200TF_KERNEL_TIMES=1 timeout 150 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-parity --workloads $w > /dev/null 2> $O/final_ktimes_$w.txt
  grep -A30 "kernel times" $O/final_ktimes_$w.txt | cut -c1-160 > $O/final_kernel_times_$w.txt; head -4 $O/final_kernel_times_$w.txt | cut -c1-120
done
for w in mlp lenet mlp_bf16; do
  B200TF_CUDA_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/final_launches_$w.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --workloads $w > $O/final_ncu_$w.log 2>&1
  wc -l $O/final_launches_$w.csv
done
