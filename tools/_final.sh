set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 300 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_err.log; tail -c 600 gpurun_out/bench_n1_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1_err.log; tail -c 400 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 24 -c 8 -o gpurun_out/r01_gemm_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r01_gemm_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,launch__registers_per_thread,launch__grid_size,launch__cluster_size > gpurun_out/r01_gemm_full_summary.csv 2>&1
timeout 300 python tools/op_bench.py --json gpurun_out/op_rooflines.json > gpurun_out/op_bench.log 2>&1; tail -32 gpurun_out/op_bench.log
