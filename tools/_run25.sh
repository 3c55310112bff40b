timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
run() { tag=$1; shift
env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_err.log
python -c "
import json,sys
d=json.load(open('gpurun_out/bench_n2_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['roofline']['us_per_launch'],2), d['config']['host_enqueue_us_per_step'])" || tail -5 gpurun_out/bench_n2_err.log; }
run ov1 X=1
run ov0 B200TF_COLLECTIVE_OVERLAP=0
run ov0_one B200TF_COLLECTIVE_OVERLAP=0 B200TF_BUCKET_BYTES=none
run ov1_one B200TF_BUCKET_BYTES=none
run ov1_cta8 NCCL_MAX_CTAS=8
run ov1_cta4 NCCL_MAX_CTAS=4
run ov1_noarena B200TF_GRADIENT_ARENA=0 B200TF_ALLREDUCE_PACK=0
