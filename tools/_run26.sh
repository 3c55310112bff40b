N=$1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/allreduce_probe.py gpurun_out/allreduce_probe_n$N.json 2> gpurun_out/probe_err.log | tee gpurun_out/probe_n$N.log | grep -i "busbw\|NVLS\|channels\|Connected" | cut -c1-200 | tail -30
run() { tag=$1; shift
env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_err.log
python -c "
import json,sys
d=json.load(open('gpurun_out/bench_n${N}_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['roofline']['us_per_launch'],2), d['config']['host_enqueue_us_per_step'])" || tail -5 gpurun_out/bench_n${N}_err.log; }
run ov1 X=1
run ov0_one B200TF_COLLECTIVE_OVERLAP=0 B200TF_BUCKET_BYTES=none
