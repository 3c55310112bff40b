mkdir -p gpurun_out/r02
run() { tag=$1; shift; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 100 --warmup 20 --no-cpu-baseline --workloads mlp > gpurun_out/r02/ov_$tag.json 2> gpurun_out/r02/ov_$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/ov_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), d["parity"]["ok"], d["config"].get("collective"), d.get("cuda_graph"))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02/ov_$tag.err").read()[-600:])
P
}
run base X=1
run ov_nograph B200TF_COLLECTIVE_OVERLAP=1 B200TF_BUCKET_BYTES=4000000 B200TF_PEER_CTAS=16 B200TF_CUDA_GRAPH=0
run ov_graph16 B200TF_COLLECTIVE_OVERLAP=1 B200TF_BUCKET_BYTES=4000000 B200TF_PEER_CTAS=16
run ov_graph32 B200TF_COLLECTIVE_OVERLAP=1 B200TF_BUCKET_BYTES=4000000 B200TF_PEER_CTAS=32
run ov_graph16_b2m B200TF_COLLECTIVE_OVERLAP=1 B200TF_BUCKET_BYTES=2000000 B200TF_PEER_CTAS=16
