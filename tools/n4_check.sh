# round-2 multi-GPU check (4 GPUs): default bench (NVLS overlap) for all workloads, exposed-NVLS comparison
mkdir -p gpurun_out/r02
N=${1:-4}
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 100 --warmup 20 --no-cpu-baseline $wl > gpurun_out/r02/n${N}_$tag.json 2> gpurun_out/r02/n${N}_$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/n${N}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), d["parity"]["ok"], d["config"].get("collective"), "e2e", round(d["e2e"]["value"]), {k:(round(v["ms_per_step"],4), v["parity"]["ok"], v["config"]["collective"]["per_step"]) for k,v in d.get("workloads",{}).items()})
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02/n${N}_$tag.err").read()[-800:])
P
}
run default "" X=1
run exposed "--workloads mlp" B200TF_COLLECTIVE_OVERLAP=0
