# usage: bash tests/gpu_checks/dp_sweep.sh N  -- bench.py at N GPUs for each gradient-exchange transport
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 30 --warmup 5 --no-decode "$@" 2> gpurun_out/dp${N}_$name.err | tail -1 > gpurun_out/dp${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp${N}_$name.json")); dp=d.get("dp",{})
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-allreduce", round(dp.get("ms_per_step_without_allreduce",0),3), dp.get("allreduce","")[:100])
except Exception as e:
    print("$name FAILED", e)
PY
}
run nccl --dp-backend nccl
run nvls_shared32_inplace --dp-backend nvls --nvls-shared --nvls-blocks 32 --nvls-threads 512 --nvls-inplace
run nvls_shared32_mc32 --dp-backend nvls --nvls-shared --nvls-blocks 32 --nvls-threads 512
