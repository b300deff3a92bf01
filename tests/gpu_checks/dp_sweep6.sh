# usage: bash tests/gpu_checks/dp_sweep6.sh N -- grid shapes of the exchange with the bf16 gradient arena (final build)
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dp6_${N}_$name.err | tail -1 > gpurun_out/dp6_${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp6_${N}_$name.json")); dp=d.get("dp",{})
    k=d["roofline"]["kernel_ms_per_step"]
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "gemm_ms", round(sum(v for a,v in k.items() if a.startswith("gemm")),3), "ln_bwd", round(k.get("layernorm_bwd",0),3), flush=True)
except Exception as e:
    print("$name FAILED", e, flush=True)
PY
}
run default_32x512
run 148x128 --nvls-blocks 148 --nvls-threads 128
run 74x256 --nvls-blocks 74 --nvls-threads 256
run 16x1024_u8 --nvls-blocks 16 --nvls-threads 1024 --nvls-unroll 8
run nccl --dp-backend nccl
run bucket8 --bucket-mb 8
