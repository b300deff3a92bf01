# usage: bash tests/gpu_checks/dp_sweep7.sh N -- candidates for the default exchange configuration at N GPUs (final build)
N=${1:-4}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dp7_${N}_$name.err | tail -1 > gpurun_out/dp7_${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp7_${N}_$name.json")); dp=d.get("dp",{})
    k=d["roofline"]["kernel_ms_per_step"]
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "gemm_ms", round(sum(v for a,v in k.items() if a.startswith("gemm")),3), "parity", dp.get("parity_rel_err",{}).get("graph_replay"), flush=True)
except Exception as e:
    print("$name FAILED", e, flush=True)
PY
}
run nvls_32x512 --dp-backend nvls
run nvls_16x1024_u8 --dp-backend nvls --nvls-blocks 16 --nvls-threads 1024 --nvls-unroll 8
if [ "$N" -le 4 ]; then run nccl --dp-backend nccl; fi
