# usage: bash tests/gpu_checks/dp_sweep4.sh N -- the three candidate exchange configurations at N GPUs (graph replay)
N=${1:-8}
mkdir -p gpurun_out
export B200B_ATTN_TC=${B200B_ATTN_TC:-0}
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dp4_${N}_$name.err | tail -1 > gpurun_out/dp4_${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp4_${N}_$name.json")); dp=d.get("dp",{})
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "parity", dp.get("parity_rel_err",{}).get("eager"), dp.get("parity_rel_err",{}).get("graph_replay"), "gemm_ms", round(d["roofline"]["kernel_ms_per_step"].get("gemm_tcgen05_pair_kernel",0),3), flush=True)
except Exception as e:
    print("$name FAILED", e, flush=True)
PY
}
run default --dp-fp32-grads
run lazy
run lazy_nccl --dp-backend nccl
run lazy_64x256 --nvls-blocks 64 --nvls-threads 256
