# usage: bash tests/gpu_checks/dp_sweep5.sh N -- is it the exchange or the symmetric-memory arenas that slows the backward GEMMs?
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dp5_${N}_$name.err | tail -1 > gpurun_out/dp5_${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp5_${N}_$name.json")); dp=d.get("dp",{})
    k=d["roofline"]["kernel_ms_per_step"]
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "parity", dp.get("parity_rel_err",{}).get("eager"), "gemm_ms", round(k.get("gemm_tcgen05_pair_kernel",0),3), {a:round(b,3) for a,b in k.items() if not a.startswith("gemm")}, flush=True)
except Exception as e:
    print("$name FAILED", e, flush=True)
PY
}
run lazy
run lazy_skip_exchange --diag-dp-skip-exchange
run lazy_nccl_skip_exchange --diag-dp-skip-exchange --dp-backend nccl
run lazy_b256 --bucket-mb 256
