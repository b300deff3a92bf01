# usage: bash tests/gpu_checks/dp_sweep3.sh N -- round-2 exchange sweep: units in flight per thread, grid shape,
# exclusive SMs, and the averaged gradients left in the bf16 arena (no fp32 pass)
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dp3_${N}_$name.err | tail -1 > gpurun_out/dp3_${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dp3_${N}_$name.json")); dp=d.get("dp",{})
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "parity", dp.get("parity_rel_err",{}).get("eager"), dp.get("parity_rel_err",{}).get("graph_replay"), "gemm_ms", round(d["roofline"]["kernel_ms_per_step"].get("gemm_tcgen05_pair_kernel",0),3), flush=True)
except Exception as e:
    print("$name FAILED", e, flush=True)
PY
}
run default --dp-fp32-grads
run u16 --nvls-unroll 16
run lazy
run lazy_u16 --nvls-unroll 16
run lazy_16x1024_u16 --nvls-unroll 16 --nvls-blocks 16 --nvls-threads 1024
run lazy_8x1024_u16 --nvls-unroll 16 --nvls-blocks 8 --nvls-threads 1024
run lazy_excl4_u16 --nvls-unroll 16 --nvls-blocks 4 --nvls-threads 1024 --nvls-exclusive
run lazy_148x512_u8_b128 --nvls-unroll 8 --nvls-blocks 148 --nvls-threads 512 --bucket-mb 128
run lazy_nccl --dp-backend nccl
