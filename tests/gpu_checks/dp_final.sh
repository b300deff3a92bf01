# usage: bash tests/gpu_checks/dp_final.sh N -- the default data-parallel bench line at N GPUs (no decode leg)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 20 --warmup 5 --no-decode 2> gpurun_out/r2y_bench_${N}gpu.err | tail -1 > gpurun_out/r2y_bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2y_bench_${N}gpu.json")); dp=d.get("dp",{})
print("N=$N", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms", d["launch_mode"][:30], "e2e", round(d["e2e"]["value"],1), "no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3), "parity", dp.get("parity_rel_err"), flush=True)
PY
