# usage: bash tests/gpu_checks/dp_sweep2.sh N -- exchange parameter sweep (bucket size, grid shape, cost of the fp32 pass)
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --graph-only "$@" 2> gpurun_out/dps${N}_$name.err | tail -1 > gpurun_out/dps${N}_$name.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dps${N}_$name.json")); dp=d.get("dp",{})
    print("N=$N $name", round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; no-exchange", round(dp.get("ms_per_step_without_allreduce",0),3))
except Exception as e:
    print("$name FAILED", e)
PY
}
run default
run bucket128 --bucket-mb 128
run bucket16 --bucket-mb 16
run ctas16 --nvls-blocks 16 --nvls-threads 512
run ctas64x256 --nvls-blocks 64 --nvls-threads 256
run noconvert --diag-dp-skip-convert
