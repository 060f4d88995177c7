# usage: bash tools/gpu_sweep.sh <tag> "<name> ENV=..;<name> ENV=.."   (runs the GPU tests once, then one bench per config)
tag=${1:-x}; cfgs=${2:-"default X=1"}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
fi
IFS=';' read -ra arr <<< "$cfgs"
for c in "${arr[@]}"; do
  set -- $c; name=$1; shift
  env "$@" timeout -s KILL 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$name.log").read().strip().splitlines()[-1])
    r=d["roofline"]; g=d.get("roofline_gate_up") or r; dn=d["roofline_down"]
    print("$name tok/s %.1f ms %.3f e2e %.1f launches %d | top %.3f %.1fus | gu %.3f %.1fus dn %.3f %.1fus" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r["frac"], r["ms_per_launch"]*1e3, g["frac"], g["ms_per_launch"]*1e3, dn["frac"], dn["ms_per_launch"]*1e3))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${tag}_$name.log").read()[-1500:])
PY
done
