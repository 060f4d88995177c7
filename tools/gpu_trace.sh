# usage: bash tools/gpu_trace.sh "<name> ENV=..;<name> ENV=.."   -> phase traces of the block kernel per config
mkdir -p gpurun_out
IFS=';' read -ra arr <<< "$1"
for c in "${arr[@]}"; do
  set -- $c; name=$1; shift
  echo "== $name $@"
  env "$@" timeout -s KILL 300 python profiles/block_trace.py 2>&1 | tee gpurun_out/block_trace_$name.txt | tail -11
done
