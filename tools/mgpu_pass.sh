#!/bin/bash
# One gpurun --gpus G call: sharded-vs-single-GPU check, then bench.py under torchrun for a list of row-band settings.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/mgpu_pass.sh 2 tag "" "RDB_BAND_VCYCLE=4"'
G=${1:-2}
TAG=${2:-mgpu}
shift 2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
PORT=29511
run() {  # name, seconds, env string, script + args
  local name=$1 secs=$2 envs=$3
  shift 3
  echo "=== $name [$envs]" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  env $envs timeout "$secs" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$G" --master-addr 127.0.0.1 \
      --master-port $PORT "$@" >"$OUT/$name.log" 2>&1
  local rc=$?
  PORT=$((PORT + 1))
  echo "rc=$rc ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/summary.txt"
  grep -a -E '^\{|Error|error|Traceback' "$OUT/$name.log" | tail -n 6 | cut -c1-2500 | tee -a "$OUT/summary.txt"
}
if [ -z "$SKIP_CHECK" ]; then
  run check 400 "RDB_BAND_MULTIGRID=8" tools/mgpu_check.py
fi
if [ -z "$SKIP_TRACE" ]; then
  # TRACE_VARIANTS: ';'-separated configurations of rdb200_set_param switches, all timed inside one torchrun
  IFS=';' read -ra TV <<< "${TRACE_VARIANTS:-}"
  run band_profile 400 "" tools/band_profile.py 32768 "${TV[@]}"
  grep -a "trace\] rank 0\|^rep\|^== variant" "$OUT/band_profile.log" | tail -n ${TRACE_LINES:-70} | tee -a "$OUT/summary.txt"
fi
i=0
for e in "$@"; do
  run "bench_$i" 500 "$e" bench.py --gpus "$G" --steps ${STEPS:-4} --warmup 2 ${BENCH_FLAGS:---no-65536}
  i=$((i + 1))
done
echo done | tee -a "$OUT/summary.txt"
