#!/bin/bash
# One gpurun call: GPU parity suite, fill variants with the V-cycle timeline, bench, ncu launch list, reference arm.
#   gpurun --timeout 1500 -- 'bash tools/gpu_pass.sh 32768 tag'
# Every step runs under its own `timeout` and writes to gpurun_out/<tag>/.
N=${1:-32768}
TAG=${2:-pass}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
cd "$(dirname "$0")/.."

step() {  # name, seconds, command...
  local name=$1 secs=$2
  shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$secs" "$@" >"$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc ($(( $(date +%s) - t0 )) s)" | tee -a "$OUT/summary.txt"
  tail -n "${TAILN:-12}" "$OUT/$name.log" | cut -c1-1500 | tee -a "$OUT/summary.txt"
}

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >"$OUT/gpu.txt" 2>&1
if [ -z "$SKIP_TESTS" ]; then
  step pytest_gpu 900 python -m pytest tests -m gpu -x -q
fi
if [ -z "$LITE" ]; then
TAILN=40 step fill_trace 300 python tools/fill_profile.py "$N" "fill_trace=1"
step fill_variants 400 python tools/fill_profile.py "$N" "" "fill_vcycle=4" "fill_vcycle=12"
TAILN=60 step band_profile 300 python tools/band_profile.py "$N"
step flats_profile 240 env RDB200_PROFILE=1 python tools/flats_profile.py "$N"
fi
TAILN=40 step dinf_engines 300 python tools/dinf_profile.py "$N" "${DINF_CONFIGS:-accum_dinf_packed=0;accum_dinf_packed=1 accum_dinf_stats=1}"
step d8_scan 200 python tools/dinf_profile.py "$N" d8 "${D8_CONFIGS:-;accum_walk_scan=0}"
TAILN=3 step bench 600 python bench.py --steps 5 --warmup 3
grep -a '"metric"' "$OUT/bench.log" | tail -1 > "$OUT/bench.json"
[ -z "$LITE" ] && step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-verify
if [ -n "$NCU_DINF" ]; then
  step ncu_dinf 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flow_code|dinf|deps_gather" -c 60 --csv \
    --log-file "$OUT/launches_dinf.csv" python tools/dinf_profile.py "$N" "accum_dinf_packed=1"
fi
if [ -n "$NCU_FULL" ]; then
  # full-set captures of the dominant kernels (one replayed launch each; read here with `ncu -i ... --page raw --csv`):
  # the heaviest sweep launch is the first round of the 32768^2 level (after the coarse levels' ~80 small launches)
  step ncu_full_sweep 600 ncu --set full --clock-control none --import-source on -k regex:fill_sweep_kernel --launch-skip 78 --launch-count 8 \
    -o "$OUT/r2_fill_sweep_full" -f python tools/fill_profile.py "$N" ""
  step ncu_full_fa 600 ncu --set full --clock-control none --import-source on -k regex:"fa_d8_prep_rolling|accum_walk_packed_lanes" --launch-count 2 \
    -o "$OUT/r2_fa_d8_full" -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-configs --no-verify
fi
if [ -n "$REF_SIZE" ]; then
  TAILN=3 step bench_reference 900 python bench.py --impl reference --steps 1 --warmup 0 --ref-size "$REF_SIZE"
  grep -a '"impl"' "$OUT/bench_reference.log" | tail -1 > "$OUT/bench_reference.json"
fi
echo "done" | tee -a "$OUT/summary.txt"
