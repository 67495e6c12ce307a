#!/bin/bash
# One gpurun call that validates and times every switch round 1 left off by default.
# Every step runs under its own `timeout` and writes to gpurun_out/first_pass/, so a hang or crash in one
# experimental kernel costs that step only.
#
#   gpurun --timeout 1500 -- 'bash tools/first_pass.sh 32768'
#
# Order: cheap parity first (a switch that fails parity is not timed), then timings at N (default 16384).
N=${1:-16384}
OUT=gpurun_out/first_pass
mkdir -p "$OUT"
cd "$(dirname "$0")/.."

step() {  # name, seconds, command...
  local name=$1 secs=$2
  shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  timeout "$secs" "$@" >"$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a "$OUT/summary.txt"
  tail -n 12 "$OUT/$name.log" | tee -a "$OUT/summary.txt"
}

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >"$OUT/gpu.txt" 2>&1

# 1. parity of each switch against the reference on small rasters (opt-in test file)
for k in fill_multigrid accum_fused_prep accum_walk_lanes accum_agg accum_tail accum_async flats_uf_tiled flowdirs_rolling fill_async eight_receiver; do
  RDB_TEST_EXPERIMENTAL=1 step "parity_$k" 240 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q -k "$k"
done

# 2. timings
step time_fa_d8 240 python tools/accum_switches.py "$N"
step time_fa_dinf 400 python tools/accum_switches.py "$N" --dinf
step time_flats_base 240 env RDB200_PROFILE=1 python tools/flats_profile.py "$N"
step time_flats_uf_tiled 240 env RDB200_PROFILE=1 python tools/flats_profile.py "$N" flats_uf_tiled=1
step time_fill 400 python tools/fill_profile.py "$N" "" "fill_ordered=0" "fill_multigrid=8" "fill_multigrid=4" "fill_multigrid=8,fill_vcycle=4" "fill_multigrid=8,fill_vcycle=8" "fill_multigrid=4,fill_vcycle=4" "fill_multigrid=8,fill_async=1" "fill_multigrid=8,fill_vcycle=4,fill_async=1" "fill_async=1"
step time_fill_async_unordered 200 python tools/fill_profile.py "$N" "fill_async=1,fill_ordered=0"

# 3. the whole GPU suite and the bench under the most promising combination
CAND="fill_multigrid=8,accum_fused_prep=1,accum_walk_lanes=1"
step suite_under_candidates 400 env RDB200_PARAMS="$CAND" python -m pytest tests -m gpu -x -q
step bench_under_candidates 300 env RDB200_PARAMS="$CAND" python bench.py --steps 3 --warmup 3

echo "done" | tee -a "$OUT/summary.txt"
