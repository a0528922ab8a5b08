#!/bin/bash
# compute-sanitizer over the code paths added in round 2 (small tests only: the tools slow kernels down 10-100x)
TAG=${1:-r2i}
OUT=gpurun_out/${TAG}_sanitizer.txt
mkdir -p gpurun_out
if [ "$2" != "compact" ]; then
echo "compute-sanitizer on a B200 (gpurun), library built from the current tree (round 2: spatial aliasing decode, re-draws, batched options, terminal observations, state-based sampler):" > $OUT
echo "== memcheck: batched options (same setup / other side / terminal observations), device resets, spatial aliasing (micro, tiny)" >> $OUT
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_batched_options_gpu.py tests/test_device_reset_gpu.py tests/test_spatial_alias_gpu.py -x -q \
   -k "micro or tiny or toy or overflow or random_player" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 >> $OUT
echo "== memcheck: state-based policy sampler (micro, octa)" >> $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_policy_sampling_gpu.py -x -q -k "state_based and (micro or octa) and dtype0" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid" | head >> $OUT
echo "== racecheck: state-based policy sampler (shared-memory running counts) and terminal observations (micro)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_policy_sampling_gpu.py tests/test_batched_options_gpu.py -x -q -k "(state_based and micro and dtype0) or (terminal and micro)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard" | head >> $OUT
echo "== racecheck: smoke (10x10 fused kernel)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python __graft_entry__.py smoke 2>&1 | grep -E "COMPUTE-SANITIZER|smoke ok|RACECHECK SUMMARY|hazard" | head >> $OUT
echo "== synccheck: repeat-from-other-side reset + terminal observations (short_barrage)" >> $OUT
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_batched_options_gpu.py -x -q -k "short_barrage and (other_side or terminal)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY" | head >> $OUT
cat $OUT
fi
if [ "$2" = "compact" ]; then
OUT2=gpurun_out/${TAG}_sanitizer_compact.txt
echo "compute-sanitizer over gen_moves<.., COMPACT> (movers handed to lanes): Standard, 192 games x 6 fused steps + observe + state-based sampler, compact 0 and 1 compared (tools/sanitize_compact.py)" > $OUT2
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $OUT2
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_compact.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize run ok|SUMMARY|hazard|Invalid|Error" | head -12 >> $OUT2
done
cat $OUT2
fi
