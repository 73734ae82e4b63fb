#!/bin/bash
# Runs the reference's own CUDA drivers (compiled unmodified for sm_100a into oracle/_ref by `make -C oracle ref_gpu`)
# and the B200 library's mirrors of the same drivers on this GPU, so the two can be read side by side.
# Test tooling: output goes to gpurun_out/ and is summarised in profiles/.
cd "$(dirname "$0")/../.."
out=${1:-gpurun_out/reference_drivers.log}
: > "$out"
for b in ref_v1_base ref_v1_opt1 ref_tiled_d_base ref_tiled_d_opt ref_v2_base ref_v2_opt; do
  if [ -x oracle/_ref/$b ]; then
    echo "===== reference $b (sm_100a build of the reference's driver.cu)" >> "$out"
    timeout 300 oracle/_ref/$b 2>&1 | grep -E "Kernel|B=|CPU|GPU time|Speedup|absolute|PASS|FAIL|Grid|KV_TILES" >> "$out"
  fi
done
for w in v1 tiled_d v2; do
  echo "===== libfa_b200 mirror: tests/drivers/driver.py $w" >> "$out"
  timeout 300 python -m tests.drivers.driver $w --heads 16 2>&1 | grep -E "Kernel|B=|CPU time|GPU time|Speedup|Throughput|absolute|PASS|FAIL" >> "$out"
done
cat "$out"
