#!/bin/bash
# Last GPU call of the round: the whole GPU suite and smoke() on the committed build, then grid-size sweeps of the
# 3D kernels (environment knobs only, no rebuild).
mkdir -p gpurun_out
python -u -m pytest tests -m gpu -x -q --timeout=200 --timeout-method=thread --durations=6 2>&1 | tail -14 > gpurun_out/gputest_summary.txt
tail -3 gpurun_out/gputest_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
for kv in "FIDGET_B200_BLOCKS_PER_SM=6" "FIDGET_B200_BLOCKS_PER_SM=4" "FIDGET_B200_BLOCKS_PER_SM=8" "FIDGET_B200_BLOCKS_PER_SM=12" \
          "FIDGET_B200_VOXEL_BLOCKS_PER_SM=10" "FIDGET_B200_VOXEL_BLOCKS_PER_SM=16"; do
  echo "$kv"
  env $kv timeout 60 python -u scripts/bench_configs.py slab bear 2>&1 | grep "^{" | cut -c1-330
done 2>&1 | tee gpurun_out/sweep.log
