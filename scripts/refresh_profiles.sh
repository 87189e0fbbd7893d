#!/bin/bash
# Run on the GPU box (one GPU): the bench line, its ncu launch list, and `ncu --set full` captures of one 2D frame
# and of a 3D render, summarised on the box (gpurun_out/ may not exceed 64 MiB: the 3D report is deleted after its
# summaries are written).
mkdir -p gpurun_out
python -u bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -s 14 -c 7 -f -o gpurun_out/frame2d \
    python scripts/profile_targets.py frame2d > gpurun_out/ncu_frame2d.log 2>&1
python scripts/ncu_table.py gpurun_out/frame2d.ncu-rep > gpurun_out/frame2d_table.json
python scripts/ncu_lines.py gpurun_out/frame2d.ncu-rep 25 > gpurun_out/frame2d_lines.txt
ncu --set full --clock-control none --import-source on -k regex:'k_voxels_3d|k_interval_level|k_interval_root_coop|k_normals_3d' \
    -s 5 -c 5 -f -o /tmp/volume3d python scripts/profile_targets.py volume3d > gpurun_out/ncu_volume3d.log 2>&1
python scripts/ncu_table.py /tmp/volume3d.ncu-rep > gpurun_out/volume3d_table.json
python scripts/ncu_lines.py /tmp/volume3d.ncu-rep 25 > gpurun_out/volume3d_lines.txt
rm -f gpurun_out/frame2d.ncu-rep
du -sh gpurun_out; tail -c 400 gpurun_out/bench_n1.json
