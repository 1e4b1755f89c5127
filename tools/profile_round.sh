#!/bin/bash
# Regenerates the ncu evidence of a round on a GPU box (run through gpurun; every program first runs plain, then under ncu).
#   bash tools/profile_round.sh <out_dir under gpurun_out>      reports go to /tmp (too large to bring back), summaries to <out_dir>
set -u
OUT=${1:-gpurun_out/prof}
mkdir -p "$OUT"
PS="python tools/profile_scan.py"
run() { echo "+ $*" >> "$OUT/log.txt"; "$@" >> "$OUT/log.txt" 2>&1; }
# ---- 1M points: launch list + full capture of two scans (the summaries keep the last, steady-state one) ----
run $PS --scans 3 || exit 1
run ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches_1m.csv" $PS --scans 3
run ncu --set full --clock-control none --import-source on -o /tmp/full_1m -f $PS --scans 2
ncu -i /tmp/full_1m.ncu-rep --page raw --csv > "$OUT/full_1m_raw.csv" 2>> "$OUT/log.txt"
# ---- 10M points (map size): the HBM-bound kernels, dense and sort-based VoxelGrid ----
run $PS --scans 2 --points 10000000 --radius 0.02 || exit 1
run ncu --set full --clock-control none -k 'regex:k_crop|k_voxel|k_rs2_|k_normals|k_count|k_compact|k_cell' -o /tmp/full_10m -f $PS --scans 2 --points 10000000 --radius 0.02
ncu -i /tmp/full_10m.ncu-rep --page raw --csv > "$OUT/full_10m_raw.csv" 2>> "$OUT/log.txt"
run $PS --scans 2 --points 10000000 --radius 0.02 --voxel-mode 1 || exit 1
run ncu --set full --clock-control none -k 'regex:k_voxel|k_rs2_' -o /tmp/full_10m_sv -f $PS --scans 2 --points 10000000 --radius 0.02 --voxel-mode 1
ncu -i /tmp/full_10m_sv.ncu-rep --page raw --csv > "$OUT/full_10m_sortvox_raw.csv" 2>> "$OUT/log.txt"
# ---- brute-force counting (FP32-pipe figure) and the k-NN normals ----
run $PS --scans 2 --count-mode 1 || exit 1
run ncu --set full --clock-control none -k 'regex:k_count_' -o /tmp/full_brute -f $PS --scans 2 --count-mode 1
ncu -i /tmp/full_brute.ncu-rep --page raw --csv > "$OUT/full_brute_raw.csv" 2>> "$OUT/log.txt"
run $PS --scans 2 --knn 32 --radius 0.07 --knn-cap 0.25 || exit 1
run ncu --set full --clock-control none -k regex:k_normals_knn -o /tmp/full_knn -f $PS --scans 2 --knn 32 --radius 0.07 --knn-cap 0.25
ncu -i /tmp/full_knn.ncu-rep --page raw --csv > "$OUT/full_knn_raw.csv" 2>> "$OUT/log.txt"
ls -la "$OUT"
