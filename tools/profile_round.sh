#!/bin/bash
# Regenerates the ncu evidence of a round on a GPU box (run through gpurun; every program first runs plain, then under ncu).
#   bash tools/profile_round.sh <out_dir under gpurun_out> [sections]
# sections (default: all): 1m 10m sortvox brute knn.  Reports go to /tmp (too large to bring back), raw CSV pages to <out_dir>.
# Measured on a B200 box: 1m ~9 min (a launch list + a full capture of 60 launches), 10m ~6 min, sortvox ~3 min, brute ~1 min,
# knn ~1 min -- `ncu --set full` replays every kernel ~40 times, so give gpurun a limit per section, not one for everything.
set -u
OUT=${1:-gpurun_out/prof}
SECTIONS=${2:-"1m 10m sortvox brute knn"}
mkdir -p "$OUT"
PS="python tools/profile_scan.py"
run() { echo "+ $*" >> "$OUT/log.txt"; "$@" >> "$OUT/log.txt" 2>&1; }
has() { case " $SECTIONS " in *" $1 "*) return 0;; *) return 1;; esac; }
if has 1m; then
  # ---- 1M points: launch list + full capture of two scans (the summaries keep the last, steady-state one) ----
  run $PS --scans 3 || exit 1
  run ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches_1m.csv" $PS --scans 3
  run ncu --set full --clock-control none --import-source on -o /tmp/full_1m -f $PS --scans 2
  ncu -i /tmp/full_1m.ncu-rep --page raw --csv > "$OUT/full_1m_raw.csv" 2>> "$OUT/log.txt"
fi
if has 10m; then
  # ---- 10M points (map size): the HBM-bound kernels, dense VoxelGrid ----
  run $PS --scans 2 --points 10000000 --radius 0.02 || exit 1
  run ncu --set full --clock-control none -k 'regex:k_crop|k_voxel|k_rs2_|k_normals|k_count|k_compact|k_cell' -o /tmp/full_10m -f $PS --scans 2 --points 10000000 --radius 0.02
  ncu -i /tmp/full_10m.ncu-rep --page raw --csv > "$OUT/full_10m_raw.csv" 2>> "$OUT/log.txt"
fi
if has sortvox; then
  # ---- 10M points, sort-based VoxelGrid: its own kernels and the sort (the last scan only: 8 + 1 + 6 + 2 launches match) ----
  run $PS --scans 2 --points 10000000 --radius 0.02 --voxel-mode 1 || exit 1
  run ncu --set full --clock-control none -k 'regex:k_voxel_keys|k_voxel_heads|k_voxel_centroids|k_rs2_' --launch-skip 17 -o /tmp/full_10m_sv -f $PS --scans 2 --points 10000000 --radius 0.02 --voxel-mode 1
  ncu -i /tmp/full_10m_sv.ncu-rep --page raw --csv > "$OUT/full_10m_sortvox_raw.csv" 2>> "$OUT/log.txt"
fi
if has brute; then
  # ---- brute-force counting (FP32-pipe figure) ----
  run $PS --scans 2 --count-mode 1 || exit 1
  run ncu --set full --clock-control none -k 'regex:k_count_' -o /tmp/full_brute -f $PS --scans 2 --count-mode 1
  ncu -i /tmp/full_brute.ncu-rep --page raw --csv > "$OUT/full_brute_raw.csv" 2>> "$OUT/log.txt"
fi
if has knn; then
  # ---- the k-NN normals ----
  run $PS --scans 2 --knn 32 --radius 0.07 --knn-cap 0.25 || exit 1
  run ncu --set full --clock-control none -k regex:k_normals_knn -o /tmp/full_knn -f $PS --scans 2 --knn 32 --radius 0.07 --knn-cap 0.25
  ncu -i /tmp/full_knn.ncu-rep --page raw --csv > "$OUT/full_knn_raw.csv" 2>> "$OUT/log.txt"
fi
ls -la "$OUT"
