#!/bin/bash
# Run ON THE GPU BOX (gpurun): the ncu evidence committed under profiles/ for the final code state.
#   bash tools/capture_profiles.sh gpurun_out/r2p
# Outputs (small text files only -- the .ncu-rep files stay in /tmp on the box):
#   launches.csv                 every launch of `bench.py --steps 2 --warmup 1` with its device time
#   ncu_raw_<mode>.csv           ncu --set full of the conv launches of one step, raw page
#   ncu_raw_other.csv            ncu --set full of conv1_1, BiLSTM and the proposal kernels
set -u
out=${1:-gpurun_out/r2p}
mkdir -p "$out"
python -c "import bench; print(bench.sources_sha256())" > "$out/sources_sha256.txt"      # what the captures below were taken from
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/launches.csv" \
    python bench.py --steps 2 --warmup 1 --cpu-sample 0 --alt-modes 0 > "$out/bench_under_ncu.log" 2>&1
for m in f16f8:40 bf16x2:16; do
  mode=${m%%:*}; skip=${m##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s "$skip" -c 16 -o "/tmp/prof_$mode" \
      python tools/ncu_step.py --mode "$mode" --steps 2 > "$out/ncu_$mode.log" 2>&1
  ncu -i "/tmp/prof_$mode.ncu-rep" --page raw --csv > "$out/ncu_raw_$mode.csv" 2>/dev/null
done
# everything that is not conv_tc: second step of the f16f8 run (the first step has 12 matching launches: conv1_1 x 3 with the
# calibration passes, BiLSTM, split_heads, 7 proposal / NMS kernels)
timeout 600 ncu --set full --clock-control none -k regex:"conv1_tc|bilstm|proposal|split_heads|nms_" -s 12 -c 10 -o /tmp/prof_other \
    python tools/ncu_step.py --mode f16f8 --steps 2 > "$out/ncu_other.log" 2>&1
ncu -i /tmp/prof_other.ncu-rep --page raw --csv > "$out/ncu_raw_other.csv" 2>/dev/null
ls -la "$out"
