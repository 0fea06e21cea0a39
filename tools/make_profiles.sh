#!/bin/bash
# Regenerates the profiling artefacts under gpurun_out/ (run on the GPU box through gpurun, one GPU):
#   bench line, ncu launch list of the SAME bench command, ncu --set full of the two neighbour kernels, clocks.
# Afterwards, here: tools/ncu_summary.py / tools/launch_shares.py / tools/kernel_traffic.py turn them into the
# committed summaries under profiles/ (kernel_traffic.json is keyed by the sha256 of the library that was profiled).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 --no-scaling-baseline > gpurun_out/${TAG}_bench_steps20.json 2>> gpurun_out/${TAG}_bench.err
kill $SMI
# launch list of the same command (shorter timed region; the pre-roll launches are skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1450 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 5 --e2e-steps 2 --no-cpu-baseline --no-scaling-baseline > gpurun_out/${TAG}_bench_under_ncu.json 2>&1
# full sections for the two dominant kernels (step 201)
ncu --set full --clock-control none --import-source on -k regex:"k_density_mask|k_forces_mask" -s 400 -c 2 -o gpurun_out/${TAG}_neighbour_kernels \
    python tools/profile_step.py > gpurun_out/${TAG}_profile_step.log 2>&1
# summaries of the capture (the same commands work here, without a GPU, on the .ncu-rep that comes back)
ncu -i gpurun_out/${TAG}_neighbour_kernels.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_neighbour_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_ncu_neighbour_kernels_raw.csv > gpurun_out/${TAG}_ncu_neighbour_kernels.txt
python tools/kernel_traffic.py gpurun_out/${TAG}_ncu_neighbour_kernels_raw.csv profiles/${TAG}_ncu_neighbour_kernels.txt > gpurun_out/${TAG}_kernel_traffic.json
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_shares.txt
python tools/timeline.py > gpurun_out/${TAG}_timeline_graph.txt 2>&1
# then, here: cp the ${TAG}_* summaries into profiles/ and ${TAG}_kernel_traffic.json to profiles/kernel_traffic.json
tail -c 600 gpurun_out/${TAG}_bench.json
