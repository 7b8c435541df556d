#!/bin/bash
# ncu captures of the headline kernel only (GPU box, one gpurun call): launch list of the default stub bench (episode stream)
# and one `--set full` capture of k_episode with ONE episode per game (the configuration roofline.traffic is quoted for);
# summarised on the box into gpurun_out/prof (scripts/summarise_profiles.py keeps the other entries of r02_traffic.json).
set -x
S="python bench.py --workload stub --steps 2 --warmup 1 --no-cpu --edge-frac 0.12"
B="$S --stub-stream-mult 1"
$S > gpurun_out/r02_plain_stub.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stub.csv $S > gpurun_out/r02_ncu_launches.log 2>&1
$B > gpurun_out/r02_plain_stub2.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_episode -s 1 -c 1 -o gpurun_out/r02_k_episode $B > gpurun_out/r02_ncu_episode.log 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/summarise_profiles.py gpurun_out/prof
ncu -i gpurun_out/r02_k_episode.ncu-rep --page source --csv --print-source sass > gpurun_out/prof/r02_k_episode_source.csv 2>/dev/null
rm -f gpurun_out/r02_k_episode.ncu-rep
