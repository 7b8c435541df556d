#!/bin/bash
# ncu captures of this round (GPU box, one gpurun call): launch list of the bench command, and `--set full` captures of the
# dominant kernels; the reports land in gpurun_out/ and are summarised HERE (no GPU) by scripts/summarise_profiles.py.
set -x
B="python bench.py --workload stub --steps 2 --warmup 1 --no-cpu --edge-frac 0.12"
$B > gpurun_out/r02_plain_stub.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stub.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
$B > gpurun_out/r02_plain_stub2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_episode -s 1 -c 1 -o gpurun_out/r02_k_episode $B > gpurun_out/r02_ncu_episode.log 2>&1
for cfg in "15 15 8192 bf16" "15 15 8192 bf16x3" "20 20 8192 bf16" "15 15 2800 bf16 6"; do
  tag=$(echo $cfg | tr ' ' '_')
  python scripts/net_once.py $cfg > gpurun_out/r02_plain_net_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_net_ -s 4 -c 2 -o gpurun_out/r02_net_$tag python scripts/net_once.py $cfg > gpurun_out/r02_ncu_net_$tag.log 2>&1
done
R="python bench.py --workload real20 --no-cpu --steps 2 --warmup 1 --stream-mult 1"
$R > gpurun_out/r02_plain_real20.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/r02_launches_real20.csv $R > gpurun_out/r02_ncu_launches_real20.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the reports together exceed the 64 MiB gpurun copies back: summarise them here, keep one (source-level view) and drop the rest
python scripts/summarise_profiles.py gpurun_out/prof
ncu -i gpurun_out/r02_net_15_15_8192_bf16.ncu-rep --page source --csv --print-source sass > gpurun_out/prof/r02_net_15_15_8192_bf16_source.csv 2>/dev/null
rm -f gpurun_out/r02_k_episode.ncu-rep gpurun_out/r02_net_15_15_8192_bf16x3.ncu-rep gpurun_out/r02_net_20_20_8192_bf16.ncu-rep gpurun_out/r02_net_15_15_2800_bf16_6.ncu-rep
