set -x
cd $GRAFT_REPO_ROOT
cat > /tmp/lr_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from test_gpu_learner import _examples, _module
from resource_packing_self_play_b200.nnet import DeviceLearner
B = int(sys.argv[1])
ops, recs, items, pis, vs = _examples(B, 15, 15, 10, seed=3)
m = _module(15, 15, 10)
L = DeviceLearner(15, 15, 10, max_batch=B); L.load_state_dict(m.state_dict())
for _ in range(4): L.grad(recs, items, pis, vs)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lr_ -s 39 -c 13 --csv --log-file gpurun_out/lr_launches_fused_512.csv python /tmp/lr_one.py 512 > /dev/null 2>&1
BPP_LEARNER_UNFUSED=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lr_ -s 123 -c 41 --csv --log-file gpurun_out/lr_launches_unfused_512.csv python /tmp/lr_one.py 512 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_lr_stage -s 18 -c 6 -o gpurun_out/lr_stage_512 python /tmp/lr_one.py 512 > /dev/null 2>&1
ncu -i gpurun_out/lr_stage_512.ncu-rep --page raw --csv > gpurun_out/lr_stage_512_raw.csv
ncu -i gpurun_out/lr_stage_512.ncu-rep --page source --csv --kernel-name regex:k_lr_stage_fwd --launch-skip 0 --launch-count 1 > gpurun_out/lr_stage_fwd0_source.csv 2>/dev/null
ls -la gpurun_out/ | head -30
rm -f gpurun_out/lr_stage_512.ncu-rep
