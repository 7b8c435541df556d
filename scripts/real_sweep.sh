#!/bin/bash
# resident-games sweep of the real-net workloads (GPU only): value, evaluator batch, steps
out=gpurun_out/r02_real_sweep.txt
: > $out
for g in 4096 8192 16384; do
  for w in real20 real15; do
    python bench.py --workload $w --no-cpu --steps 10 --games $g --stream-mult 4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['secondary']['$w']
if 'error' in r: print('$w G=$g', r); raise SystemExit
rf = r['roofline']
print('$w G=$g mode %s value %.1f M sims/s e2e %.1f evals %.2f M/s steps/batch %.0f fill %.3f eval %.0f us exp %.0f us TF %.1f' % (r['precision_mode'], r['value']/1e6, r['e2e']['value']/1e6, r['leaf_evals_per_sec']/1e6, r['lockstep_steps_per_batch'], r['leaf_batch_fill'], rf['evaluator_us_per_launch'], rf['expand_select_us_per_launch'], rf['achieved'] or 0))
if 'bf16' in r and r['bf16'] and 'value' in r['bf16']: print('     bf16 value %.1f M sims/s' % (r['bf16']['value']/1e6))
" >> $out
  done
done
cat $out
