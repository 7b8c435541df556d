#!/bin/bash
# A/B sweep of the role-kernel plans (CTAs per SM per role, group sizes) with scripts/net_perf.py; GPU only
out=gpurun_out/r02_role_sweep.txt
: > $out
for cfg in "2,2,2" "2,1,1" "1,1,1" "2,2,1" "2,1,2"; do
  echo "== BPP_ROLE_CTAS=$cfg" >> $out
  BPP_ROLE_CTAS=$cfg BPP_TC_VERBOSE=1 python scripts/net_perf.py bf16 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); b = d['bf16']; print(d['cfg'], d['B'], '%.2f M evals/s' % (b['evals_per_s'] / 1e6))
    elif 'role' in line and '(bf16)' in line: print(line.strip())
" >> $out
done
echo "== x3 (roles + tensor-core heads)" >> $out
BPP_TC_VERBOSE=1 python scripts/net_perf.py bf16x3 2>&1 | grep -v "^bpp_net: tcgen05" | cut -c1-400 >> $out
echo "== x3 one-kernel trunk" >> $out
BPP_NO_ROLES=1 python scripts/net_perf.py bf16x3 2>&1 | cut -c1-200 >> $out
