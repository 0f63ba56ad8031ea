for lib in libdavo_b200.so libdavo_b200_noyield.so libdavo_b200_cap275.so; do
  for cfg in cfg4 cfg2; do
    DAVO_B200_LIB=$PWD/deep-attention-visual-odometry_b200/$lib timeout 300 python bench.py --config $cfg --no-side-configs --no-e2e --no-cpu-baseline --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib $cfg', 'seq %.2f pipe %.2f' % (d['ms_per_step_sequential'], d['ms_per_step_pipelined'] or 0), d['reasons_rank0'], 'conv %.4f' % d['converged_frac'])
"
  done
done
