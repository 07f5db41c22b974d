cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q -k "vgg or conv or filter_grad or wgrad or reproducible" 2>&1 | tail -4
for lib in base new; do
  if [ $lib = base ]; then export AGB200_LIB=$GRAFT_REPO_ROOT/rust-autograd_b200/lib/libagb200_base.so; else unset AGB200_LIB; fi
  timeout 150 python bench.py --mode 3xtf32 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$lib', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['clocks'])"
  timeout 100 python scripts/bench_conv3x.py 0 2>&1 | tail -5
done
