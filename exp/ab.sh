#!/bin/bash
# A/B the linearize kernel variants built into exp/libvo_*.so (see exp/README.md)
for lib in exp/libvo_*.so; do
  VO_B200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'us/launch %.1f' % d['roofline']['us_per_launch'], 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'pose_err %.2e' % d['final']['pose_err_vs_gt'])"
done
