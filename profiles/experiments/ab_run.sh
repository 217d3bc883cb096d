#!/usr/bin/env bash
# A/B harness for the gpurun box: for every build/variants/libp3d_<name>.so given on the command line run two fused parity
# fixtures, the bench (20 steps, device-timed) and one P3D_WS_TIMING cycle account.  Output: gpurun_out/ab_<tag>.txt
tag=$1; shift
out=gpurun_out/ab_$tag.txt
: > $out
for name in "$@"; do
  lib=$PWD/build/variants/libp3d_$name.so
  echo "=== $name" >> $out
  P3D_LIBP3D=$lib timeout 300 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "fused_tc_3xbf16 and (config1 or headline96 or mid_eval96 or fused48_shared)" 2>&1 | tail -2 >> $out
  P3D_LIBP3D=$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-variants --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('views/s %.1f  ms/step %.3f  kernel ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch']))" >> $out 2>&1
  P3D_LIBP3D=$lib P3D_WS_TIMING=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-variants --no-cpu-baseline 2>&1 | grep "p3d ws timing" | tail -1 >> $out
done
cat $out
