#!/usr/bin/env bash
# Interleaved A/B timing of build/variants/libp3d_<name>.so: 3 rounds x (each variant: bench.py, 100 device-timed steps).
# Box-to-box and run-to-run noise of a single 20-step run is +-3%, the effects being compared are of that size.
# usage: ab_bench.sh <tag> <names...>  ->  gpurun_out/abb_<tag>.txt
tag=$1; shift
out=gpurun_out/abb_$tag.txt
: > $out
for rep in 1 2 3; do
  for name in "$@"; do
    lib=$PWD/build/variants/libp3d_$name.so
    P3D_LIBP3D=$lib timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-variants --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name rep$rep views/s %.1f  ms/step %.4f' % (d['value'], d['ms_per_step']))" >> $out 2>&1
  done
done
sort $out
