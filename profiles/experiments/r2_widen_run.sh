#!/usr/bin/env bash
# One GPU-box call for the round-2 widening rows (8f-3 paste, 8f-4 image output) + the final ncu evidence of the default
# renderer:  gpurun --timeout 900 -- 'bash profiles/experiments/r2_widen_run.sh'
mkdir -p gpurun_out/widen
timeout 300 python -m pytest tests/test_imageio_gpu.py tests/test_paste_gpu.py -x -q 2>&1 | tail -25 | tee gpurun_out/widen/tests.log
timeout 100 python bench_paste.py 2>&1 | tail -2 | tee gpurun_out/widen/bench_paste_ours.json
timeout 100 python bench_paste.py --impl reference 2>&1 | tail -2 | tee gpurun_out/widen/bench_paste_ref.json
timeout 200 python bench_imageio.py 2>&1 | tail -2 | tee gpurun_out/widen/bench_imageio.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_paste_front -s 3 -c 1 -f -o gpurun_out/widen/r2_paste \
    python bench_paste.py > gpurun_out/widen/ncu_paste.log 2>&1
timeout 300 bash profiles/experiments/ncu_ws3.sh r2_final
ls -la gpurun_out/widen gpurun_out/r2_final* 2>/dev/null | tail -20
