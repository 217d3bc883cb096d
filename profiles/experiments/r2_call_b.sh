#!/usr/bin/env bash
# Round-2 call B: paste inside the training phases (mode 'Agrad', both arms), config 3 with the occlusion pass rendered from the
# view's tri-planes, per-kernel device-time table of the config-3 sweep.
mkdir -p gpurun_out/call_b
timeout 120 python -m pytest tests/test_paste_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/call_b/paste_tests.log
timeout 300 python bench_config5.py --arm reference --paste --iters 2 2>&1 | tail -1 | tee gpurun_out/call_b/c5p_ref.json
timeout 300 python bench_config5.py --arm ours --paste --iters 2 2>&1 | tail -3 | tee gpurun_out/call_b/c5p_ours.json
timeout 300 python bench_config3.py --arm ours --paste --reuse_triplane --reps 1 --profile gpurun_out/call_b/c3p_reuse_kernels.json 2>&1 | tail -3 | tee gpurun_out/call_b/c3p_ours_reuse.json
timeout 300 python bench_config3.py --arm ours --reps 1 --profile gpurun_out/call_b/c3_kernels.json 2>&1 | tail -2 | tee gpurun_out/call_b/c3_ours.json
