#!/usr/bin/env bash
# Round-2 verification call: whole GPU suite, paste bench, config 3 with the eval script's paste_params (both arms), default bench.
mkdir -p gpurun_out/call_a
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/call_a/gpu_tests.log
timeout 100 python bench_paste.py 2>&1 | tail -1 | tee gpurun_out/call_a/bench_paste_ours.json
timeout 400 python bench_config3.py --arm reference --paste --out gpurun_out/call_a/c3p_ref.pt 2>&1 | tail -1 | tee gpurun_out/call_a/c3p_ref.json
timeout 400 python bench_config3.py --arm ours --paste --out gpurun_out/call_a/c3p_ours.pt 2>&1 | tail -4 | tee gpurun_out/call_a/c3p_ours.json
timeout 100 python bench_config3.py --compare gpurun_out/call_a/c3p_ref.pt gpurun_out/call_a/c3p_ours.pt 2>&1 | tail -1 | tee gpurun_out/call_a/c3p_compare.json
rm -f gpurun_out/call_a/*.pt
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/call_a/bench.json
