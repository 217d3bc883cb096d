#!/usr/bin/env bash
# Round-2 final verification: whole GPU suite (incl. CUDA-graph tests), smoke(), config 3 with the backbone / SR head replayed as
# CUDA graphs (plain sweep, and the eval script's paste_params with the occlusion pass rendered from the view's tri-planes).
mkdir -p gpurun_out/final
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/final/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final/smoke.log
timeout 300 python bench_config3.py --arm ours --graphs --reps 2 2>&1 | tail -4 | tee gpurun_out/final/c3_graphs.json
timeout 300 python bench_config3.py --arm ours --graphs --paste --reuse_triplane --reps 1 2>&1 | tail -4 | tee gpurun_out/final/c3p_graphs_reuse.json
