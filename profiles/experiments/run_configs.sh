#!/usr/bin/env bash
# BASELINE configs[2] (G.f sweep, SR on) and configs[4] (training phases) on the GPU box: the unmodified reference on its own GPU
# path (eager PyTorch + its JIT plugins, prebuilt by baseline/install_ref.sh) next to the same reference modules with our drop-in.
# Outputs: gpurun_out/c3_*.json, c5_*.json (the .pt image dumps are compared and deleted: they exceed the merge-back limit).
set -u
out=gpurun_out
export TORCH_EXTENSIONS_DIR=$PWD/baseline/_ref/_torch_ext
t() { local t0=$SECONDS; "$@"; echo "[$((SECONDS - t0)) s wall]"; }
t timeout 900 python bench_config3.py --arm reference --out $out/c3_ref.pt > $out/c3_ref.json 2> $out/c3_ref.err; tail -2 $out/c3_ref.err; cat $out/c3_ref.json
t timeout 900 python bench_config3.py --arm ours --out $out/c3_ours.pt > $out/c3_ours.json 2> $out/c3_ours.err; tail -2 $out/c3_ours.err; cat $out/c3_ours.json
timeout 300 python bench_config3.py --compare $out/c3_ref.pt $out/c3_ours.pt > $out/c3_compare.json 2> $out/c3_compare.err; cat $out/c3_compare.json; tail -2 $out/c3_compare.err
rm -f $out/c3_ref.pt $out/c3_ours.pt
t timeout 900 python bench_config5.py --arm reference > $out/c5_ref.json 2> $out/c5_ref.err; tail -2 $out/c5_ref.err; cat $out/c5_ref.json
t timeout 900 python bench_config5.py --arm ours > $out/c5_ours.json 2> $out/c5_ours.err; tail -2 $out/c5_ours.err; cat $out/c5_ours.json
