#!/bin/bash
# The BASELINE configs that are not the bench line, through the headless C++ driver (run on a GPU box):
#   tools/other_configs.sh > gpurun_out/r2_other_configs.jsonl
B=gmu-water-simulation_b200/sph_bench
run() { echo -n "{\"cmd\": \"sph_bench $*\", \"result\": "; $B "$@" 2>/dev/null | tr -d '\n'; echo "}"; }
run --scenario dam_break --box 3.62 --steps 300 --warmup 200
run --scenario dam_break --box 3.62 --steps 100 --warmup 200 --phases
run --scenario dam_break --box 0.9 --steps 1000 --warmup 100
run --scenario dam_break --box 1.14 --brute --steps 100 --warmup 10 --phases
run --scenario fountain --box 2.28 --nozzles 64 --steps 600 --warmup 0
run --scenario fountain --box 4.57 --nozzles 256 --steps 1200 --warmup 0
run --scenario dam_break --box 3.62 --steps 20 --warmup 200 --mirror 2
run --scenario dam_break --box 3.62 --steps 100 --warmup 200 --mirror 3
run --scenario dam_break --box 3.62 --steps 100 --warmup 200 --mirror 3 --mirror-stride 4
