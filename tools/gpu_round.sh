#!/bin/bash
# One GPU-box visit: GPU tests, bench line, reference arm, launch list, ncu captures.
set -u
mkdir -p gpurun_out
./tools/mb_fp64 > gpurun_out/mb_fp64.json 2>&1; cat gpurun_out/mb_fp64.json
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
tail -3 gpurun_out/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
# main launch of the grouped hierarchical kernel: launches alternate validation (256 points) / batch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_hier_groups -s 5 -c 1 -f -o gpurun_out/h2_c2 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_h2.log 2>&1; echo "ncu h2 rc=$?"
SGM_JIT_SYNC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgm_small_jit -s 2 -c 1 -f -o gpurun_out/jit_ex1 python tools/prof_ex1.py > gpurun_out/ncu_jit.log 2>&1; echo "ncu jit rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lc_brownian_tiles -s 2 -c 1 -f -o gpurun_out/lc_bulk python tools/exp_lc.py > gpurun_out/ncu_lc.log 2>&1; echo "ncu lc rc=$?"
