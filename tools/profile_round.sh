#!/bin/bash
# One-GPU evidence run for profiles/: bench line, ncu launch list of the same command, full ncu capture of the dominant
# kernels, HBM-kernel bench.  Everything lands in gpurun_out/ (scratch); the summaries are copied into profiles/ by hand.
# usage (under gpurun):  bash tools/profile_round.sh r01
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
python bench.py --symmetric 0 --steps 10 --no-cpu-baseline > $out/${tag}_bench_full.json 2>> $out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python tools/bench_hbm.py > $out/${tag}_hbm.jsonl 2>> $out/${tag}_bench.err
python tools/bench_configs.py --c5 > $out/${tag}_configs.jsonl 2>> $out/${tag}_bench.err
python tools/bench_hbm_napm.py --frames 64 > $out/${tag}_hbm_napm.jsonl 2>> $out/${tag}_bench.err
# launch list of the bench command (cold-cache, serialised: shares must agree with bench.py, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_l.log 2>&1
# full capture of one k3_rdf launch (symmetric kernel of the bench) and of the full-evaluation kernel
ncu --set full --clock-control none --import-source on -k regex:k3_rdf --launch-skip 6 --launch-count 2 -o $out/${tag}_k3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --frames-per-step 2 > $out/${tag}_ncu_k3.log 2>&1
# K1 / K2 / K4 of the HBM bench: one launch each (mangled names select the instantiation; launch-skip passes the warm-ups)
for k in 0ELi3ELi0 1ELi3ELi0 2ELi3ELi1; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_streamILi$k --launch-skip 2 --launch-count 1 \
      -o $out/${tag}_hbm_$k python tools/bench_hbm.py --reps 2 > $out/${tag}_ncu_hbm_$k.log 2>&1
done
ls -la $out | tail -15
