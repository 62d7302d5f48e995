#!/bin/bash
# One-GPU evidence run for profiles/: bench line, ncu launch list of the same command, full ncu capture of the dominant
# kernels.  Everything lands in gpurun_out/ (scratch); the summaries are copied into profiles/ by hand.
# usage (under gpurun):  bash tools/profile_round.sh r02
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 400 $out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python tools/bench_k4.py --frames 64 > $out/${tag}_k4_64.jsonl 2>> $out/${tag}_bench.err
python tools/bench_k4.py --frames 128 > $out/${tag}_k4_128.jsonl 2>> $out/${tag}_bench.err
python tools/bench_k4.py --timeline > $out/${tag}_k4_timeline.txt 2>> $out/${tag}_bench.err
# launch list of the bench command (cold-cache, serialised: shares must agree with bench.py, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-host-cpp > $out/${tag}_ncu_l.log 2>&1
# full capture of one k3_rdf launch of the bench (symmetric kernel, 64 frames: the launch the headline times)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k3_rdfILi8ELi128ELb1ELb1ELb0ELi1 --launch-skip 3 --launch-count 1 -o $out/${tag}_k3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-others --no-host-cpp > $out/${tag}_ncu_k3.log 2>&1
# K4 of the config-4 box, distinct frames
ncu --set full --clock-control none --import-source on -k regex:k_stream -c 1 -o $out/${tag}_k4 python tools/bench_k4.py --one > $out/${tag}_ncu_k4.log 2>&1
# config 5: per-pair times of one frame, launch list, full capture of the water-water launch of the cell kernel
python tools/bench_configs.py --only-c5 > $out/${tag}_c5_pairs.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_c5_launches.csv \
    python tools/bench_configs.py --only-c5 > $out/${tag}_ncu_c5l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_rdf_cellw --launch-skip 0 --launch-count 1 -o $out/${tag}_cellw \
    python tools/bench_configs.py --only-c5 > $out/${tag}_ncu_cellw.log 2>&1
ls -la $out | tail -12
