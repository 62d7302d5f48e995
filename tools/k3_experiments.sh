set -u
B="python bench.py --no-others --no-host-cpp --no-cpu-baseline --steps 10"
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["frames_per_s"],1), "frames/s  frac", round(d["roofline"]["frac"],4), "k3_ms", round(d["kernel_ms"]["k3_rdf"],3), d["check"]["parity"][:40])'
for rep in 1 2; do
$B | python -c "$pick" default
B2A_LIB=$PWD/build/exp/libb2a_dbuf.so $B | python -c "$pick" dbuf
B2A_LIB=$PWD/build/exp/libb2a_nokmin.so $B | python -c "$pick" nokmin
done
for c in c3 c5; do
python bench.py --config $c --no-others --no-host-cpp --no-cpu-baseline --steps 5 | python -c "$pick" default-$c
B2A_LIB=$PWD/build/exp/libb2a_dbuf.so python bench.py --config $c --no-others --no-host-cpp --no-cpu-baseline --steps 5 | python -c "$pick" dbuf-$c
done
ncu --set full --clock-control none --import-source on -k regex:k3_rdf --launch-skip 6 --launch-count 2 -o gpurun_out/r02_k3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-others --no-host-cpp --frames-per-step 2 > gpurun_out/r02_ncu_k3.log 2>&1; tail -2 gpurun_out/r02_ncu_k3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-host-cpp > gpurun_out/r02_ncu_l.log 2>&1; tail -1 gpurun_out/r02_ncu_l.log | cut -c1-200
