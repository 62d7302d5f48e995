cd /root/repo
for fr in 128 256 384; do
  timeout 600 python bench.py --config c4 --frames-per-step $fr --no-others --no-host-cpp --no-cpu-baseline --steps 10 2>gpurun_out/c4_$fr.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('F $fr', round(d['frames_per_s'],1), round(d['roofline']['frac'],4), d['check']['parity'], round(d['e2e']['frames_per_s'],1), d['kernel_ms'] if 'kernel_ms' in d else '')"
done
python tools/bench_k4.py --frames 256 | cut -c1-200
