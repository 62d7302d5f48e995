set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -k "rdf or stress or config" 2>&1 | tail -5
B="python bench.py --config c5 --no-others --no-host-cpp --no-cpu-baseline --steps 5"
for v in default blocktile w16 nw1 nw4 nw4w24; do
  case $v in
    default) env="";;
    blocktile) env="B2A_CELL_BLOCKTILE=1";;
    *) env="B2A_LIB=/root/repo/build_exp/libb2a_$v.so";;
  esac
  env $env timeout 600 $B > gpurun_out/c5_$v.json 2> gpurun_out/c5_$v.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/c5_$v.json') if l.startswith('{')][0]
    print('$v', round(d['frames_per_s'],2), d['roofline']['frac'], d['check'])
except Exception as e:
    print('$v', 'failed', e); print(open('gpurun_out/c5_$v.err').read()[-1500:])
PY
done
