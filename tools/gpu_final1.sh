#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2f_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2f_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2f_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2f_sanitizer_racecheck.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2f_bench_ref.json','gpurun_out/r2f_bench_c4.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f, round(d['value']/1e6,2),'Mbp/s', round(d['ms_per_step'],1), d.get('e2e',{}).get('ms_per_step'), d.get('families_match_oracle_golden'), d.get('sample_bp'), (d.get('roofline') or {}).get('kernel','')[:30], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('traffic'))
PY
