set -x
python -m pytest tests -m gpu -q -k "cov or c3 or extrinsics" 2>&1 | tail -2
python tools/solve_probe.py c3 2>/dev/null
python bench.py --workload c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err || tail -5 gpurun_out/r2_bench_c3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3.json')); print(d['e2e']['wall_s'], d['e2e']['solve_ms'], d['e2e']['create_ms'])"
