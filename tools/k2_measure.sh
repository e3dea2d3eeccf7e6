# One GPU: Schur-path GPU tests; optimize_extrinsics at c5 size (8 cameras x 100 k views); small solves.
set -x
python -m pytest tests -m gpu -q -k "extrinsics or cov or c3" > gpurun_out/s3_pytest_gpu_k2.log 2>&1; tail -3 gpurun_out/s3_pytest_gpu_k2.log
for i in 1 2 3 4 5; do python tools/extr_probe.py 8 100000 2>/dev/null | tail -1; done > gpurun_out/s3_extr_probe.txt
python tools/solve_probe.py c1 c3 > gpurun_out/s3_solve_probe.jsonl 2>/dev/null
cat gpurun_out/s3_extr_probe.txt gpurun_out/s3_solve_probe.jsonl
