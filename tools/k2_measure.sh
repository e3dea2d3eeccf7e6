# One GPU: Schur-path GPU tests; one LM iteration's launch list of optimize_extrinsics at c5 size (8 cameras x 100 k views).
set -x
python -m pytest tests -m gpu -q -k "extrinsics or intrinsics or cov or c3 or c1 or cpp or wide" > gpurun_out/s3_pytest_gpu_k2.log 2>&1; tail -3 gpurun_out/s3_pytest_gpu_k2.log
python tools/extr_probe.py 8 100000 > gpurun_out/plain_e2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r2_launches_extrinsics_c5size.csv python tools/extr_probe.py 8 100000 > gpurun_out/ncu_e2.log 2>&1
tail -1 gpurun_out/plain_e2.log
