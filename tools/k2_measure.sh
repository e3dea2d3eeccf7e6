# One GPU: GPU suite, the DMMA microbenchmark, optimize_extrinsics at c5 size (8 cameras x 100 k views) with per-kernel times.
set -x
python -m pytest tests -m gpu -q > gpurun_out/s3_pytest_gpu.log 2>&1; tail -3 gpurun_out/s3_pytest_gpu.log
tools/ubench_dmma > gpurun_out/r2_ubench_dmma.txt 2>&1
python tools/extr_probe.py 8 100000 2>/dev/null | tail -1 > gpurun_out/s3_extr_probe.txt
CALIB_B200_BUILD_DIR=_build_f3 python tools/extr_probe.py 8 100000 2>/dev/null | tail -1 > gpurun_out/s3_extr_probe_f3.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 26 --csv --log-file gpurun_out/s3_launches_extrinsics_c5size.csv python tools/extr_probe.py 8 100000 > gpurun_out/ncu_e2.log 2>&1
CALIB_B200_BUILD_DIR=_build_f3 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 26 --csv --log-file gpurun_out/s3_launches_extrinsics_c5size_f3.csv python tools/extr_probe.py 8 100000 > gpurun_out/ncu_e3.log 2>&1
python tools/solve_probe.py c1 c3 c4 > gpurun_out/s3_solve_probe.jsonl 2>/dev/null
cat gpurun_out/s3_extr_probe.txt gpurun_out/s3_extr_probe_f3.txt gpurun_out/s3_solve_probe.jsonl
