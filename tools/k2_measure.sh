# One GPU: Schur-path GPU tests; optimize_extrinsics at c5 size (8 cameras x 100 k views) with per-kernel times; ncu of the K2 kernels.
set -x
python -m pytest tests -m gpu -q -k "extrinsics or intrinsics or schur or cov or c1 or c3 or cpp" > gpurun_out/s3_pytest_gpu_k2.log 2>&1; tail -3 gpurun_out/s3_pytest_gpu_k2.log
python tools/extr_probe.py 8 100000 2>/dev/null | tail -1 > gpurun_out/s3_extr_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/s3_launches_extrinsics_c5size.csv python tools/extr_probe.py 8 100000 > gpurun_out/ncu_e2.log 2>&1
python tools/extr_probe.py 8 20000 > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_schur_syrk|k_schur_factor|k_reduced_solve|k_backsub|k_view_chol|k_schur_reduce|k_view_gather" -s 7 -c 7 -f -o gpurun_out/r2_k2_schur python tools/extr_probe.py 8 20000 > gpurun_out/ncu_e.log 2>&1
python tools/solve_probe.py c1 c3 > gpurun_out/s3_solve_probe.jsonl 2>/dev/null
cat gpurun_out/s3_extr_probe.txt gpurun_out/s3_solve_probe.jsonl
