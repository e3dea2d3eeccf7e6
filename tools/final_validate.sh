# One GPU, the round's closing run: GPU suite, smoke, every bench record, Schur-path probes and ncu captures.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/final_measure.sh > gpurun_out/final_measure.log 2>&1
python tools/extr_probe.py 8 20000 > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_schur_syrk|k_schur_factor|k_reduced_solve|k_backsub|k_view_chol|k_schur_reduce|k_view_gather" -s 7 -c 7 -f -o gpurun_out/r2_k2_schur python tools/extr_probe.py 8 20000 > gpurun_out/ncu_e.log 2>&1
python tools/extr_probe.py 8 100000 > gpurun_out/plain_e2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r2_launches_extrinsics_c5size.csv python tools/extr_probe.py 8 100000 > gpurun_out/ncu_e2.log 2>&1
ls -la gpurun_out/r2_k2_schur.ncu-rep
