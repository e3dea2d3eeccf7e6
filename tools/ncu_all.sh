set -x
B="--steps 2 --warmup 1 --no-cpu-baseline --no-board-probe"
python bench.py $B > gpurun_out/plain_k1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 4 -c 1 -f -o gpurun_out/r2_k1_final python bench.py $B > gpurun_out/ncu_k1f.log 2>&1
python bench.py $B > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_c5.csv python bench.py $B > gpurun_out/ncu_l.log 2>&1
R="--workload c2 --ransac-problems 20000 --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $R > gpurun_out/plain_r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ransac -s 1 -c 1 -f -o gpurun_out/r2_k_ransac python bench.py $R > gpurun_out/ncu_r.log 2>&1
A="--workload c4-axxb --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $A > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_axxb_otf -s 2 -c 1 -f -o gpurun_out/r2_k_axxb_otf python bench.py $A > gpurun_out/ncu_a.log 2>&1
python tools/extr_probe.py 8 20000 > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_schur_syrk|k_schur_factor|k_reduced_solve|k_backsub" -s 4 -c 4 -f -o gpurun_out/r2_k2_schur python tools/extr_probe.py 8 20000 > gpurun_out/ncu_e.log 2>&1
python tools/perf_probe.py seed 20000 > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_view_dlt -s 1 -c 1 -f -o gpurun_out/r2_k_view_dlt python tools/perf_probe.py seed 20000 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep
