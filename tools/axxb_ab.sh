# One GPU: bench record and ncu capture of the on-the-fly AX = XB kernel.
set -x
python bench.py --workload c4-axxb > gpurun_out/r2_bench_c4-axxb.json 2> gpurun_out/r2_bench_c4-axxb.err || tail -5 gpurun_out/r2_bench_c4-axxb.err
A="--workload c4-axxb --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $A > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_axxb_otf -s 2 -c 1 -f -o gpurun_out/r2_k_axxb_otf python bench.py $A > gpurun_out/ncu_a.log 2>&1
python tools/perf_probe.py axxb 3000 2>/dev/null | tail -1 > gpurun_out/r2_axxb_pairs_probe.json
tail -c 400 gpurun_out/r2_bench_c4-axxb.json
