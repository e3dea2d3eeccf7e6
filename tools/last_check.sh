set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_ref.err
python bench.py --workload c3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
python bench.py --workload c1 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
tail -c 300 gpurun_out/r2_bench_n1.json
