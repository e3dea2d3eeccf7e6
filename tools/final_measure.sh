# One GPU: the JSON lines kept under profiles/ (bench of every named shape), the FP64 microbenchmarks, small-solve timings.
set -x
for w in c1 c3 c4 c4-axxb c2; do python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err || tail -5 gpurun_out/r2_bench_$w.err; done
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_ref.err
{ tools/ubench_fp64_occ; tools/ubench_fp64_mix; tools/ubench_k1_exchange; } > gpurun_out/r2_ubench_fp64.txt 2>&1
tools/ubench_dmma > gpurun_out/r2_ubench_dmma.txt 2>&1
python tools/solve_probe.py c1 c3 c4 > gpurun_out/r2_solve_probe.jsonl 2>/dev/null
for i in 1 2 3; do python tools/extr_probe.py 8 100000 2>/dev/null | tail -1; done > gpurun_out/r2_extr_probe.txt
tail -n 3 gpurun_out/r2_solve_probe.jsonl gpurun_out/r2_extr_probe.txt
