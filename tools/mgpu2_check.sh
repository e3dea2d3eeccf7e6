# Two GPUs: the view-sharded optimize_extrinsics (solve, covariance), sharded AX = XB and the RANSAC split against one GPU; the c5 bench on two ranks.
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_views_check.py > gpurun_out/r2_mgpu_views_n2.txt 2>&1; tail -3 gpurun_out/r2_mgpu_views_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_misc_check.py > gpurun_out/r2_mgpu_misc_n2.txt 2>&1; tail -4 gpurun_out/r2_mgpu_misc_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 600 gpurun_out/r2_bench_n2.json
