# One GPU: the AX = XB pair kernels at 1, 2 (the default build) and 3 CTAs per SM, same box; then the bench record and the ncu capture.
# The comparison builds are made HERE (no GPU needed) before the gpurun call:
#   for m in 1 3; do CALIB_B200_BUILD_DIR=_build_ax$m CALIB_B200_NVCC_EXTRA="-DCALK_AXXB_MINB=$m" python -c "from calibration_b200 import build as b; b.build(force=False)"; done
set -x
for d in _build_ax1 _build _build_ax3 _build_ax1 _build _build_ax3; do [ -d calibration_b200/$d ] || continue; CALIB_B200_BUILD_DIR=$d python bench.py --workload c4-axxb --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$d', 'otf pairs/s %.4g' % d['value'], 'ms/pass %.4f' % d['ms_per_step'], 'e2e solve ms', round(d['e2e'].get('solve_ms',0),2))"; done
for d in _build_ax1 _build _build_ax3; do [ -d calibration_b200/$d ] || continue; CALIB_B200_BUILD_DIR=$d python tools/perf_probe.py axxb 3000 2>/dev/null | tail -1 | cut -c1-200; done
python bench.py --workload c4-axxb > gpurun_out/r2_bench_c4-axxb.json 2> gpurun_out/r2_bench_c4-axxb.err || tail -5 gpurun_out/r2_bench_c4-axxb.err
A="--workload c4-axxb --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $A > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_axxb_otf -s 2 -c 1 -f -o gpurun_out/r2_k_axxb_otf python bench.py $A > gpurun_out/ncu_a.log 2>&1
