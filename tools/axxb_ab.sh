# One GPU: the AX = XB pair kernels at 2 CTAs per SM (the default build, _build) against 3 (_build_ax1, -DCALK_AXXB_MINB=3), same box.
set -x
python -m pytest tests -m gpu -q -k "axxb or handeye" 2>&1 | tail -2
for d in _build _build_ax1 _build _build_ax1; do CALIB_B200_BUILD_DIR=$d python bench.py --workload c4-axxb --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$d', 'otf pairs/s %.4g' % d['value'], 'ms/pass %.4f' % d['ms_per_step'], 'e2e solve ms', round(d['e2e'].get('solve_ms',0),2))"; done
for d in _build _build_ax1; do CALIB_B200_BUILD_DIR=$d python tools/perf_probe.py axxb 3000 2>/dev/null | tail -1 | cut -c1-200; done
