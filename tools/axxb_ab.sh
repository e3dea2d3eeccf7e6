# One GPU: k_axxb_otf at two CTAs per SM (128 registers) against one (186 registers), same box, same process sequence.
set -x
python -m pytest tests -m gpu -q -k "axxb or handeye" 2>&1 | tail -2
for d in _build _build_ax1 _build _build_ax1; do CALIB_B200_BUILD_DIR=$d python bench.py --workload c4-axxb --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$d', 'pairs/s %.4g' % d['value'], 'ms/pass %.4f' % d['ms_per_step'], 'e2e solve ms', round(d['e2e'].get('solve_ms',0),2))"; done
