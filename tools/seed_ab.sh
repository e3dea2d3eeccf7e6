# One GPU: the seeding DLT kernel at 2, 3 (the default build) and 4 CTAs per SM; ncu launch times of k_view_dlt.
# The comparison builds are made HERE (no GPU needed) before the gpurun call:
#   for m in 2 4; do CALIB_B200_BUILD_DIR=_build_d$m CALIB_B200_NVCC_EXTRA="-DCALK_DLT_MINB=$m" python -c "from calibration_b200 import build as b; b.build(force=False)"; done
for d in _build_d2 _build _build_d4 _build_d2 _build _build_d4; do echo "== $d"; CALIB_B200_BUILD_DIR=$d python tools/perf_probe.py seed 20000 2>/dev/null | tail -1 | cut -c1-300; done
for d in _build_d2 _build _build_d4; do CALIB_B200_BUILD_DIR=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_view_dlt --csv python tools/perf_probe.py seed 20000 2>/dev/null | grep k_view_dlt | awk -F'","' '{print "'$d'", $5, $NF}' | head -4; done
