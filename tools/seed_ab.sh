# One GPU: the seeding DLT kernel at 2 (default build), 3 and 4 CTAs per SM; ncu launch times of k_view_dlt.
for d in _build _build_d3 _build_d4 _build _build_d3 _build_d4; do echo "== $d"; CALIB_B200_BUILD_DIR=$d python tools/perf_probe.py seed 20000 2>/dev/null | tail -1 | cut -c1-300; done
for d in _build _build_d3 _build_d4; do CALIB_B200_BUILD_DIR=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_view_dlt --csv python tools/perf_probe.py seed 20000 2>/dev/null | grep k_view_dlt | awk -F'","' '{print "'$d'", $5, $NF}' | head -4; done
