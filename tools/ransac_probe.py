"""K3 timing probe for kernel A/B runs: CUDA-event time of k_ransac on n_problems x 500 resident on the device.
    CALIB_B200_BUILD_DIR=_build_x python tools/ransac_probe.py 100000"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from calibration_b200 import abi, capi, synth
npb = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x, y, u, v, _ = synth.synth_ransac(seed=17, n_problems=npb, n=500)
dev = [torch.from_numpy(a).cuda() for a in (x, y, u, v)]
res = torch.empty(npb * C.sizeof(abi.RansacResult), dtype=torch.uint8, device="cuda")
mask = torch.empty(npb * 500, dtype=torch.uint8, device="cuda")
opts = abi.RansacOptions.default(); ms = C.c_float(); L = capi.lib(); t = []
for rep in range(6):
    rc = L.cal_ransac_homography_batch_dev(npb, 500, *[C.c_void_p(d.data_ptr()) for d in dev], C.byref(opts), 1, C.c_void_p(res.data_ptr()), C.c_void_p(mask.data_ptr()), C.byref(ms))
    assert rc == 0
    t.append(ms.value)
print(json.dumps({"build": os.environ.get("CALIB_B200_BUILD_DIR", "_build"), "problems": npb, "ms": t[1:], "checksum": int(mask.sum().item())}))
