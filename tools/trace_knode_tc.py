"""dev: timeline of one CTA of the tensor-core forward march (library built with make EXTRA=-DKC_TC_TRACE)"""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "knode-cosserat_b200"); sys.path.insert(0, "tools")
import _kc, _ops
from time_knode import main as _m
buf = torch.zeros(3 * 8 * 32 + 128, dtype=torch.int64, device="cuda")
L = C.CDLL(_kc.LIB_PATH)
L.kc_knode_tc_set_trace.argtypes = [C.c_void_p]
assert L.kc_knode_tc_set_trace(C.c_void_p(buf.data_ptr())) == 0
os.environ["KC_TIME_T"] = "8"
sys.argv = ["x", "1024"]
_m()
starts = buf.cpu().numpy()[3 * 8 * 32:]
print('gaps between consecutive node-evaluation starts:', np.diff(starts).tolist())
t = buf.cpu().numpy()[:3 * 8 * 32].reshape(3, 8, 32)
names = {0: "start", 1: "X-arrive", 20: "O-ready", 21: "end"}
for ev in range(2, 3):
    base = t[0, ev, 0]
    print(f"--- evaluation {ev}: previous-eval end -> start (physics) = {t[0, ev, 0] - t[0, ev - 1, 21]}")
    rows = []
    for role, rn in enumerate(["phys", "help", "mma "]):
        for i in range(32):
            if t[role, ev, i]:
                rows.append((int(t[role, ev, i] - base), rn, i))
    for dt, rn, i in sorted(rows):
        print(f"   {dt:7d}  {rn} ev{i}")
