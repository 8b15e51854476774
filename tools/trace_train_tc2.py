"""dev: timeline of CTA 0 of kc_train_tc2_kernel / kc_train_tc3_kernel (library built with make EXTRA=-DKC_TC2_TRACE or
-DKC_TC3_TRACE; KC_TRACE_GEN=2|3 selects the kernel): python tools/trace_train_tc2.py [B] [tile]"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
GEN = os.environ.get("KC_TRACE_GEN", "3")
os.environ["KC_TRAIN_TC"] = GEN
import _kc, _ops
from cosserat_ode import CosseratRod
from cosserat_ode_torch import CosseratRodTorch
from knode import setup_robot
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
want = int(sys.argv[2]) if len(sys.argv) > 2 else 2
robot = CosseratRod(use_fsolve=True); setup_robot(robot)
ctl = torch.tensor(synthetic_tensions(B, 30, robot.del_t, seed=0), device="cuda")
traj, _, _ = _ops.rollout(_kc.rod_params(robot), None, ctl)
torch.manual_seed(0)
tr = CosseratRodTorch("cuda", 512); setup_robot(tr)
for _ in range(3):
    tr.teacher_forced_step(traj, ctl, [3, 5, 7, 9])
torch.cuda.synchronize()
buf = torch.zeros(3 * 2048, dtype=torch.int64, device="cuda")
L = C.CDLL(_kc.LIB_PATH)
set_trace = getattr(L, f"kc_train_tc{GEN}_set_trace")
set_trace.argtypes = [C.c_void_p]
assert set_trace(C.c_void_p(buf.data_ptr())) == 0
tr.teacher_forced_step(traj, ctl, [3, 5, 7, 9])
torch.cuda.synchronize()
assert set_trace(C.c_void_p(0)) == 0
t = buf.cpu().numpy().reshape(3, 2048)
ev = [[(int(v >> 48), int(v & ((1 << 48) - 1))) for v in t[r] if v] for r in range(3)]
t0 = min(e[0][1] for e in ev if e)
print("events per role", [len(e) for e in ev], "kernel span (cycles)", max(e[-1][1] for e in ev if e) - t0)
tiles = [c for i, c in ev[0] if i == 0]
print("tile starts (epilogue warp 0):", [c - t0 for c in tiles], "diffs", np.diff(tiles).tolist())
lo = tiles[want]; hi = tiles[want + 1] if want + 1 < len(tiles) else 1 << 62
rows = []
for r, nm in enumerate(["epi0", "epi4", "mma "]):
    for i, c in ev[r]:
        if lo - 3000 <= c < hi:
            rows.append((c - lo, nm, i))
names = {0: "tile start", 1: "X stored", 30: "O ready", 31: "dO stored", 90: "tile end", 100: "mma: X ready", 140: "mma: dO ready"}
def name(i):
    if i in names: return names[i]
    tab2 = [(10, "fwd zf_rdy got s="), (20, "fwd done s="), (40, "bwd zb_rdy got s="), (50, "bwd loaded s="), (60, "bwd computed s="),
            (70, "bwd gdone ok s="), (80, "bwd stored s="), (120, "mma: zf_used got s="), (130, "mma: gemm2+gemm1 issued s="),
            (150, "mma: zb_used got s="), (160, "mma: g13 issued s="), (170, "mma: tile_rdy got s="), (180, "mma: grad issued s=")]
    tab3 = [(10, "fwd zf_rdy got s="), (20, "fwd done s="), (40, "bwd zb_rdy got j="), (50, "bwd loaded j="), (60, "bwd done j="),
            (120, "mma: zf_used got s="), (130, "mma: gemm2+gemm1 issued s="), (150, "mma: zb_done got j="),
            (160, "mma: grads issued j="), (170, "mma: refill issued j=")]
    for b, n in (tab3 if GEN == "3" else tab2):
        if b <= i < b + 10: return n + str(i - b)
    return str(i)
prev = {}
for dt, nm, i in sorted(rows):
    d = dt - prev.get(nm, dt); prev[nm] = dt
    print(f"{dt:8d} (+{d:6d})  {nm}  {name(i)}")
