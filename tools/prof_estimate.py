"""kc_estimate_state at dataset scale: python tools/prof_estimate.py [B] [T] [N]
Prints kernel time (CUDA events, L2 flushed between iterations), algorithmic GB/s ((7 + 4/N + 25) values per node-step)
against the measured HBM copy peak."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200")); 
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10
Pn = CosseratRod(); Pn.N = N; Pn.compute_intermediate_terms()
P = _kc.rod_params(Pn)
def measurements(P, B, T, seed):
    """Smooth synthetic measurements (non-unit quaternions) and random tensions (same family as tests/test_gpu_estimate.py)."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0, P.L, P.N)[None, None, :]
    t = (np.arange(T) * P.del_t)[None, :, None]
    w = 2 * np.pi / (rng.uniform(15, 40, (B, 1, 1)) * P.del_t)
    ph = rng.uniform(0, 2 * np.pi, (B, 1, 1))
    p = np.stack([0.3 * np.sin(w * t + ph) * s ** 2, 0.2 * np.cos(1.3 * w * t) * s ** 2, s + 0 * t + 0 * w], 2)
    h = np.stack([1 + 0 * s + 0 * t + 0 * w, 0.8 * s * np.sin(w * t), 0.6 * s * np.cos(0.7 * w * t + ph),
                  0.3 * s * np.sin(0.4 * w * t + 1)], 2)
    return np.concatenate([p, h], 2), 5 + 5 * rng.random((B, T, 4))
data1, ctl1 = measurements(Pn, 1, T, seed=0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dt in (torch.float64, torch.float32):
    data = torch.tensor(data1, dtype=dt, device="cuda").expand(B, T, 7, N).contiguous()
    ctl = torch.tensor(ctl1, dtype=dt, device="cuda").expand(B, T, 4).contiguous()
    for _ in range(3): est = _ops.estimate_state(P, Pn.L, Pn.del_t, data, ctl)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); est = _ops.estimate_state(P, Pn.L, Pn.del_t, data, ctl); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nbytes = B * T * (32 * N + 4) * data.element_size()
    print("%s B %d T %d N %d: %.3f ms  %.3e rod-node-steps/s  %.0f GB/s algorithmic (HBM copy peak %s)" % (
        str(dt).split(".")[-1], B, T, N, ms, B * T * N / ms * 1e3, nbytes / ms / 1e6, peak.get("hbm_gbs", "?")))
