"""Host-side trace of ht_1d_moments steps on the bench workload: wall time of every C-ABI call and the gaps between
them, printed for the slowest steps (the device stage times are constant, so a slow step is a host stall)."""
import os, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scrna-parameter-estimation_b200"))
import torch
import memento_b200 as memento
from memento_b200 import synth, _lib
ad = synth.make_counts_fast(25000, 10000, n_conditions=2, n_types=8, seed=7, device="cuda")
memento.setup_memento(ad, "q", profile=True); memento.create_groups(ad, ["stim", "cell"]); memento.compute_1d_moments(ad)
cov, tr = synth.design_from_groups(ad.uns["memento"]["groups"], ["stim", "cell"])
trace = []
orig = _lib.call
def traced(name, *a, **k):
    t0 = time.perf_counter(); r = orig(name, *a, **k); trace.append((name, t0, time.perf_counter())); return r
_lib.call = traced
import memento_b200.engine as E, memento_b200.gev as Gv, memento_b200.main as M
for mod in (E, Gv, M):
    if hasattr(mod, "_lib"): mod._lib.call = traced
_empty = torch.empty
def t_empty(*a, **k):
    t0 = time.perf_counter(); r = _empty(*a, **k); t1 = time.perf_counter()
    if t1 - t0 > 1e-3: trace.append(("torch.empty %s (%.0f MB)" % (k.get("dtype"), r.numel() * r.element_size() / 1e6), t0, t1))
    return r
torch.empty = t_empty
_mgi = torch.cuda.mem_get_info
def t_mgi(*a, **k):
    t0 = time.perf_counter(); r = _mgi(*a, **k); trace.append(("mem_get_info", t0, time.perf_counter())); return r
torch.cuda.mem_get_info = t_mgi
for _ in range(3):
    memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=1)
gc.collect(); gc.disable()
steps = []
E2E = bool(os.environ.get("MM_DIAG_E2E"))      # host-staged mode: the matrix is uploaded inside every call
st = ad.uns["memento"]["_b200"]
for i in range(40):
    if E2E:
        st.offload()
    trace.clear(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    memento.ht_1d_moments(ad, cov, tr, num_boot=10000, resampling="bootstrap", seed=10 + i)
    t1 = time.perf_counter()
    steps.append((t1 - t0, [(n, a - t0, b - t0) for n, a, b in trace]))
ts = sorted(s[0] for s in steps)
print("steps ms: min %.1f median %.1f max %.1f" % (ts[0] * 1e3, ts[len(ts) // 2] * 1e3, ts[-1] * 1e3))
for label, (dt, tr_) in (("FASTEST", min(steps, key=lambda s: s[0])), ("SLOWEST", max(steps, key=lambda s: s[0]))):
    print(label, "%.1f ms" % (dt * 1e3))
    prev = 0.0
    for n, a, b in tr_:
        print("   +%7.2f gap %6.2f  call %6.2f  %s" % (a * 1e3, (a - prev) * 1e3, (b - a) * 1e3, n)); prev = b
    print("   end gap %.2f" % ((dt - prev) * 1e3))
