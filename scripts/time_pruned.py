"""Developer aid: full-length vs z-pruned correlation with the compact synthetic tip."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["pyridine", "tma_cu111", "large_slab"]:
    cfg = synth.CONFIGS[name]
    dr = np.array(cfg.cell) / np.array(cfg.shape)
    rhoS, potS, _ = synth.make_sample(cfg, device=dev)
    rhoT, nrhoT = synth.make_tip(cfg, device=dev, variant="compact")
    sup = ops.tip_zsupport(nrhoT)
    z_lo, z_hi = 76, 130
    out = torch.empty(cfg.shape[:2] + (z_hi - z_lo,), dtype=torch.float64, device=dev)
    res = {}
    for label, kw in (("full", {}), ("pruned", {"tip_support": sup}), ("pruned_slab7", {"tip_support": sup, "slab": 7})):
        zl, zh = (z_lo, z_lo + kw.pop("slab")) if "slab" in kw else (z_lo, z_hi)
        o = out[:, :, : zh - zl].contiguous()
        for _ in range(3):
            ops.density_overlap_window(potS, nrhoT, dr, zl, zh, out=o, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.density_overlap_window(potS, nrhoT, dr, zl, zh, out=o, **kw); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[label] = float(np.median(ts))
        if label == "full": ref = o.clone()
        if label == "pruned": err = float((o - ref).abs().max() / ref.abs().max())
    SKIP = {"x_fwd": 1 << 8, "y_fwd": 1 << 9, "z": 1 << 10, "y_inv": 1 << 11, "x_inv": 1 << 12}
    st = {}
    for nm, bit in SKIP.items():
        ts = []
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.density_overlap_window(potS, nrhoT, dr, z_lo, z_hi, out=out, flags=sum(SKIP.values()) & ~bit, tip_support=sup); b.record(); torch.cuda.synchronize()
            if it: ts.append(a.elapsed_time(b))
        st[nm] = round(float(np.mean(ts)), 3)
    print("   pruned stages", st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.tip_zsupport(nrhoT); b.record(); torch.cuda.synchronize()
    print(f"{name}: support {sup} of nz={cfg.shape[2]} | es full {res['full']:.3f} ms, pruned {res['pruned']:.3f} ms "
          f"(err {err:.1e}), pruned 7-plane slab {res['pruned_slab7']:.3f} ms, support scan {a.elapsed_time(b):.3f} ms", flush=True)
