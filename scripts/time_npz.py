"""Developer aid: np.load + upload vs the pinned reader, one 1 GB array."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200.npzio import read_npz
dev = torch.device("cuda", 0)
d = tempfile.mkdtemp()
a = np.random.rand(512, 512, 512)
f = os.path.join(d, "big.npz")
np.savez(f, rhoS=a)
torch.zeros(1, device=dev); torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); x = np.load(f)["rhoS"]; t1 = time.perf_counter()
    g = torch.from_numpy(x).to(dev); torch.cuda.synchronize(); t2 = time.perf_counter()
    r = read_npz(f, device=dev)["rhoS"]; t3 = time.perf_counter()
    assert torch.equal(r, g)
    print(f"1.07 GB: np.load {t1 - t0:.3f} s + pageable upload {t2 - t1:.3f} s = {t2 - t0:.3f} s | pinned reader + upload {t3 - t2:.3f} s", flush=True)
    del g, r
os.remove(f)
