"""KTS on the device: timings for dataset-sized inputs (and the launch sequence ncu profiles)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.utils import kts
rng = np.random.default_rng(0)
for n, m in ((600, 100), (2000, 300), (4000, 500)):
    x = torch.from_numpy(rng.standard_normal((n, 1024)).astype(np.float32)).cuda()
    kts.kts(x, m); torch.cuda.synchronize(); t0 = time.time(); c = kts.kts(x, m); torch.cuda.synchronize()
    print(f"KTS n={n} max_ncp={m}: {1e3*(time.time()-t0):.1f} ms, {len(c)} change points", flush=True)
