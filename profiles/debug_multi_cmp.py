import numpy as np
a=np.load("gpurun_out/dbg_dev.npz"); h=np.load("gpurun_out/dbg_host.npz")
for key in sorted(a.files):
    x, y = a[key], h[key]
    if x.shape != y.shape: print(key, "shape", x.shape, y.shape); continue
    if not np.array_equal(x, y):
        d = np.abs(x - y); i = int(np.flatnonzero(d.ravel() > 0)[0])
        print(key, "first diff at", i, "of", x.size, "max", d.max())
