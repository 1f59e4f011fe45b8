import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
for k in a.files:
    x, y = a[k], b[k]
    print(k, "bit-identical" if np.array_equal(x, y) else f"DIFF max abs {np.abs(x - y).max():.3e} rel {np.abs(x - y).max() / (np.abs(x).max() + 1e-30):.3e}")
