"""Print the debug timeline of the last matvec_panel launch (OQ_TIMELINE=<file> python bench.py ...)."""
import sys
import numpy as np
names = {0: "start", 1: "prod_prefetched", 2: "prod_pdlwait", 3: "prod_ep", 20: "first_x_issued", 21: "prod_end", 22: "prol_begin",
         23: "prol_stores", 24: "prol_arrived", 25: "cons_first_x", 26: "cons_end", 27: "epi_end", 28: "published"}
for r in range(16):
    names[4 + r] = f"peer{r}_ok"
for f in sys.argv[1:]:
    a = np.fromfile(f, dtype=np.uint64).reshape(160, 32).astype(np.int64)
    used = a[:, 0] > 0
    t0 = a[used, 0].min()
    print(f, "CTAs", used.sum())
    for slot in sorted(names):
        col = a[used, slot]
        col = col[col > 0]
        if col.size:
            print(f"  {names[slot]:>16s}: min {1e-3*(col.min()-t0):8.2f} us  median {1e-3*(np.median(col)-t0):8.2f}  max {1e-3*(col.max()-t0):8.2f}   (n={col.size})")
