import numpy as np, time, os, ctypes
from concurrent.futures import ThreadPoolExecutor
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| defrag:", open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip())
print("cpus:", len(os.sched_getaffinity(0)))
n = 348_000_000
src = np.ones(n // 8)   # 348 MB source reused
def fill(out, workers):
    pool = ThreadPoolExecutor(workers)
    step = src.shape[0]
    t0 = time.perf_counter()
    futs = []
    for lo in range(0, n, step):
        hi = min(lo + step, n)
        parts = workers; st = (hi - lo + parts - 1) // parts
        futs += [pool.submit(np.copyto, out[lo + q * st:min(lo + (q + 1) * st, hi)], src[q * st:min((q + 1) * st, hi - lo)]) for q in range(parts)]
    for f in futs: f.result()
    return time.perf_counter() - t0
for w in (4, 8, 16, 32):
    out = np.empty(n)
    t = fill(out, w)
    t2 = fill(out, w)
    print("workers %2d: first touch %.1f ms (%.1f GB/s), second pass %.1f ms (%.1f GB/s)" % (w, t * 1e3, n * 8 / t / 1e9, t2 * 1e3, n * 8 / t2 / 1e9))
    del out
