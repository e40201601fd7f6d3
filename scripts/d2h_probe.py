"""Device -> pageable host copy of a large result column set (what `DistContext.all_gather_rows` does on the consumer):
torch's `.cpu()` per column against a chunked copy through two small pinned staging buffers drained by host threads.
   python scripts/d2h_probe.py [total_MB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from liana_py_b200.method._dist import d2h_chunked

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
dev = torch.device("cuda", 0)
cols = [torch.randint(0, 255, (mb * (1 << 20) // 13,), dtype=torch.uint8, device=dev) for _ in range(13)]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); a = [c.cpu().numpy() for c in cols]; t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); b = d2h_chunked(cols); t2 = time.perf_counter() - t0
    ok = all(np.array_equal(x, y) for x, y in zip(a, b))
    print(f"{mb} MB in 13 columns: .cpu() {t1 * 1e3:.0f} ms ({mb / 1e3 / t1:.1f} GB/s), chunked pinned {t2 * 1e3:.0f} ms ({mb / 1e3 / t2:.1f} GB/s), equal={ok}", flush=True)
    del a, b
