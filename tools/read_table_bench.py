"""readfastq_table from an io.BytesIO and from a regular file, for several FQB_READ_THREADS (host staging rate)."""
import io
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402,F401
import fastqandfurious_b200 as fq  # noqa: E402

n = (1 << 30) // 337
data = fq.synth_fixed(n).cpu().numpy().tobytes()
tmpdir = '/dev/shm' if os.path.isdir('/dev/shm') else None
with tempfile.NamedTemporaryFile(dir=tmpdir, delete=False) as f:
    f.write(data)
    path = f.name
try:
    for threads in [int(x) for x in (sys.argv[1:] or ['1', '2', '4', '8', '16'])]:
        os.environ['FQB_READ_THREADS'] = str(threads)
        for name, make in (('bytesio', lambda: io.BytesIO(data)), ('file', lambda: open(path, 'rb'))):
            best = None
            for rep in range(3):
                fh = make()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tab = fq.readfastq_table(fh)
                dt = time.perf_counter() - t0
                fh.close()
                assert len(tab) == n and int(tab[-1, 0]) == (n - 1) * 337
                best = dt if best is None else min(best, dt)
            print('threads %2d %-8s %.3f s = %.1f GB/s' % (threads, name, best, len(data) / best / 1e9), flush=True)
finally:
    os.unlink(path)
