"""Time the scan kernel of whatever build FQB200_LIB points at on the bench workload (1 GiB fixed150) -- used with the
-DFQB_SCAN_LOADS_ONLY build to see what the staging pipeline alone (TMA bulk copies, mbarrier waits, one barrier per
tile, no row scan) delivers: the ceiling of the kernel's memory side with this geometry."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402,F401
import fastqandfurious_b200 as fq  # noqa: E402
from fastqandfurious_b200 import _lib, device  # noqa: E402

L = _lib.lib()
n = (1 << 30) // 337
buf = fq.synth_fixed(n)
table = torch.empty((n + 64, 6), dtype=torch.int64, device='cuda')
result = torch.empty(16, dtype=torch.int64, device='cuda')
for cfg in [int(c) for c in (sys.argv[1:] or ['0'])]:
    flags = _lib.FLAG_CFG(cfg) | _lib.FLAG_FAST_ONLY
    for _ in range(5):
        device.parse_raw(buf, 1, -1, table, None, 0, result, flags)
    torch.cuda.synchronize()
    _lib.check(L.fqb_profile_enable(1), 'profile')
    for _ in range(100):
        device.parse_raw(buf, 1, -1, table, None, 0, result, flags)
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    _lib.check(L.fqb_profile_read(ctypes.byref(tot), ctypes.byref(cnt)), 'read')
    L.fqb_profile_enable(0)
    ms = tot.value / max(1, cnt.value)
    print('lib %s cfg %d scan kernel %.4f ms = %.1f GB/s' % (os.path.basename(_lib.LIBPATH), cfg, ms, buf.numel() / ms / 1e6))
