"""Developer probe: pinned H2D / D2H bandwidth of the box through the engine's own memcpy helpers (1D), alone and concurrently."""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vocoderproject_b200 as vp

eng = vp.Engine(48000.0, 1024, 32, 4)
rows, cols = 1024, 1 << 20  # 4 GiB
a = vp.PinnedArray(rows, cols)
b = vp.PinnedArray(rows, cols)
a.array[:] = 1.0
d1 = eng.device_alloc(rows * cols * 4)
d2 = eng.device_alloc(rows * cols * 4)
lib = eng.lib
for name, fn in (("H2D", lambda: lib.vp_memcpy_h2d(eng.h, d1, a.ptr, rows * cols * 4)),
                 ("D2H", lambda: lib.vp_memcpy_d2h(eng.h, b.ptr, d2, rows * cols * 4))):
    fn()
    t = time.time(); fn(); dt = time.time() - t
    print("%s alone: %.1f GB/s" % (name, rows * cols * 4 / dt / 1e9))
