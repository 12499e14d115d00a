# minimal CPU stand-in for the torch calls bench.py makes (control-flow test under the SIMT emulator only)
import numpy as np
uint8 = np.uint8
float64 = np.float64
class _T:
    def __init__(self, a): self.a = a
    def fill_(self, v): self.a[...] = v; return self
    def pin_memory(self): return self
    def data_ptr(self): return self.a.ctypes.data
    def __getitem__(self, i): return self.a[i]
    def tolist(self): return self.a.tolist()
def empty(n, dtype=None, device=None): return _T(np.empty(min(n, 1 << 20), dtype=np.uint8))
def from_numpy(a): return _T(a)
def tensor(x, dtype=None, device=None): return _T(np.array(x))
class device:
    def __init__(self, *a): pass
class cuda:
    @staticmethod
    def is_available(): return True
    @staticmethod
    def set_device(d): pass
    @staticmethod
    def synchronize(): pass
