"""Times the two whole-chip passes of one L-BFGS direction computation (pinn_lbfgs_direction_probe) -- development."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinn_depthestimation_b200 import _cabi

lib = _cabi.lib()
dev = torch.device("cuda:0")
al = lambda x: (x + 255) & ~255
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n_par, m in ((41703, 100), (462852, 100), (462852, 20), (4000000, 50)):
    nb = C.c_size_t(0)
    _cabi.check(lib.pinn_lbfgs_workspace_bytes(n_par, m, C.byref(nb)))
    ws = torch.zeros(nb.value + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    tail = 6 * al(4 * n_par) + 2 * al((m + 1) * n_par * 4)
    ws[off + nb.value - tail: off + nb.value].view(torch.float32).normal_(0.0, 1e-3)
    g = torch.randn(n_par, device=dev) * 1e-3
    byts = C.c_double(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for _ in range(5):
        e0.record()
        _cabi.check(lib.pinn_lbfgs_direction_probe(C.c_void_p(ws.data_ptr() + off), n_par, m, _cabi.ptr(g), C.byref(byts), st))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"P={n_par} m={m}: {best * 1e3:.1f} us, {byts.value / 1e6:.1f} MB algorithmic -> {byts.value / (best * 1e-3) / 1e9:.0f} GB/s")
