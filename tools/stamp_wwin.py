"""Phase timestamps (clock64, CTA 0) of wgrad_win_kernel per stage: producer thread 0 and the MMA lane."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _lib
from bayesde_b200.arch import FxSpec
L = _lib.lib(); dev = "cuda:0"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
L.bsde_debug_set_ww_stamps.argtypes = [C.c_void_p]
assert L.bsde_debug_set_ww_stamps(C.c_void_p(buf.data_ptr())) == 0
S, B = 8, 128
fx = FxSpec(0, (1, 32, 32, 5), 64, "softplus")
for li in (0, 2, 3):
    lay = fx.conv_layer_structs()[li]
    x = torch.randn(S, B, lay.hin, lay.win, lay.cin, device=dev)
    dy = torch.randn(S, B, lay.hout, lay.wout, lay.cout, device=dev)
    gw = torch.zeros(S, fx.P, device=dev)
    nb = L.bsde_tconv_wgrad_workspace_bytes(C.byref(lay), S, B)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    for it in range(2):
        buf.zero_()
        st = L.bsde_tconv_wgrad(C.byref(lay), S, B, _lib.ptr(x), _lib.ptr(dy), 0.3, 1.0, _lib.ptr(gw), fx.P, 1, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
        torch.cuda.synchronize()
    t = buf.cpu().tolist()
    n = max(t[5], 1)
    print(f"layer {li}: stages {t[5]}  producer cyc/stage: tables {t[0]/n:.0f} bar {t[1]/n:.0f} wait_empty {t[2]/n:.0f} copies {t[3]/n:.0f} arrive {t[4]/n:.0f}"
          f" | mma: wait_full {t[8]/n:.0f} issue+commit {t[9]/n:.0f}")
