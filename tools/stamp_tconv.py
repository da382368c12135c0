import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _lib
from bayesde_b200.arch import FxSpec
L = _lib.lib(); dev = "cuda:0"
buf = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
L.bsde_debug_set_tc_stamps.argtypes = [C.c_void_p]
assert L.bsde_debug_set_tc_stamps(C.c_void_p(buf.data_ptr())) == 0
S, B = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 128
li = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fx = FxSpec(0, (1, 32, 32, 5), 64, "softplus")
lay = fx.conv_layer_structs()[li]
w = torch.randn(S, fx.P, device=dev) * 0.05
x = torch.randn(S, B, lay.hin, lay.win, lay.cin, device=dev)
out = torch.empty(S, B, lay.hout, lay.wout, lay.cout, device=dev)
nb = L.bsde_tconv_fwd_workspace_bytes(C.byref(lay), S, 1)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
for it in range(3):
    buf.zero_()
    st = L.bsde_tconv_fwd(C.byref(lay), S, B, _lib.ptr(x), _lib.ptr(w), fx.P, 0.3, 0, 1, 1.0, None, _lib.ptr(out), 1, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
    torch.cuda.synchronize()
    t = buf.view(64, 8).cpu()
    for b in (0, 1, 5, 40):
        r = t[b]
        print(it, b, "setup", int(r[1]-r[0]), "prod_loop", int(r[2]-r[1]), "mma_done(thread128)", int(r[3]-r[1]), "tmem_full_wait", int(r[4]-r[2]), "epilogue", int(r[5]-r[4]), "dealloc", int(r[6]-r[5]), "total", int(r[6]-r[0]))
