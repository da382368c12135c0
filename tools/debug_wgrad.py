import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _lib
from bayesde_b200.arch import FxSpec
from oracle import jax_path as jp
L = _lib.lib(); DEV = "cuda:0"
S, B, t = 2, 3, 0.61
for (C_in, fx_dim, H) in [(5, 64, 8), (20, 64, 16), (80, 64, 8)]:
    fx = FxSpec(0, (1, H, H, C_in), fx_dim, "softplus")
    layout, P = jp.fx_layout(0, C_in, fx_dim)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(S, P, generator=g, dtype=torch.float64) * 0.2
    convs = [e for e in layout if e["kind"] == "conv"]
    for li, (cs, e) in enumerate(zip(fx.conv_layer_structs(), convs)):
        x = torch.randn(S, B, cs.hin, cs.win, cs.cin, generator=g, dtype=torch.float64)
        n = e["k"] ** 2 * (e["cin"] + 1) * e["cout"]
        dy = torch.randn(S, B, cs.hout, cs.wout, cs.cout, generator=g, dtype=torch.float64)
        gref = []
        for s_ in range(S):
            ws_ = w[s_].clone().requires_grad_(True)
            o_ = jp.concat_conv2d(x[s_], t, ws_[e["w_off"]:e["w_off"] + n].reshape(e["w_shape"]), ws_[e["b_off"]:e["b_off"] + e["cout"]], e["stride"], e["transpose"])
            (g_,) = torch.autograd.grad((o_ * dy[s_]).sum(), ws_)
            gref.append(g_)
        gref = torch.stack(gref)
        xd = x.float().to(DEV).contiguous(); dyd = dy.float().to(DEV).contiguous()
        for math in (0, 1):
            gwd = torch.zeros(S, P, device=DEV)
            nb = L.bsde_tconv_wgrad_workspace_bytes(C.byref(cs), S, B)
            ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
            st = L.bsde_tconv_wgrad(C.byref(cs), S, B, _lib.ptr(xd), _lib.ptr(dyd), t, 1.0, _lib.ptr(gwd), P, math, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
            _lib.check(st, "wgrad"); torch.cuda.synchronize()
            lo, hi = e["w_off"], e["b_off"] + e["cout"]
            a = gwd[:, lo:hi].double().cpu(); b = gref[:, lo:hi]
            W = (a[:, :n] - b[:, :n]).reshape(S, 9, e["cin"] + 1, e["cout"]).abs()
            print(f"C{C_in} L{li} math{math} cin{cs.cin} cout{cs.cout} sd{cs.sd}: max err {float((a-b).abs().max()):.3e} ref {float(b.abs().max()):.3e} | per-tap err {[round(float(W[:, k, :-1].max()),3) for k in range(9)]} time-row err {float(W[:, :, -1].max()):.2e}")
