"""Per-layer timing of bsde_tconv_fwd (forward geometry and data gradient) with CUDA events.
Usage: python tools/bench_tconv.py [--B 128] [--S 8] [--math tc|fp32] [--group 0|1|2]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _lib  # noqa: E402
from bayesde_b200.arch import FxSpec  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--B", type=int, default=128)
p.add_argument("--S", type=int, default=8)
p.add_argument("--math", default="tc")
p.add_argument("--groups", default="0,1,2")
p.add_argument("--reps", type=int, default=10)
p.add_argument("--warm", type=int, default=3)
p.add_argument("--no_wgrad", action="store_true")
p.add_argument("--epi_bias", action="store_true", help="EPI_BIAS for the forward geometry too (the taps-on-N path serves BIAS / EULER / ACCUM)")
p.add_argument("--layers", default="", help="comma list of layer indices (default all)")
p.add_argument("--dact", action="store_true", help="data gradients with the EPI_DACT epilogue (scale * acc * act'(aux)), as the reverse pass runs them")
a = p.parse_args()
L = _lib.lib()
dev = "cuda:0"
math = _lib.MATH[a.math]
shapes = {0: (32, 5), 1: (16, 20), 2: (8, 80)}


def dgrad_of(cs):
    D = _lib.ConvLayer()
    for f_, _t in _lib.ConvLayer._fields_:
        setattr(D, f_, getattr(cs, f_))
    D.cin, D.cout = cs.cout, cs.cin
    D.hin, D.win, D.hout, D.wout = cs.hout, cs.wout, cs.hin, cs.win
    D.sn, D.kn, D.off, D.sd = cs.sd, -cs.kn, -cs.off, cs.sn
    D.has_time, D.time_first, D.b_off = 0, 0, -1
    D.w_k_stride, D.w_n_stride = cs.w_n_stride, cs.w_k_stride
    return D


for g in [int(v) for v in a.groups.split(",")]:
    H, Cc = shapes[g]
    fx = FxSpec(0, (1, H, H, Cc), 64, "softplus")
    w = torch.randn(a.S, fx.P, device=dev) * 0.05
    for li, cs in enumerate(fx.conv_layer_structs()):
        if a.layers and str(li) not in a.layers.split(","):
            continue
        for kind, lay in (("fwd", cs), ("dgrad", dgrad_of(cs))):
            x = torch.randn(a.S, a.B, lay.hin, lay.win, lay.cin, device=dev)
            out = torch.empty(a.S, a.B, lay.hout, lay.wout, lay.cout, device=dev)
            nb = L.bsde_tconv_fwd_workspace_bytes(C.byref(lay), a.S, math)
            ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
            epi = _lib.EPI_BIAS_ACT if (kind == "fwd" and not a.epi_bias) else _lib.EPI_BIAS
            aux = None
            if kind == "dgrad" and a.dact:
                epi = _lib.EPI_DACT
                aux = torch.rand(a.S, a.B, lay.hout, lay.wout, lay.cout, device=dev) + 0.1

            def run():
                st = L.bsde_tconv_fwd(C.byref(lay), a.S, a.B, _lib.ptr(x), _lib.ptr(w), fx.P, 0.3, 0, epi, 1.0, _lib.ptr(aux) if aux is not None else None,
                                      _lib.ptr(out), math, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
                _lib.check(st, "tconv")
            for _ in range(a.warm):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.reps
            pix = lay.hout * lay.wout if lay.sd == 1 else lay.hin * lay.win
            if kind == "dgrad" and cs.sd == 1 and cs.sn == 2:
                pix = cs.hout * cs.wout
            fl = 2.0 * a.S * a.B * pix * lay.ksize ** 2 * (lay.cin + lay.has_time) * lay.cout
            if kind == "fwd" and not a.no_wgrad:
                dy = torch.randn(a.S, a.B, cs.hout, cs.wout, cs.cout, device=dev)
                gw = torch.zeros(a.S, fx.P, device=dev)
                nbw = L.bsde_tconv_wgrad_workspace_bytes(C.byref(cs), a.S, a.B)
                wsw = torch.empty(nbw, dtype=torch.uint8, device=dev)

                def runw():
                    st = L.bsde_tconv_wgrad(C.byref(cs), a.S, a.B, _lib.ptr(x), _lib.ptr(dy), 0.3, 1.0, _lib.ptr(gw), fx.P, math,
                                            C.c_void_p(wsw.data_ptr()), nbw, _lib.stream_ptr())
                    _lib.check(st, "wgrad")
                for _ in range(2):
                    runw()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.reps):
                    runw()
                e1.record()
                torch.cuda.synchronize()
                usw = e0.elapsed_time(e1) * 1e3 / a.reps
                print(f"g{g} L{li} wgrad                                        : {usw:9.1f} us  {fl / usw / 1e6:8.2f} TFLOP/s", flush=True)
            print(f"g{g} L{li} {kind:5s} cin {lay.cin:3d} cout {lay.cout:3d} {lay.hin:2d}->{lay.hout:2d} sd{lay.sd} sn{lay.sn}: "
                  f"{us:9.1f} us  {fl / us / 1e6:8.2f} TFLOP/s", flush=True)
