"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU in the build container.

Run here only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures pin oracle/quan_oracle.py and oracle/torch_port.py (tests/test_oracle_golden.py) and are the
ground truth the CUDA path is compared with on the GPU (tests/test_gpu_parity.py).

Reference entry points exercised:
  ultralytics.nn.modules.conv.QConv2D / IQBN / Conv / QUpsample  (M_A; CUDA_EXT forced True so IQBN uses batch
      statistics — SURVEY §0.2; safe on CPU because QConv2D also checks x.is_cuda, conv.py:453)
  classification.quaternion.qconv.QConv2D / IQBN                (M_B)
Backward values come from torch.autograd on those modules.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="yolo_cfg_"))
sys.dont_write_bytecode = True

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "thop"):      # the only missing imports
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "matplotlib":
                m.use = lambda *a, **k: None
                m.rcParams = {}
                m.rc = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "classification"))
    import ultralytics.nn.modules.conv as uconv
    uconv.CUDA_EXT = True
    import quaternion.qconv as cconv
    return uconv, cconv


def t2n(t):
    return t.detach().cpu().numpy()


def conv_case(mod_cls, name, cin, cout, k, s, p, d, g, bias, B, H, W, seed, extra=None):
    torch.manual_seed(seed)
    m = mod_cls(cin * 4, cout * 4, k, stride=s, padding=p, dilation=d, groups=g, bias=bias).double()
    x = torch.randn(B, cin, H, W, 4, dtype=torch.float64, requires_grad=True)
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    rec = dict(x=t2n(x), y=t2n(y), dy=t2n(dy), dx=t2n(x.grad),
               w_r=t2n(m.weight_r), w_i=t2n(m.weight_i), w_j=t2n(m.weight_j), w_k=t2n(m.weight_k),
               dw_r=t2n(m.weight_r.grad), dw_i=t2n(m.weight_i.grad), dw_j=t2n(m.weight_j.grad),
               dw_k=t2n(m.weight_k.grad),
               conf=np.array([cin, cout, k, s, p, d, g, int(bias)], dtype=np.int64))
    if bias:
        rec["bias_r"] = t2n(m.bias_r)
        rec["db_r"] = t2n(m.bias_r.grad)
    return {f"{name}/{k_}": v for k_, v in rec.items()}


def main():
    uconv, cconv = import_reference()
    torch.set_num_threads(4)
    data = {}
    # ---- QConv2D, both mixing matrices ------------------------------------------------------------------------
    cases = [  # name, cin_q, cout_q, k, s, p, d, g, bias, B, H, W
        ("k3s1", 4, 8, 3, 1, 1, 1, 1, False, 2, 6, 7),
        ("k3s2", 4, 4, 3, 2, 1, 1, 1, True, 2, 9, 8),
        ("k1", 8, 4, 1, 1, 0, 1, 1, False, 2, 5, 5),
        ("dw", 4, 4, 3, 1, 1, 1, 4, False, 1, 6, 6),
        ("g2d2", 4, 8, 3, 1, 2, 2, 2, True, 1, 7, 7),
        ("k7s2", 2, 4, 7, 2, 3, 1, 1, True, 1, 12, 12),
    ]
    for i, c in enumerate(cases):
        data.update(conv_case(uconv.QConv2D, "convA_" + c[0], *c[1:], seed=100 + i))
        data.update(conv_case(cconv.QConv2D, "convB_" + c[0], *c[1:], seed=200 + i))

    # ---- first layer: RGB -> Poincare -> conv ------------------------------------------------------------------
    torch.manual_seed(7)
    m = uconv.QConv2D(3, 16, 3, stride=2, padding=1, bias=False).double()
    rgb = torch.rand(2, 3, 8, 8, dtype=torch.float64, requires_grad=True)
    q = m._rgb_to_quaternion(rgb)
    gq = torch.randn_like(q)
    (grgb,) = torch.autograd.grad(q, rgb, gq, retain_graph=True)
    y = m(rgb)
    data.update({"poincare/rgb": t2n(rgb), "poincare/q": t2n(q), "poincare/gq": t2n(gq), "poincare/grgb": t2n(grgb),
                 "poincare/y": t2n(y), "poincare/w_r": t2n(m.weight_r), "poincare/w_i": t2n(m.weight_i),
                 "poincare/w_j": t2n(m.weight_j), "poincare/w_k": t2n(m.weight_k)})
    torch.manual_seed(8)
    rgbn = torch.randn(2, 3, 5, 6, dtype=torch.float64)           # CIFAR/ImageNet style mean/std-normalised input
    data.update({"poincare_n/rgb": t2n(rgbn), "poincare_n/q": t2n(cconv.QConv2D(3, 8, 3)._rgb_to_quaternion(rgbn))})

    # ---- IQBN train / eval (both packages are the same math; record both) --------------------------------------
    for tag, cls in (("iqbnA", uconv.IQBN), ("iqbnB", cconv.IQBN)):
        torch.manual_seed(11)
        bn = cls(6 * 4).double()
        with torch.no_grad():
            bn.gamma.copy_(torch.randn(6, 4).double() * 0.5 + 1)
            bn.beta.copy_(torch.randn(6, 4).double() * 0.3)
            bn.running_mean.copy_(torch.randn(6, 4).double() * 0.1)
            bn.running_var.copy_(torch.rand(6, 4).double() + 0.5)
        rm0, rv0 = t2n(bn.running_mean).copy(), t2n(bn.running_var).copy()
        x = (torch.randn(3, 6, 5, 4, 4, dtype=torch.float64) * 1.7 + 0.4).requires_grad_(True)
        bn.train()
        y = bn(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        data.update({f"{tag}/x": t2n(x), f"{tag}/gamma": t2n(bn.gamma), f"{tag}/beta": t2n(bn.beta),
                     f"{tag}/rm0": rm0, f"{tag}/rv0": rv0, f"{tag}/y": t2n(y), f"{tag}/dy": t2n(dy),
                     f"{tag}/dx": t2n(x.grad), f"{tag}/dgamma": t2n(bn.gamma.grad), f"{tag}/dbeta": t2n(bn.beta.grad),
                     f"{tag}/rm1": t2n(bn.running_mean), f"{tag}/rv1": t2n(bn.running_var),
                     f"{tag}/nbt": np.array(int(bn.num_batches_tracked))})
        bn.eval()
        data[f"{tag}/y_eval"] = t2n(bn(x))

    # ---- Conv block = SiLU(IQBN(QConv2D)) (conv.py:788-809), fwd + bwd -------------------------------------------
    torch.manual_seed(21)
    blk = uconv.Conv(16, 32, 3, 1).double()
    blk.train()
    x = torch.randn(2, 4, 6, 6, 4, dtype=torch.float64, requires_grad=True)
    y = blk(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    data.update({"block/x": t2n(x), "block/y": t2n(y), "block/dy": t2n(dy), "block/dx": t2n(x.grad),
                 "block/w_r": t2n(blk.conv.weight_r), "block/w_i": t2n(blk.conv.weight_i),
                 "block/w_j": t2n(blk.conv.weight_j), "block/w_k": t2n(blk.conv.weight_k),
                 "block/dw_r": t2n(blk.conv.weight_r.grad), "block/dw_k": t2n(blk.conv.weight_k.grad),
                 "block/dgamma": t2n(blk.bn.gamma.grad), "block/dbeta": t2n(blk.bn.beta.grad),
                 "block/rm1": t2n(blk.bn.running_mean), "block/rv1": t2n(blk.bn.running_var)})

    # ---- QUpsample ----------------------------------------------------------------------------------------------------
    torch.manual_seed(31)
    up = uconv.QUpsample(2, "nearest")
    x = torch.randn(2, 3, 4, 5, 4, dtype=torch.float64, requires_grad=True)
    y = up(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    data.update({"upsample/x": t2n(x), "upsample/y": t2n(y), "upsample/dy": t2n(dy), "upsample/dx": t2n(x.grad)})

    np.savez_compressed(OUT / "quan_layers.npz", **data)
    print("wrote", OUT / "quan_layers.npz", f"{(OUT / 'quan_layers.npz').stat().st_size / 1024:.1f} KiB,",
          len(data), "arrays")


if __name__ == "__main__":
    main()
