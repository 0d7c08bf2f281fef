"""Bind an importable checkout of bryceag11/QUAN_ultralytics to the B200 ops (INTEGRATION.md levels 1 and 2).

    import quan_ultralytics_b200.install as qi
    qi.install(ultralytics=True, classification=True)       # swap the layer classes
    with torch.device("cuda"):                               # the reference runs a probe forward while building
        model = OBBModel("yolo11n-obb-quan.yaml", ch=3, nc=15)

Nothing here is needed on a box without the reference; it is plumbing, not compute.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

from . import modules as M
from . import quaternion_ops as shim

_SWAP = ("QConv2D", "IQBN", "Conv", "DWConv", "QUpsample", "QuaternionMaxPool", "QER")
_originals = []          # (module, attribute, reference class) of every swap, for uninstall()
_derived = {}            # reference class -> our subclass of it (QAttention, OBB): one subclass per reference class


def install_extension_shim(mixing: str = "A") -> types.ModuleType:
    """Level 1: make `import quaternion_ops` (conv.py:47-60) resolve to the ctypes-backed shim."""
    shim.set_mixing(mixing)
    sys.modules["quaternion_ops"] = shim
    return shim


def _swap(mod, name, new) -> None:
    old = getattr(mod, name)
    if old is not new:
        _originals.append((mod, name, old))
        setattr(mod, name, new)


def uninstall() -> None:
    """Put the reference's own classes back (tests build the swapped and the unswapped graph in one process).  A module that was
    first imported AFTER a swap (`from quaternion.qconv import QConv2D` in models/quaternion_blocks.py) bound our class as its
    "original": every loaded module that still holds one of the installed classes gets the reference class back too."""
    restored = {}
    while _originals:
        mod, name, old = _originals.pop()
        restored[id(getattr(mod, name))] = old
        setattr(mod, name, old)
    _rebind(restored)


_REF_PREFIXES = ("ultralytics", "quaternion", "models")


def _rebind(mapping: dict) -> None:
    """Replace classes by identity (`mapping`: id(class) -> replacement) wherever the reference's loaded modules still hold them:
    module globals bound by `from x import Y` before / after a swap, and DEFAULT ARGUMENTS evaluated at import time
    (`norm_class: nn.Module = IQBN`, classification/models/quaternion_blocks.py:91)."""
    if not mapping:
        return
    for modname, mod in list(sys.modules.items()):
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict) or not modname.split(".")[0] in _REF_PREFIXES:
            continue
        for name, val in list(d.items()):
            if not isinstance(val, type):
                continue
            if id(val) in mapping:
                setattr(mod, name, mapping[id(val)])
                continue
            if getattr(val, "__module__", None) != modname:
                continue
            init = val.__dict__.get("__init__")
            if init is None:
                continue
            if init.__defaults__ and any(id(v) in mapping for v in init.__defaults__):
                init.__defaults__ = tuple(mapping.get(id(v), v) for v in init.__defaults__)
            if init.__kwdefaults__ and any(id(v) in mapping for v in init.__kwdefaults__.values()):
                init.__kwdefaults__ = {k: mapping.get(id(v), v) for k, v in init.__kwdefaults__.items()}


_SUPPORTED_HEADS = ((1, 2), (2, 4), (4, 8), (8, 16))       # (key_dim, head_dim) instantiations of quan_qattention_*


def _qattention_forward(self, x):
    """QAttention.forward (block.py:1511-1546) with the attention arithmetic fused (quan_qattention_fwd/bwd): the qkv / pe / proj
    QConv2D layers are the module's own (already the B200 classes), the split / reshape / matmul / softmax / matmul / reshape between
    them is one kernel.  Head shapes the library was not built for keep the reference's own forward."""
    from . import functional as QF
    from . import ops
    if (self.key_dim, self.head_dim) not in _SUPPORTED_HEADS or not ops.on_device(x):
        return self._reference_forward(x)
    o = QF.qattention(self.qkv(x), self.num_heads, self.key_dim, self.head_dim, self.scale)
    return self.proj(o + self.pe(o))


def _make_qattention(ref_cls):
    """Subclass of the REFERENCE's QAttention: its constructor (and therefore every parameter, buffer and state-dict key, including the
    unused IQLN `norm`, block.py:1506) stays the reference's; only `forward` is ours."""
    if getattr(ref_cls, "_quan_fused", False):
        return ref_cls
    if ref_cls not in _derived:
        _derived[ref_cls] = type("QAttention", (ref_cls,), {"forward": _qattention_forward, "_reference_forward": ref_cls.forward,
                                                            "_quan_fused": True, "__module__": __name__,
                                                            "__doc__": _qattention_forward.__doc__})
    return _derived[ref_cls]


class _TorchProxy(types.ModuleType):
    """Stands in for the name `torch` in the globals of the reference's block modules: every attribute is torch's own except `cat`,
    which is functional.cat (the library's channel concatenation when the operands are quaternion activations in the tensor-core
    layout, torch.cat otherwise).  The reference calls `torch.cat(y, 1)` from C2f / C3k2 / C3 / QSPPF / QC2PSA / Concat
    (block.py:350-352 ..., conv.py Concat) — rebinding the module global is the only hook that leaves its source untouched."""

    def __init__(self):
        super().__init__("torch")
        self.__dict__["_real"] = torch

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    @staticmethod
    def cat(tensors, dim=0, *, out=None):
        from . import functional as QF
        return QF.cat(tensors, dim, out=out)


_torch_proxy = _TorchProxy()
_branch_streams = {}


def _head_streams(device, n):
    key = (device.type, device.index)
    pool = _branch_streams.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _obb_forward(self, x):
    """OBB.forward / Detect.forward in training mode (head.py:137-147, :338-350) with the nine independent branches — box (cv2), class
    (cv3) and angle (cv4) towers of the three pyramid levels, each two or four `Conv` blocks and a QER — issued on nine streams:
    the P4 / P5 towers are latency-bound chains of 5-20 us kernels that fit beside the P3 towers instead of queueing behind them.
    The fork / join is stream events only, so a captured step replays it as parallel graph branches and autograd runs every
    tower's backward on its forward stream.  Same outputs as the reference: ([cat(box_i, cls_i, 1)], angle)."""
    import math
    if (not self.training or self.end2end or not x[0].is_cuda or os.environ.get("QUAN_HEAD_STREAMS", "1") == "0"):
        return self._reference_forward(x)
    from . import functional as QF
    bs = x[0].shape[0]
    main = torch.cuda.current_stream(x[0].device)
    towers = [(i, j, t[i]) for i in range(self.nl) for j, t in enumerate((self.cv2, self.cv3, self.cv4))]
    streams = _head_streams(x[0].device, len(towers))
    # box / class towers stop before their QER when the pair can write the concatenated tensor directly (functional.qer_cat)
    pair = [isinstance(self.cv2[i][-1], M.QER) and isinstance(self.cv3[i][-1], M.QER) and
            os.environ.get("QUAN_QER_CAT", "1") != "0" for i in range(self.nl)]
    outs = {}
    for (i, j, tower), s in zip(towers, streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            y = x[i]
            mods = list(tower)
            for m in (mods[:-1] if (j < 2 and pair[i]) else mods):
                y = m(y)
            if j < 2 and pair[i] and not mods[-1].fused_ok(y):
                y = mods[-1](y)                     # a shape the kernel does not serve: the module's own fallback
                pair[i] = None
            outs[i, j] = y
        x[i].record_stream(s)
    for (i, j, _), s in zip(towers, streams):
        main.wait_stream(s)
        outs[i, j].record_stream(main)
    angle = torch.cat([outs[i, 2].view(bs, self.ne, -1) for i in range(self.nl)], 2)
    angle = (angle.sigmoid() - 0.25) * math.pi
    for i in range(self.nl):
        if pair[i]:
            qa, qb = self.cv2[i][-1].output_proj, self.cv3[i][-1].output_proj
            x[i] = QF.qer_cat(outs[i, 0], qa.weight, qa.bias, outs[i, 1], qb.weight, qb.bias)
        else:
            a, b = outs[i, 0], outs[i, 1]
            if pair[i] is None:                     # one of the two already went through its QER: finish the other
                a = a if a.dim() == 4 else self.cv2[i][-1](a)
                b = b if b.dim() == 4 else self.cv3[i][-1](b)
            x[i] = torch.cat((a, b), 1)
    return x, angle


def _c2f_forward(self, x):
    """C2f.forward (block.py:348-352; C3k2 inherits it) with the chunk's backward and the concatenation on the library's row-copy
    kernels: same tensors, same order of modules."""
    from . import functional as QF
    y = list(QF.chunk2(self.cv1(x)))
    y.extend(m(y[-1]) for m in self.m)
    return self.cv2(QF.cat(y, 1))


def _make_obb(ref_cls):
    """Subclass of the REFERENCE's OBB head: constructor, parameters and every non-training path are the reference's."""
    if getattr(ref_cls, "_quan_streams", False):
        return ref_cls
    if ref_cls in _derived:
        return _derived[ref_cls]
    _derived[ref_cls] = type("OBB", (ref_cls,), {"forward": _obb_forward, "_reference_forward": ref_cls.forward, "_quan_streams": True,
                                    "__module__": __name__, "__doc__": _obb_forward.__doc__})
    return _derived[ref_cls]


def install(ultralytics: bool = True, classification: bool = True) -> dict:
    """Level 2: replace the reference's layer classes with the B200 modules in every namespace that re-exports them.
    Returns {module_name: [swapped names]} for logging."""
    done = {}
    if ultralytics:
        for modname in ("ultralytics.nn.modules.conv", "ultralytics.nn.modules", "ultralytics.nn.modules.block",
                        "ultralytics.nn.modules.head", "ultralytics.nn.tasks"):
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            names = [n for n in _SWAP if hasattr(mod, n)]
            for n in names:
                _swap(mod, n, getattr(M, n))
            if hasattr(mod, "QAttention"):            # block.py:1485-1546: the reference's class with a fused forward
                ref_cls = importlib.import_module("ultralytics.nn.modules.block").QAttention
                _swap(mod, "QAttention", _make_qattention(ref_cls))
                names.append("QAttention")
            if hasattr(mod, "OBB") and isinstance(getattr(mod, "OBB"), type):   # head.py:322-350: branch-parallel training forward
                ref_cls = importlib.import_module("ultralytics.nn.modules.head").OBB
                _swap(mod, "OBB", _make_obb(ref_cls))
                names.append("OBB")
            done[modname] = names
            if modname == "ultralytics.nn.modules.block" and hasattr(mod, "C2f") and os.environ.get("QUAN_FAST_CAT", "1") != "0":
                _swap(mod.C2f, "forward", _c2f_forward)         # a method of the reference's class, restored by uninstall()
                names.append("C2f.forward")
            if modname in ("ultralytics.nn.modules.conv", "ultralytics.nn.modules.block") and getattr(mod, "torch", None) is torch \
                    and os.environ.get("QUAN_FAST_CAT", "1") != "0":
                _originals.append((mod, "torch", torch))
                mod.torch = _torch_proxy
                names.append("torch.cat")
    if classification:
        for modname in ("quaternion.qconv", "quaternion", "models.quaternion_blocks", "models.quaternion_models",
                        "models.blocks.quaternion_blocks"):
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            names = []
            if hasattr(mod, "QConv2D"):
                _swap(mod, "QConv2D", M.QConv2D_B)      # classification/quaternion/qconv.py:606-609 mixes with M_B
                names.append("QConv2D")
            if hasattr(mod, "IQBN"):
                _swap(mod, "IQBN", M.IQBN)
                names.append("IQBN")
            if hasattr(mod, "QuaternionMaxPool"):        # models/blocks/quaternion_blocks.py:236-260 (Q-ResNet stems)
                _swap(mod, "QuaternionMaxPool", M.QuaternionMaxPool)
                names.append("QuaternionMaxPool")
            done[modname] = names
    _rebind({id(old): getattr(mod, name) for mod, name, old in _originals})
    return done
