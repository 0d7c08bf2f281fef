"""Record a layer trace of the REAL reference's Q-WRN-16-2 (BASELINE config[0]) training step on CPU:
    python tests/golden/make_qwrn_trace.py          # build container only (needs /root/reference)
create_qwrn_16_2(num_classes=10) (classification/models/quaternion_models.py:80-90), seed 0, x = randn(2,3,32,32),
cross-entropy on the returned quaternion-norm logits, one backward.  Forward / backward hooks capture, for a handful of
layers that span the model (RGB first layer with the Poincare map, a stride-2 conv, a deep conv, two IQBNs), the
input, output, gradient w.r.t. output and input, and the parameter gradients.  tests/test_gpu_parity.py replays those
layers through the CUDA path (classification flavour: mixing matrix M_B, bias on every conv).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import import_reference, t2n  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    _, cconv = import_reference()
    from models.quaternion_models import create_qwrn_16_2
    torch.manual_seed(0)
    torch.set_num_threads(4)
    net = create_qwrn_16_2(num_classes=10, mapping_type="poincare").double().train()
    x = torch.randn(2, 3, 32, 32, dtype=torch.float64)
    labels = torch.tensor([3, 7])
    convs = [(n, m) for n, m in net.named_modules() if isinstance(m, cconv.QConv2D)]
    bns = [(n, m) for n, m in net.named_modules() if isinstance(m, cconv.IQBN)]
    strided = [nm for nm in convs if tuple(nm[1].stride) == (2, 2) and nm[1].kernel_size[0] == 3]
    picks = {"conv_first": convs[0], "conv_s2": strided[0], "conv_deep": convs[-1], "bn_first": bns[0], "bn_last": bns[-1]}
    rec = {}

    def hook_fwd(tag):
        def f(mod, inp, out):
            rec[f"{tag}/x"] = t2n(inp[0])
            rec[f"{tag}/y"] = t2n(out)
        return f

    def hook_bwd(tag):
        # module-level hook: grad_input is THIS layer's contribution only (the block input also feeds the shortcut branch)
        def f(mod, gin, gout):
            rec[f"{tag}/dy"] = t2n(gout[0])
            if gin[0] is not None:
                rec[f"{tag}/dx"] = t2n(gin[0])
        return f

    for tag, (name, mod) in picks.items():
        mod.register_forward_hook(hook_fwd(tag))
        mod.register_full_backward_hook(hook_bwd(tag))
        rec[f"{tag}/name"] = np.array(name)
    out = net(x)
    loss = torch.nn.functional.cross_entropy(out, labels)
    loss.backward()
    rec["loss"] = t2n(loss)
    for tag, (name, mod) in picks.items():
        if isinstance(mod, cconv.QConv2D):
            rec[f"{tag}/conf"] = np.array([mod.kernel_size[0], mod.stride[0], mod.padding[0], mod.dilation[0], mod.groups,
                                           int(mod.bias_r is not None)], dtype=np.int64)
            for c in "rijk":
                w = getattr(mod, f"weight_{c}")
                rec[f"{tag}/w_{c}"] = t2n(w)
                rec[f"{tag}/dw_{c}"] = t2n(w.grad)
            if mod.bias_r is not None:
                rec[f"{tag}/bias_r"] = t2n(mod.bias_r)
                rec[f"{tag}/db_r"] = t2n(mod.bias_r.grad)
        else:
            for k in ("gamma", "beta"):
                rec[f"{tag}/{k}"] = t2n(getattr(mod, k))
                rec[f"{tag}/d{k}"] = t2n(getattr(mod, k).grad)
            rec[f"{tag}/eps"] = np.array(mod.eps)
    data = {k: (v.astype(np.float32) if v.dtype == np.float64 and v.ndim > 0 else v) for k, v in rec.items()}
    np.savez_compressed(OUT / "qwrn_trace.npz", **data)
    tot = sum(v.nbytes for v in data.values())
    print(f"wrote {OUT / 'qwrn_trace.npz'}: {len(data)} arrays, {tot / 1e6:.2f} MB raw; loss {float(loss):.6f}")
    for tag, (name, mod) in picks.items():
        print(tag, name, {k.split('/')[1]: v.shape for k, v in data.items() if k.startswith(tag + '/') and getattr(v, 'ndim', 0) > 1})


if __name__ == "__main__":
    main()
