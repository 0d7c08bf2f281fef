"""Probe the REAL reference models for their QConv2D call histogram (build container only: needs /root/reference):
    python tools/probe_model_trace.py            # writes tests/golden/model_traces.json
Forward hooks on every QConv2D / IQBN of OBBModel(yolo11{n,s}-obb-quan.yaml, nc=15) (ultralytics/nn/tasks.py) at 256^2
(spatial sizes scaled x4 to 1024^2) and of create_qrn34_imagenet(1000) (classification/models/quaternion_models.py) at
224^2 record, per call: C_i, C_o (quaternion channels), k, stride, groups, H_o, whether an IQBN consumes the output
directly, whether the conv has a bias, and the call count.  bench.py replays these shapes (`--workload yolo11s_trace`,
`qresnet34_trace`); SURVEY §8(a) quotes the same histograms (4.21 / 14.34 / 1.85 separable GFLOP fwd per image).
"""
import collections
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from make_golden import import_reference  # noqa: E402


def _one(v):
    return v[0] if isinstance(v, (tuple, list)) else v


def trace(model, conv_cls, bn_cls, x, scale):
    events, keep = [], []

    def conv_hook(mod, inp, out):
        keep.append(out)
        events.append(["conv", out, [mod.in_channels_per_comp, mod.out_channels_per_comp, _one(mod.kernel_size), _one(mod.stride),
                                     mod.groups, out.shape[2] * scale, False, getattr(mod, "bias_r", None) is not None]])

    def bn_hook(mod, inp, out):
        # the conv whose output this IQBN normalises: the very tensor, else (a dropout / identity in between, as in the
        # classification blocks) the most recent not yet matched conv output of the same shape
        for same in (True, False):
            for e in reversed(events):
                if not e[2][6] and (e[1] is inp[0] if same else e[1].shape == inp[0].shape):
                    e[2][6] = True
                    return

    hs = [m.register_forward_hook(conv_hook) for m in model.modules() if isinstance(m, conv_cls)]
    hs += [m.register_forward_hook(bn_hook) for m in model.modules() if isinstance(m, bn_cls)]
    with torch.no_grad():
        model(x)
    for h in hs:
        h.remove()
    hist = collections.OrderedDict()
    for _, _, row in events:
        hist[tuple(row)] = hist.get(tuple(row), 0) + 1
    rows = [list(k) + [c] for k, c in hist.items()]
    gflop = sum(c * 4 * 2 * ho * ho * co * (ci // g) * k * k for ci, co, k, s, g, ho, bn, b, c in rows) / 1e9
    melem = sum(c * co * ho * ho * 4 for ci, co, k, s, g, ho, bn, b, c in rows if bn) / 1e6
    return {"columns": ["C_i", "C_o", "k", "stride", "groups", "H_o", "iqbn", "bias", "count"], "rows": rows,
            "calls": sum(r[-1] for r in rows), "gflop_fwd_per_image": gflop, "iqbn_melems_per_image": melem}


def main():
    uconv, cconv = import_reference()
    torch.set_num_threads(8)
    from ultralytics.nn.tasks import OBBModel
    from models.quaternion_models import create_qrn34_imagenet
    out = {}
    for name in ("yolo11n", "yolo11s"):
        m = OBBModel(f"/root/reference/ultralytics/cfg/models/11/{name}-obb-quan.yaml", ch=3, nc=15, verbose=False).train()
        out[name] = dict(trace(m, uconv.QConv2D, uconv.IQBN, torch.rand(2, 3, 256, 256), 4), image=1024, mix="A")
    m = create_qrn34_imagenet(1000).train()
    out["qresnet34"] = dict(trace(m, cconv.QConv2D, cconv.IQBN, torch.randn(2, 3, 224, 224), 1), image=224, mix="B")
    (ROOT / "tests" / "golden" / "model_traces.json").write_text(json.dumps(out, indent=None, separators=(",", ":")) + "\n")
    for k, v in out.items():
        print(k, v["calls"], "calls", round(v["gflop_fwd_per_image"], 3), "GFLOP fwd/img", round(v["iqbn_melems_per_image"], 2), "M IQBN elems/img",
              file=sys.stderr)


if __name__ == "__main__":
    main()
