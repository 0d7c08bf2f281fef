"""The BASELINE.json configurations as runnable workloads: the reference's OWN model graphs (built by the reference's own
Python from its own yaml / factory functions, found through refenv) with the B200 layer classes installed — or left
untouched (`swapped=False`: the reference PyTorch path, the parity oracle of BASELINE.json and the `--impl reference` arm) —
plus the synthetic batches of SURVEY §8(d).  Host plumbing only: nothing here computes.

  config[0]  Q-WRN-16-2, 128x3x32x32 ~ N(0,1), CE on the returned logits, SGD(.1, .9, 1e-4, nesterov) + clip 1.0
             (classification/models/quaternion_models.py:80-90, classification.py:202-203, utils/training.py:78)
  config[2]  QUAN-YOLO11n-OBB, nc=15, 16x3x1024x1024 in [0,1], ~40 rotated boxes / image, v8OBBLoss, SGD(nesterov) + clip 10
             (nn/tasks.py:942 parse_model, utils/loss.py:941, engine/trainer.py:586-594)
  config[3]  Q-ResNet-34, 256x3x224x224 ~ N(0,1), CE (quaternion_models.py:173-253)
  config[4]  QUAN-YOLO11s-OBB, batch 8
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import install as _install
from . import refenv


def _device_ctx(device):
    return torch.device(device)


def build_yolo_obb(scale: str = "n", nc: int = 15, device="cuda", swapped: bool = True, verbose: bool = False):
    """`OBBModel('yolo11{scale}-obb-quan.yaml', ch=3, nc=nc)` exactly as the reference builds it (the scale letter is parsed from the
    file name, nn/tasks.py:1109-1132), `model.args = get_cfg(DEFAULT_CFG)` for the loss gains.  swapped=False: the untouched
    reference with its IQBN on the batch-statistics branch (see reference_batch_stat_iqbn)."""
    refenv.activate()
    import ultralytics.nn.modules.conv as uconv
    if swapped:
        _install.install(ultralytics=True, classification=False)
    else:
        _install.uninstall()
    from ultralytics.cfg import get_cfg
    from ultralytics.nn.tasks import OBBModel
    from ultralytics.utils import DEFAULT_CFG
    with _device_ctx(device):        # the reference runs a 1x3x256x256 probe forward while building (nn/tasks.py:341)
        model = OBBModel(f"yolo11{scale}-obb-quan.yaml", ch=3, nc=nc, verbose=verbose)
    model.args = get_cfg(DEFAULT_CFG)
    if not swapped:
        reference_batch_stat_iqbn(model)
    model = model.to(device)
    if swapped:
        from . import modules
        modules.pool_batch_counters(model)
    return model


def reference_batch_stat_iqbn(model) -> None:
    """The untouched reference's IQBN only takes its batch-statistics branch when the module global `CUDA_EXT` is true
    (conv.py:537, a defect — SURVEY §0.2), but the same global would send QConv2D into the compiled extension (conv.py:453).
    The PyTorch path BASELINE.json names as the oracle = PyTorch QConv2D + batch-statistics IQBN: hooks flip the global around
    every training-mode IQBN.forward and nothing else; no reference code is modified."""
    import ultralytics.nn.modules.conv as uconv

    def pre(mod, args):
        if mod.training:
            uconv.CUDA_EXT = True

    def post(mod, args, out):
        uconv.CUDA_EXT = False

    for m in model.modules():
        if type(m) is uconv.IQBN or type(m).__name__ == "IQBN" and type(m).__module__ == uconv.__name__:
            m.register_forward_pre_hook(pre)
            m.register_forward_hook(post)


def synthetic_obb_batch(B: int, size: int, device="cpu", nc: int = 15, boxes_per_image: int = 40, seed: int = 1) -> Dict[str, torch.Tensor]:
    """img ~ U[0,1); per image `boxes_per_image` rotated boxes xy ~ U(.1,.9), wh ~ U(.02,.12),
    theta ~ U(-pi/4, 3pi/4); format of data/dataset.py:232-248 (what utils/loss.py:961-966 reads)."""
    g = torch.Generator().manual_seed(seed)
    n = B * boxes_per_image
    img = torch.rand(B, 3, size, size, generator=g)
    batch_idx = torch.arange(B).repeat_interleave(boxes_per_image).float()
    cls = torch.randint(0, nc, (n, 1), generator=g).float()
    xy = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(n, 2, generator=g) * 0.10 + 0.02
    th = torch.rand(n, 1, generator=g) * math.pi - math.pi / 4
    batch = {"img": img, "batch_idx": batch_idx, "cls": cls, "bboxes": torch.cat([xy, wh, th], 1)}
    return {k: v.to(device) for k, v in batch.items()}


def build_classifier(name: str = "qwrn16_2", num_classes: Optional[int] = None, device="cuda", swapped: bool = True):
    """`create_qwrn_16_2(num_classes=10, mapping_type='poincare')` / `create_qrn34_imagenet(1000)` from the reference's
    classification/models/quaternion_models.py, with the B200 classes installed in the classification namespaces."""
    refenv.activate()
    if swapped:
        _install.install(ultralytics=False, classification=True)
    else:
        _install.uninstall()
    import models.quaternion_models as qm
    with _device_ctx(device):
        if name == "qwrn16_2":
            model = qm.create_qwrn_16_2(num_classes=num_classes or 10, mapping_type="poincare")
        elif name == "qresnet34":
            model = qm.create_qrn34_imagenet(num_classes or 1000)
        else:
            raise ValueError(name)
    model = model.to(device)
    if swapped:
        from . import modules
        modules.pool_batch_counters(model)
    return model


def synthetic_classification_batch(B: int, size: int, num_classes: int, device="cpu", seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, size, size, generator=g).to(device), torch.randint(0, num_classes, (B,), generator=g).to(device)


def yolo_param_groups(model):
    """The reference's parameter groups (engine/trainer.py:766-798 build_optimizer): (decayed weights, normalisation weights,
    biases).  IQBN gamma / beta fall in the DECAYED group because the reference only exempts `nn.*Norm*` classes (SURVEY §8(c)
    defect 5; reproduced, it is the reference's behaviour).  QER registers its bias twice (`bias` and `output_proj.bias`, head.py:38-39):
    the reference hands the duplicate to torch (which warns); here every parameter appears once."""
    import torch.nn as nn
    g = [], [], []
    seen = set()
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    for module_name, module in model.named_modules():
        for param_name, param in module.named_parameters(recurse=False):
            if id(param) in seen:
                continue
            seen.add(id(param))
            fullname = f"{module_name}.{param_name}" if module_name else param_name
            if "bias" in fullname:
                g[2].append(param)
            elif isinstance(module, bn):
                g[1].append(param)
            else:
                g[0].append(param)
    return g


def yolo_sgd(model, lr: float = 0.01, momentum: float = 0.937, decay: float = 5e-4):
    """torch.optim.SGD over the reference's groups (trainer.py:799-806: SGD(nesterov) for the biases, then the two weight groups);
    lr0 / momentum / weight_decay defaults of cfg/default.yaml."""
    g = yolo_param_groups(model)
    opt = torch.optim.SGD(g[2], lr=lr, momentum=momentum, nesterov=True)
    opt.add_param_group({"params": g[0], "weight_decay": decay})
    opt.add_param_group({"params": g[1], "weight_decay": 0.0})
    return opt
