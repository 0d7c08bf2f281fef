#!/usr/bin/env python
"""bench.py — hot-path benchmark (contract in the task statement, read per SURVEY §8(d)).

Default workload = BASELINE.json configs[2], the configuration the metric "train images/sec at 1/2/4/8 B200" is quoted on:
the QUAN-YOLO11n-OBB training step (the reference's own model graph from yolo11n-obb-quan.yaml, nc=15, with the B200 layer classes
installed) on a synthetic DOTA-shaped batch of 16 images of 1024^2 per GPU with 40 rotated boxes each, bf16 autocast: forward +
v8OBBLoss + backward + grad-clip 10 + SGD(nesterov) (engine/trainer.py:379-393, :586-594), replayed from two CUDA graphs.
`--impl reference` = the SAME model graph untouched (baseline/_ref, the reference's PyTorch path) on the host cores.
Other workloads: `--workload block_stack` (configs[1] sweep point: a stack of `depth` reference `Conv` blocks, QConv2D 3x3 C_q->C_q ->
IQBN(batch stats) -> SiLU, training step on N x C_q x H x W x 4 activations), `sweep` (the whole configs[1] layer sweep as JSON),
`yolo11s_obb` (configs[4]), `qresnet34` (configs[3]) and the per-layer `*_trace` replays.  metric = train images/s.

  value      : images/s with the batch already resident in HBM (CUDA-event timed, max over ranks)
  e2e        : same step through the public nn.Module API with HOST (pinned) inputs: H2D of the batch and D2H of the
               step's result (first block's weight gradient) inside the timed region
  roofline   : dominant kernel (by share of the step) against MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the reference's PyTorch CPU path (oracle/torch_port.py, all host threads) on a
               bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

_JSON_OUT = None


def json_only_stdout():
    """stdout carries the JSON line(s) only: whatever libraries print there (the reference's import messages, settings banners) is
    sent to stderr at the file-descriptor level, so a driver that parses stdout sees exactly one JSON object per line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="yolo11n_obb", choices=["yolo11n_obb", "yolo11s_obb", "qresnet34", "qwrn16_2", "block_stack", "sweep", "yolo11n_trace", "yolo11s_trace", "qresnet34_trace"],
                    help="block_stack: BASELINE config[1] sweep point (default, the bench line); yolo11n_trace: replay of "
                         "the 87 QConv2D / 84 IQBN+SiLU calls of QUAN-YOLO11n-OBB at 1024^2 (SURVEY 8(a) histogram); "
                         "yolo11s_trace (config[4], default 8 images) and qresnet34_trace (config[3], 224^2, M_B, biased convs, "
                         "default 256 images): the same replay from tests/golden/model_traces.json (tools/probe_model_trace.py)")
    ap.add_argument("--cq", type=int, default=256, help="quaternion channels per component")
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--n", type=int, default=None, help="images per GPU per step (default: 16 yolo11n_obb, 8 yolo11s_obb, 64 block_stack)")
    ap.add_argument("--size", type=int, default=1024, help="yolo*_obb: image side")
    ap.add_argument("--buckets", type=int, default=3, help="yolo*_obb, N > 1: gradient all-reduce buckets launched inside the captured backward")
    ap.add_argument("--no-graph", action="store_true", help="yolo*_obb: eager step (reference v8OBBLoss, host-synchronous) instead of the two CUDA graphs")
    ap.add_argument("--ref-loss", action="store_true", help="yolo*_obb: the reference's own v8OBBLoss (eager, between the two graphs) instead of loss.OBBLossStatic")
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--mix", default="A", choices=["A", "B"])
    ap.add_argument("--sync-iqbn", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=4, help="images per step of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--ddp-bucket-mb", type=int, default=512,
                    help="DDP gradient bucket size.  Default: one bucket (all-reduce once, after the last wgrad): the persistent "
                         "conv kernels own all 148 SMs, so an NCCL kernel that overlaps them delays whole CTA pairs; 25 = "
                         "torch's default overlapped buckets")
    ap.add_argument("--broadcast-buffers", action="store_true",
                    help="DDP: re-broadcast the IQBN running statistics from rank 0 before every forward (torch's default; "
                         "measured 0.3 ms/step at N=2 — off here: the batch-statistics training step does not read them)")
    ap.add_argument("--infer", action="store_true",
                    help="yolo11n_trace: eval-mode forward only under no_grad (IQBN running statistics + SiLU in the conv epilogue); "
                         "reports forward latency per batch of --n images (BASELINE config 5 asks for batch-1 latency)")
    ap.add_argument("--graph", action="store_true", help="yolo11n_trace: capture the step in a CUDA graph (removes host launch overhead)")
    a = ap.parse_args()
    if a.n is None:
        a.n = {"yolo11n_obb": 16, "yolo11s_obb": 8, "qresnet34": 256, "qwrn16_2": 128}.get(a.workload, 64)
        a.n_default = True
    else:
        a.n_default = False
    return a


def workload_name(a):
    return f"conv_block_stack(depth={a.depth},Cq={a.cq},k3s1,{a.hw}x{a.hw},N={a.n}/gpu,mix={a.mix})"


def flops_per_image(a):
    """Separable QConv2D FLOPs, train = 3 x fwd."""
    return 3 * a.depth * 4 * 2 * a.hw * a.hw * a.cq * a.cq * 9


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's PyTorch CPU path (port), all host threads
# ---------------------------------------------------------------------------------------------------------------
def run_cpu_port(a, n_images, steps, warmup):
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[TP.Conv(a.cq, a.cq, 3, 1, mix=a.mix) for _ in range(a.depth)]).train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    x = torch.randn(n_images, a.cq, a.hw, a.hw, 4)
    dy = torch.randn(n_images, a.cq, a.hw, a.hw, 4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        y = net(x)
        y.backward(dy)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_images / sec, sec, cores


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.cpu_n
    ips, sec, cores = run_cpu_port(a, n, a.steps, min(a.warmup, 1))
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "reference PyTorch CPU path (oracle/torch_port.py), fp32"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{n} images/step of the same block stack, {a.steps} timed steps"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def time_op(fn, iters, flush=None):
    """Mean device time (ms) of fn() over `iters` launches, CUDA events on the current stream."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.sum()            # read-only sweep > L2: evicts the op's tensors and leaves CLEAN lines behind
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def reference_cuda_ext(a, dev):
    """The reference's OWN CUDA extension (ultralytics/nn/cuda/quaternion_ops.cu, built by oracle/build_ref_ext.py into
    oracle/_ref/) on a bounded sample of the bench layer, next to our drop-in for the same module API
    (quan_ultralytics_b200/quaternion_ops.py): qconv_forward + qconv_backward of ONE QConv2D (C_q, 3x3, stride 1) in fp32,
    contiguous BCHWQ tensors, mixing matrix M_B (what the extension computes).  A reported baseline; None when
    the prebuilt .so is absent.  Its kernels are scalar (one block per weight element in wgrad), hence the small sample."""
    try:
        from oracle import build_ref_ext
        ref = build_ref_ext.load()
    except Exception as e:                                    # noqa: BLE001 — a baseline must never break the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if ref is None:
        return None
    from quan_ultralytics_b200 import quaternion_ops as ours
    ours.set_mixing("B")
    n, C, H = 2, a.cq, a.hw
    torch.manual_seed(0)
    x = torch.randn(n, C, H, H, 4, device=dev)
    w = [torch.randn(C, C, 3, 3, device=dev) * 0.02 for _ in range(4)]
    dy = torch.randn(n, C, H, H, 4, device=dev)
    args = ([1, 1], [1, 1], [1, 1], 1)

    def run(mod):
        y = mod.qconv_forward(x, *w, None, None, None, None, *args)
        g = mod.qconv_backward(dy, x, *w, False, *args)
        return y, g

    try:
        y_ref, g_ref = run(ref)
        y_our, g_our = run(ours)
        rel = lambda p, q: float((p.double() - q.double()).abs().max() / q.double().abs().max())
        t_ref = time_op(lambda: run(ref), 2)
        t_our = time_op(lambda: run(ours), 10)
        # the same layer the way the nn.Module path runs it: tensor-core layout (BHWQC), fp32 storage / tf32 MMA
        from quan_ultralytics_b200 import ops as qops
        xn, dyn = (t.contiguous(memory_format=torch.channels_last_3d) for t in (x, dy))
        cargs = ((1, 1), (1, 1), (1, 1), 1, qops.M_B)

        def run_native():
            qops.qconv2d_fwd(xn, w, None, *cargs, qops.ALGO_AUTO, qops.LAYOUT_BHWQC)
            qops.qconv2d_bwd(dyn, xn, w, *cargs, True, True, False)

        t_nat = time_op(run_native, 10)
        return {"op": f"QConv2D fwd+bwd (qconv_forward + qconv_backward), Cq={C} 3x3 s1 {H}x{H}, fp32, M_B", "images": n,
                "reference_ms": t_ref, "reference_images_per_sec": n / t_ref * 1e3,
                "ours_dropin_ms": t_our, "ours_dropin_images_per_sec": n / t_our * 1e3,
                "ours_native_ms": t_nat, "ours_native_images_per_sec": n / t_nat * 1e3,
                "max_rel_diff_dropin_vs_reference": {"y": rel(y_our, y_ref), "dx": rel(g_our[0], g_ref[0]),
                                                     "dw_r": rel(g_our[1], g_ref[1])},
                "note": "reference = the reference's own quaternion_ops.cu built for sm_100; dropin = our quaternion_ops shim, same "
                        "contiguous-BCHWQ fp32 contract (large layers: layout conversion in and out around the tcgen05 engine, tf32 "
                        "MMA; set_fast_layout(False) = exact-fp32 CUDA-core engine, 5.3 ms here); native = the nn.Module path's BHWQC layout on "
                        "the tcgen05 engine (tf32 MMA), host launch time included at this 2-image sample"}
    except Exception as e:                                    # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def kernels_in_step(lib, step_fn, steps, a, dtype, peaks):
    """Device time of every library kernel inside the training step itself: the library brackets each launch with a
    CUDA-event pair on its own stream (quan_kernel_timing_*), over `steps` extra steps run right after the timed region
    (same clocks / power state).  Peaks here are the SUSTAINED figures: these kernels run inside a long step."""
    import ctypes
    lib.quan_kernel_timing_enable(1)
    for _ in range(steps):
        step_fn()
    torch.cuda.synchronize()
    lib.quan_kernel_timing_enable(0)
    n = lib.quan_kernel_timing_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.quan_kernel_timing_report(buf, n + 16)
    S = a.n * a.cq * a.hw * a.hw * 4 * (2 if dtype == torch.bfloat16 else 4)
    conv_flops = 4 * 2 * a.n * a.hw * a.hw * a.cq * a.cq * 9
    tpeak = peaks["bf16_tflops_sustained"] * (1.0 if dtype == torch.bfloat16 else 0.5)
    # algorithmic work per launch: separable conv FLOPs; bytes of the streams a kernel must touch
    work = {"qconv_igemm_fwd": ("tensor", conv_flops), "qconv_igemm_dgrad": ("tensor", conv_flops),
            "qconv_wgrad_kernel": ("tensor", conv_flops), "iqbn_reduce_fwd": ("hbm", S), "iqbn_apply_fwd": ("hbm", 2 * S),
            "iqbn_reduce_bwd": ("hbm", 2 * S), "iqbn_apply_bwd": ("hbm", 3 * S), "mix": ("hbm", 2 * S)}
    rows = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, total = ln.split()
        cnt, total = int(cnt), float(total)
        if cnt == 0:
            continue
        r = {"launches_per_step": cnt / steps, "ms_per_launch": total / cnt, "ms_per_step": total / steps}
        if name in work:
            kind, wk = work[name]
            if kind == "tensor":
                ach = wk / (total / cnt) / 1e9
                r.update({"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak})
            else:
                ach = wk / (total / cnt) / 1e6
                r.update({"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]})
        rows[name] = r
    return rows


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the bench-shape kernels, from the committed
    ncu captures (profiles/r01_ncu_traffic.json, r02_ncu_traffic.json); None when a kernel has no capture."""
    out = {}
    for name in ("r01_ncu_traffic.json", "r02_ncu_traffic.json"):
        p = ROOT / "profiles" / name
        if p.exists():
            out.update(json.loads(p.read_text()))
    return out


def kernel_table(a, dev, dtype, peaks):
    """Per-op device times for one block of the stack, each timed alone with the L2 flushed between launches."""
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import ops
    L = ops.LAYOUT_BHWQC
    esz = 2 if dtype == torch.bfloat16 else 4
    N, C, H = a.n, a.cq, a.hw
    x = torch.randn(N, C, H, H, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    dy = torch.randn_like(x)
    w = [torch.randn(C, C, 3, 3, device=dev) * 0.02 for _ in range(4)]
    gamma, beta = torch.ones(C, 4, device=dev), torch.zeros(C, 4, device=dev)
    mix = ops.MIX[a.mix]
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MB
    S = x.numel() * esz
    conv_flops = 4 * 2 * N * H * H * C * C * 9
    cnt = float(N * H * H)
    stats = ops.iqbn_train_stats(x, L, gamma, beta, 1e-5, 0.1, None, None)
    sums = ops.iqbn_bwd_reduce(dy, x, L, stats, gamma, beta, Q.ACT_SILU, cnt)
    rows = {}

    def add(name, fn, kind, work):
        ms = time_op(fn, 10, flush)
        if kind == "tensor":
            peak = peaks["bf16_tflops"] * (1.0 if dtype == torch.bfloat16 else 0.5)
            ach = work / ms / 1e9
            rows[name] = {"ms": ms, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak}
        else:
            ach = work / ms / 1e6
            rows[name] = {"ms": ms, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": ach / peaks["hbm_gbs"]}

    add("qconv2d_fwd", lambda: ops.qconv2d_fwd(x, w, None, (1, 1), (1, 1), (1, 1), 1, mix, ops.ALGO_AUTO, L), "tensor",
        conv_flops)
    add("qconv2d_bwd(dgrad+wgrad+mixT)", lambda: ops.qconv2d_bwd(dy, x, w, (1, 1), (1, 1), (1, 1), 1, mix, True, True, False),
        "tensor", 2 * conv_flops)
    add("iqbn_train_stats", lambda: ops.iqbn_train_stats(x, L, gamma, beta, 1e-5, 0.1, None, None), "hbm", S)
    add("iqbn_apply_fwd_silu", lambda: ops.iqbn_apply_fwd(x, L, stats, gamma, beta, Q.ACT_SILU), "hbm", 2 * S)
    add("iqbn_bwd_reduce", lambda: ops.iqbn_bwd_reduce(dy, x, L, stats, gamma, beta, Q.ACT_SILU, cnt), "hbm", 2 * S)
    add("iqbn_bwd_apply", lambda: ops.iqbn_bwd_apply(dy, x, L, stats, gamma, beta, Q.ACT_SILU, sums, cnt), "hbm", 3 * S)
    return rows


def load_model_trace(name):
    """(rows, meta) of a model's QConv2D call histogram: rows = (C_i, C_o, k, s, groups, H_o, iqbn, bias, count), recorded
    from the real reference by tools/probe_model_trace.py and committed as tests/golden/model_traces.json."""
    t = json.loads((ROOT / "tests" / "golden" / "model_traces.json").read_text())[name]
    rows = [tuple(r) for r in t["rows"]]
    return rows, t


def run_model_trace(a, name):
    """Hot-path replay of one training step of a reference model (QUAN-YOLO11n/s-OBB at 1024^2, Q-ResNet-34 at 224^2): every
    QConv2D (+ IQBN + SiLU) call of the model with its real shape, forward + backward, on synthetic activations.  Glue
    the reference does in Python (concat, split, attention matmuls, pooling, heads, loss) is not part of the path and
    not replayed."""
    import quan_ultralytics_b200 as Q
    lib = Q._lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    B = a.n
    rows, meta = load_model_trace(name)
    image = meta["image"]
    mixB = meta["mix"] == "B"
    torch.manual_seed(0)
    layers = []
    fwd_flops = 0
    bn_elems = 0
    n_bn = 0

    class ConvB(Q.Conv):                                        # classification flavour: M_B (qconv.py:546-612)
        conv_cls = Q.QConv2D_B

    class BiasedBlock(torch.nn.Module):                         # Q-ResNet: biased QConv2D -> IQBN -> SiLU (QSiLU == SiLU)
        def __init__(self, c1, c2, k, s, g):
            super().__init__()
            self.conv = (Q.QConv2D_B if mixB else Q.QConv2D)(c1, c2, k, s, k // 2, groups=g, bias=True)
            self.bn = Q.IQBN(c2)

        def forward(self, x):
            return self.bn(self.conv(x), Q.ACT_SILU)

    for ci, co, k, s, g, ho, has_bn, has_bias, cnt in rows:
        hin = ho * s
        for rep in range(cnt):
            cin = 3 if ci == 1 and hin == image else ci * 4
            if has_bn and has_bias:
                mod = BiasedBlock(cin, co * 4, k, s, g)
            elif has_bn:
                mod = (ConvB if mixB else Q.Conv)(cin, co * 4, k, s, g=g)
            else:                                               # e.g. the three BN-less QConv2Ds of QAttention, shortcuts
                mod = (Q.QConv2D_B if mixB else Q.QConv2D)(cin, co * 4, k, s, k // 2, groups=g, bias=bool(has_bias))
            mod = mod.to(dev).train()
            if cin == 3:
                x = torch.rand(B, 3, hin, hin, device=dev) if not mixB else torch.randn(B, 3, hin, hin, device=dev)
            else:
                x = torch.randn(B, ci, hin, hin, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
                x.requires_grad_(True)
            dy = torch.randn(B, co, ho, ho, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
            layers.append((mod, x, dy))
            fwd_flops += 4 * 2 * ho * ho * co * (ci // g) * k * k
            bn_elems += co * ho * ho * 4 if has_bn else 0
            n_bn += 1 if has_bn else 0
    n_conv = len(layers)
    params = [p for m, _, _ in layers for p in m.parameters()]

    def train_step():
        for p in params:
            p.grad = None
        for mod, x, dy in layers:
            if x.grad is not None:
                x.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
                y = mod(x)
            y.backward(dy)

    def infer_step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
            for mod, x, _ in layers:
                mod(x)

    if a.infer:
        for mod, _, _ in layers:
            mod.eval()
    step = infer_step if a.infer else train_step

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if os.environ.get("QUAN_TRACE_DETAIL"):
        rows_in, rows = rows, []
        i = 0
        for ci, co, k, s, g, ho, _bn, _bias, cnt in rows_in:
            mod, x, dy = layers[i]
            i += cnt

            def f():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
                    return mod(x)
            y = f()
            tf = time_op(f, 5)
            tfb = time_op(lambda: f().backward(dy), 5)
            from quan_ultralytics_b200 import ops as _ops
            algo = [_ops.qconv2d_pick_algo((B, max(ci, 1), ho * s, ho * s, 4), (co, ci // g, k, k), (s, s), (k // 2, k // 2), (1, 1),
                                            g, dtype, _ops.LAYOUT_BHWQC, ps) for ps in range(3)]
            rows.append((cnt * tfb, f"({ci},{co},k{k},s{s},g{g},{ho}^2)x{cnt}: fwd {tf:.3f} ms, fwd+bwd {tfb:.3f} ms, algo {algo}"))
        tot = sum(r[0] for r in rows)
        for t, txt in sorted(rows, reverse=True):
            print(f"{100 * t / tot:5.1f}%  {txt}", flush=True)
    if os.environ.get("QUAN_TRACE_KERNELS"):
        # warm per-kernel device times inside the (eager) step, from the library's event pairs
        import ctypes
        lib.quan_kernel_timing_enable(1)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        lib.quan_kernel_timing_enable(0)
        n = lib.quan_kernel_timing_report(None, 0)
        buf = ctypes.create_string_buffer(n + 16)
        lib.quan_kernel_timing_report(buf, n + 16)
        rows = [ln.split() for ln in buf.value.decode().splitlines()]
        tot = sum(float(r[2]) for r in rows)
        for kname, cnt, ms in sorted(rows, key=lambda r: -float(r[2])):
            print(f"{float(ms) / 3:8.3f} ms/step {int(cnt) // 3:5d} launches {100 * float(ms) / tot:5.1f}%  "
                  f"{1e3 * float(ms) / int(cnt):7.1f} us avg  {kname}", flush=True)
        print(f"{tot / 3:8.3f} ms/step in library kernels", flush=True)
    run = step
    launches_per_step = None
    if a.graph:
        # the library's workspaces are static and every launch goes to the current stream, so a whole step captures
        n0 = lib.quan_launch_count()
        step()
        launches_per_step = lib.quan_launch_count() - n0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        run = graph.replay
        for _ in range(2):
            run()
        torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    n0 = lib.quan_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.quan_launch_count() - n0
    if launches_per_step is not None:
        launches = launches_per_step * a.steps          # replayed by the graph, counted when it was recorded
    clocks = sampler.stop()
    peaks = load_peaks()
    esz = 2 if dtype == torch.bfloat16 else 4
    t_step = ms / a.steps / 1e3
    # per-image roofline of the replayed path: sum over layers of max(FLOPs/peak, bytes/BW) is approximated by the two totals
    flops_img = 3 * fwd_flops
    bytes_img = 8 * bn_elems * esz                     # IQBN 3S fwd + 5S bwd; conv traffic comes on top
    line = {
        "metric": "infer_images_per_sec" if a.infer else "train_images_per_sec", "value": B * a.steps / (ms / 1e3),
        "unit": "images/s", "n_gpus": 1,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": f"{name}_quan_hotpath_trace({image}x{image},B={B}): {n_conv} QConv2D + {n_bn} IQBN.SiLU " +
                               ("eval-mode forward (no_grad)" if a.infer else "fwd+bwd") + ", per-layer replay",
                   "qconv_gflop_per_image_train": flops_img / 1e9, "iqbn_melems_per_image": bn_elems / 1e6},
        "gpu_launches": int(launches), "clocks": clocks, "cuda_graph": bool(a.graph),
        "hotpath_tflops": (fwd_flops if a.infer else flops_img) * B / t_step / 1e12,
        "roofline_floor_ms_per_step": 1e3 * B * max(flops_img / (peaks["bf16_tflops"] * 1e12), bytes_img / (peaks["hbm_gbs"] * 1e9)),
    }
    emit(line)




# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] / [4]: the QUAN-YOLO11-OBB training step on the reference's own model graph
# ---------------------------------------------------------------------------------------------------------------
def yolo_workload_name(a, scale, B):
    return (f"QUAN-YOLO11{scale}-OBB train step (yolo11{scale}-obb-quan.yaml, nc=15): {B} x 3 x {a.size}^2 per GPU, 40 rotated boxes/img, "
            f"fwd + v8OBBLoss + bwd + clip 10 + SGD(nesterov)")


def yolo_cpu_reference(a, scale, n_images, steps, warmup):
    """The UNMODIFIED reference (baseline/_ref: its own graph, PyTorch path — 4 x F.conv2d + mix, batch-statistics IQBN, its own
    v8OBBLoss) on the host cores: forward + loss + backward + clip_grad_norm_(10) + torch SGD(nesterov), fp32."""
    from quan_ultralytics_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = workloads.build_yolo_obb(scale, 15, "cpu", swapped=False).train()
    opt = workloads.yolo_sgd(model)
    batch = workloads.synthetic_obb_batch(n_images, a.size, "cpu", seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _ = model({k: v.clone() for k, v in batch.items()})
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_images / sec, sec, cores


def yolo_config(a, scale, B, world):
    """The workload definition both arms print (identical dicts: the driver compares them)."""
    return {"workload": yolo_workload_name(a, scale, B), "parallelism": f"dp{world}",
            "l2": "per-step activations (~3 GB saved for backward) exceed the 126 MB L2"}


def yolo_reference_arm(a):
    """`--impl reference`: the unmodified reference on the host cores, K timed steps after W warm-up steps as asked; each step is a
    bounded sample of the workload — ONE image of the 16-image batch (a 1024^2 image costs ~1-2 s of fwd + loss + bwd on 16 cores) —
    and the value is per-image normalised."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    scale = "n" if a.workload == "yolo11n_obb" else "s"
    n = 1
    ips, sec, cores = yolo_cpu_reference(a, scale, n, a.steps, a.warmup)
    emit({
        "impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": yolo_config(a, scale, a.n, a.gpus),
        "implementation": "unmodified reference model graph and PyTorch path (baseline/_ref: 4 x F.conv2d + mix, batch-statistics IQBN, "
                          "v8OBBLoss, clip_grad_norm_ + torch SGD) on the host cores, fp32",
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "reference",
                         "sample": f"{n} image/step of the same training step ({a.steps} timed steps, {a.warmup} warm-up), per-image normalised"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def parse_timing_report(lib):
    import ctypes
    n = lib.quan_kernel_timing_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.quan_kernel_timing_report(buf, n + 16)
    rows = {}
    for ln in buf.value.decode().splitlines():
        f = ln.split()
        if len(f) < 3 or int(f[1]) == 0:
            continue
        rows[f[0]] = {"launches": int(f[1]), "ms": float(f[2]), "bytes": float(f[3]) if len(f) > 3 else 0.0,
                      "flops": float(f[4]) if len(f) > 4 else 0.0}
    return rows


def run_yolo_obb(a):
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import optim, workloads
    from quan_ultralytics_b200.graphs import BucketedGradSync, GraphedTrainStep
    from quan_ultralytics_b200.loss import OBBLossFused, OBBLossStatic, pad_targets
    import torch.distributed as dist
    lib = Q._lib.load()
    scale = "n" if a.workload == "yolo11n_obb" else "s"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ac = torch.bfloat16 if a.dtype == "bf16" else None
    peaks = load_peaks()
    B, S = a.n, a.size

    torch.manual_seed(0)
    model = workloads.build_yolo_obb(scale, 15, dev, swapped=True).train()
    if world > 1:
        for t in list(model.parameters()) + [b for b in model.buffers() if b.is_floating_point()]:
            dist.broadcast(t.data, 0)
        if a.sync_iqbn:
            from quan_ultralytics_b200.distributed import convert_sync_iqbn
            convert_sync_iqbn(model)
    params = list(model.parameters())
    opt = optim.yolo_clip_sgd(model)
    batch = workloads.synthetic_obb_batch(B, S, dev, seed=1 + rank)
    tg, tm = pad_targets(batch, B)
    tg, tm = tg.to(dev), tm.to(dev)
    crit = OBBLossFused(model)
    ref_crit = lambda preds: model.loss(batch, preds)          # the reference's own criterion (utils/loss.py:941), --ref-loss / --no-graph
    sync = BucketedGradSync(params, nbuckets=a.buckets) if world > 1 else None

    def eager_step(use_ref_loss=False):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
            preds = model(batch["img"])
            loss, items = ref_crit(preds) if use_ref_loss else crit(preds, {"targets": tg, "target_mask": tm})
        loss.backward()
        if sync is not None:
            sync.finish()
        opt.step()
        return loss

    if a.no_graph:
        static_img = batch["img"]
        run = lambda: eager_step(True)
        captured = None
    elif a.ref_loss:
        gs = GraphedTrainStep(lambda img: model(img), lambda preds, img: ref_crit(preds), opt, [batch["img"]], params, autocast=ac,
                              grad_sync=sync)
        static_img = gs.static_inputs[0]
        run = lambda: gs([static_img])[0]
        captured = gs.captured_launches
    else:
        gs = GraphedTrainStep(lambda img, t, m: model(img), lambda preds, img, t, m: crit(preds, {"targets": t, "target_mask": m}), opt,
                              [batch["img"], tg, tm], params, autocast=ac, grad_sync=sync, capture_loss=True)
        static_img = gs.static_inputs[0]
        run = lambda: gs(gs.static_inputs)[0]
        captured = gs.captured_launches

    # ---- e2e leg: every step copies ITS batch from pinned host memory — uint8 images as a loader delivers them + the padded targets —
    # on a side stream into one of two staging buffers (the copy of step i+1 overlaps step i), converts on the device as the reference's
    # preprocess_batch does (`batch["img"].to(device).float() / 255`, models/yolo/detect/train.py:57-59) and reads the loss back
    g = torch.Generator().manual_seed(100 + rank)
    img_host = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory()
    tg_host, tm_host = tg.cpu().pin_memory(), tm.cpu().pin_memory()
    stage = [(torch.empty_like(img_host, device=dev), torch.empty_like(tg), torch.empty_like(tm)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    st = {"i": 0, "primed": False}

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            stage[slot][0].copy_(img_host, non_blocking=True)
            stage[slot][1].copy_(tg_host, non_blocking=True)
            stage[slot][2].copy_(tm_host, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step():
        i = st["i"]
        cur = i & 1
        if not st["primed"]:
            consumed[0].record()
            consumed[1].record()
            prefetch(cur)
            st["primed"] = True
        prefetch(cur ^ 1)
        torch.cuda.current_stream().wait_event(copied[cur])
        static_img.copy_(stage[cur][0])                       # uint8 -> float32 ...
        static_img.mul_(1.0 / 255.0)                          # ... / 255
        if not a.no_graph and not a.ref_loss:
            gs.static_inputs[1].copy_(stage[cur][1])
            gs.static_inputs[2].copy_(stage[cur][2])
        consumed[cur].record()
        loss = run()
        loss_host[cur].copy_(loss.detach().reshape(1).float(), non_blocking=True)
        done[cur].record()
        if i > 0:
            done[cur ^ 1].synchronize()
        st["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(a.warmup, 3)):
        run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.quan_launch_count()
    ms = timed(run, a.steps)
    launches = lib.quan_launch_count() - n0 if captured is None else captured * a.steps
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    st["primed"] = False
    ms_e2e = timed(e2e_step, a.steps)
    final_loss = float(loss_host[(st["i"] - 1) & 1][0])
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel device times of the library inside the (eager) step: every launch bracketed by a CUDA-event pair on its own
    # stream, with the algorithmic bytes / FLOPs each API call announces (all ranks run it: the step holds collectives)
    table = None
    if not a.no_kernel_table:
        for _ in range(2):
            eager_step()
        torch.cuda.synchronize()
        lib.quan_kernel_timing_enable(1)
        ksteps = 3
        for _ in range(ksteps):
            eager_step()
        torch.cuda.synchronize()
        lib.quan_kernel_timing_enable(0)
        table = parse_timing_report(lib)

    imgs = B * world * a.steps
    if rank == 0:
        line = {
            "metric": "train_images_per_sec", "value": imgs / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": yolo_config(a, scale, B, world),
            "implementation": {"step": ("eager" if a.no_graph else "2 CUDA graphs (forward+loss | backward+all-reduce+optimizer)" if not a.ref_loss
                                        else "2 CUDA graphs around the reference's eager v8OBBLoss"),
                               "loss": "reference v8OBBLoss" if (a.no_graph or a.ref_loss) else "OBBLossFused (decode + rotated TAL assigner + loss/gradient kernels, 9 launches)",
                               "optimizer": "ClipSGD: clip_grad_norm_(10) + SGD(lr .01, momentum .937, nesterov, wd 5e-4) in 2 launches",
                               "sync_iqbn": bool(a.sync_iqbn and world > 1), "grad_buckets": a.buckets if world > 1 else None,
                               "params": sum(p.numel() for p in params)},
            "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": img_host.numel() + 4 * (tg_host.numel() + tm_host.numel()), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "final_loss": final_loss,
        }
        if table:
            esz = 2.0 if a.dtype == "bf16" else 4.0
            tpeak = peaks["bf16_tflops"] * (1.0 if a.dtype == "bf16" else 0.5) * 1e12     # burst: every kernel here runs for microseconds
            bw = peaks["hbm_gbs"] * 1e9
            rows, lib_ms, floor_ms = {}, 0.0, 0.0
            for name, r in table.items():
                per_step = r["ms"] / ksteps
                lib_ms += per_step
                row = {"launches_per_step": r["launches"] / ksteps, "ms_per_step": per_step, "us_per_launch": 1e3 * r["ms"] / r["launches"]}
                if r["bytes"] > 0 or r["flops"] > 0:
                    t_hbm, t_tc = r["bytes"] / bw, r["flops"] / tpeak
                    bound = "hbm" if t_hbm >= t_tc else "tensor"
                    sec = r["ms"] / 1e3
                    ach = r["bytes"] / sec / 1e9 if bound == "hbm" else r["flops"] / sec / 1e12
                    pk = peaks["hbm_gbs"] if bound == "hbm" else tpeak / 1e12
                    row.update({"bound": bound, "achieved": ach, "peak": pk, "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": ach / pk,
                                "algorithmic_mb_per_step": r["bytes"] / ksteps / 1e6, "algorithmic_gflop_per_step": r["flops"] / ksteps / 1e9})
                    floor_ms += 1e3 * max(t_hbm, t_tc) / ksteps
                rows[name] = row
            dom = max((k for k in rows if "frac" in rows[k]), key=lambda k: rows[k]["ms_per_step"])
            r = {k: rows[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
            r.update({"kernel": dom, "peak_src": peaks["src"] + " (burst)", "traffic": ncu_traffic().get(dom + ":" + a.workload),
                      "us_per_launch": rows[dom]["us_per_launch"], "launches_per_step": rows[dom]["launches_per_step"],
                      "share_of_library_kernel_time": rows[dom]["ms_per_step"] / lib_ms,
                      "note": "achieved = algorithmic bytes (or FLOPs) of all launches of this kernel in one step / their summed device time, "
                              "measured with CUDA-event pairs on the launching stream inside the eager step"})
            line["roofline"] = r
            line["library_kernel_ms_per_step"] = lib_ms
            line["library_roofline_floor_ms_per_step"] = floor_ms
            line["kernels_in_step"] = dict(sorted(rows.items(), key=lambda kv: -kv[1]["ms_per_step"]))
        if world == 1 and not a.no_cpu_baseline:
            ips, sec, cores = yolo_cpu_reference(a, scale, 1, a.cpu_steps, 1)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "reference",
                                    "sample": f"1 image/step of the same training step through the unmodified reference (baseline/_ref), "
                                              f"{a.cpu_steps} timed steps after 1 warm-up, fp32, per-image normalised"}
        emit(line)
    if world > 1:
        # the captured graphs hold NCCL work on the communicator: tearing the process group down underneath them can wait forever
        # (watchdog dump after minutes); every rank has finished its timed work at this point, so leave without the teardown
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[1]: the QConv2D / IQBN layer sweep, machine-readable
# ---------------------------------------------------------------------------------------------------------------
def measure_matmul_peak(dtype, tf32):
    """cuBLAS GEMM 8192^3 on this device, best of 10 (burst): the tensor-pipe denominator for `dtype` (bf16, or fp32 with TF32 MMA)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(8192, 8192, device="cuda", dtype=dtype)
    b = torch.randn(8192, 8192, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    return 2 * 8192 ** 3 / best / 1e9


def run_sweep(a):
    """C_q in {64,128,256,512} x stride {1,2} x N in {16,64,256} x {bf16, tf32}: QConv2D fwd / dgrad / wgrad (3x3, pad 1, M_A, no bias) and the
    IQBN kernels on the conv's output shape, each op alone with the L2 swept between launches.  One JSON object on stdout:
    {"peaks": {...}, "rows": [{C, stride, N, HW, dtype, op, ms, achieved, unit, peak, frac}, ...]}."""
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import ops
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = ops.LAYOUT_BHWQC
    peaks = load_peaks()
    peaks["bf16_tflops_cublas_now"] = measure_matmul_peak(torch.bfloat16, False)
    peaks["tf32_tflops_cublas_now"] = measure_matmul_peak(torch.float32, True)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    rows = []
    Cs = [int(v) for v in os.environ.get("QUAN_SWEEP_C", "64,128,256,512").split(",")]
    Ns = [int(v) for v in os.environ.get("QUAN_SWEEP_N", "16,64,256").split(",")]
    for dt in ("bf16", "tf32"):
        dtype = torch.bfloat16 if dt == "bf16" else torch.float32
        esz = 2 if dt == "bf16" else 4
        tpeak = peaks["bf16_tflops"] if dt == "bf16" else peaks["tf32_tflops_cublas_now"]
        for C in Cs:
            H = 64 if C <= 128 else 32 if C == 256 else 16
            for s_ in (1, 2):
                for N in Ns:
                    torch.manual_seed(1234)
                    x = torch.randn(N, C, H, H, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
                    w = [torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5 for _ in range(4)]
                    args = ((s_, s_), (1, 1), (1, 1), 1, ops.M_A)
                    y = ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L)
                    dy = torch.randn_like(y)
                    Ho = y.shape[2]
                    flops = 8.0 * N * Ho * Ho * C * C * 9
                    Sy = y.numel() * esz
                    gamma, beta = torch.ones(C, 4, device=dev), torch.zeros(C, 4, device=dev)
                    stats = ops.iqbn_train_stats(y, L, gamma, beta, 1e-5, 0.1, None, None)
                    cnt = float(N * Ho * Ho)
                    sums = ops.iqbn_bwd_reduce(dy, y, L, stats, gamma, beta, Q.ACT_SILU, cnt)
                    todo = [
                        ("qconv2d_fwd", lambda: ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L), "tensor", flops),
                        ("qconv2d_dgrad", lambda: ops.qconv2d_bwd(dy, x, w, *args, True, False, False), "tensor", flops),
                        ("qconv2d_wgrad", lambda: ops.qconv2d_bwd(dy, x, w, *args, False, True, False), "tensor", flops),
                    ]
                    if s_ == 1:          # the IQBN kernels do not depend on the stride: once per (C, N, dtype)
                        todo += [
                            ("iqbn_train_stats", lambda: ops.iqbn_train_stats(y, L, gamma, beta, 1e-5, 0.1, None, None), "hbm", Sy),
                            ("iqbn_apply_fwd_silu", lambda: ops.iqbn_apply_fwd(y, L, stats, gamma, beta, Q.ACT_SILU), "hbm", 2 * Sy),
                            ("iqbn_bwd_reduce", lambda: ops.iqbn_bwd_reduce(dy, y, L, stats, gamma, beta, Q.ACT_SILU, cnt), "hbm", 2 * Sy),
                            ("iqbn_bwd_apply", lambda: ops.iqbn_bwd_apply(dy, y, L, stats, gamma, beta, Q.ACT_SILU, sums, cnt), "hbm", 3 * Sy),
                        ]
                    for name, fn, kind, work in todo:
                        ms = time_op(fn, 6, flush)
                        ach = work / ms / 1e9 if kind == "tensor" else work / ms / 1e6
                        pk = tpeak if kind == "tensor" else peaks["hbm_gbs"]
                        rows.append({"C": C, "stride": s_, "N": N, "HW": H, "dtype": dt, "op": name, "ms": ms, "bound": kind,
                                     "achieved": ach, "unit": "TFLOP/s" if kind == "tensor" else "GB/s", "peak": pk, "frac": ach / pk})
                    del x, y, dy, w
                    torch.cuda.empty_cache()
    # ---- the HBM-bound ops around the convolutions at the QUAN-YOLO11n shapes (16 x 1024^2): algorithmic bytes of SURVEY 8(d)
    def hbm_row(name, shape, fn, nbytes):
        ms = time_op(fn, 6, flush)
        ach = nbytes / ms / 1e6
        rows.append({"op": name, "shape": shape, "dtype": "bf16", "ms": ms, "bound": "hbm", "achieved": ach, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                     "frac": ach / peaks["hbm_gbs"], "algorithmic_mb": nbytes / 1e6})

    bf = torch.bfloat16
    cl = torch.channels_last_3d
    img = torch.rand(16, 3, 1024, 1024, device=dev)
    qimg = ops.poincare_fwd(img, bf)
    gq = torch.randn_like(qimg)
    npx = 16 * 1024 * 1024
    hbm_row("poincare_fwd", "16x3x1024^2", lambda: ops.poincare_fwd(img, bf), npx * (12 + 8))
    hbm_row("poincare_bwd", "16x3x1024^2", lambda: ops.poincare_bwd(img, gq), npx * (12 + 8 + 12))
    del img, qimg, gq
    for (C, H) in ((64, 32), (32, 64)):                      # yolo11-obb-quan.yaml:35,39
        xu = torch.randn(16, C, H, H, 4, device=dev).to(bf).contiguous(memory_format=cl)
        yu = ops.qupsample_fwd(xu, 2)
        S = xu.numel() * 2
        hbm_row("qupsample_fwd", f"16x{C}x{H}^2 -> {2 * H}^2", lambda: ops.qupsample_fwd(xu, 2), 5 * S)
        hbm_row("qupsample_bwd", f"16x{C}x{H}^2 <- {2 * H}^2", lambda: ops.qupsample_bwd(yu, 2), 5 * S)
    xp = torch.randn(16, 32, 32, 32, 4, device=dev).to(bf).contiguous(memory_format=cl)     # QSPPF: 5x5 stride 1 pad 2
    yp, ip = ops.qmaxpool_fwd(xp, 5, 1, 2, with_idx=True)
    Sp = xp.numel() * 2
    hbm_row("qmaxpool_fwd (5,1,2)", "16x32x32^2", lambda: ops.qmaxpool_fwd(xp, 5, 1, 2, with_idx=True), 2.5 * Sp)
    hbm_row("qmaxpool_bwd (5,1,2)", "16x32x32^2", lambda: ops.qmaxpool_bwd(yp, ip, (32, 32), 5, 1, 2), 2.5 * Sp)
    for (C, H, N) in ((16, 128, 64), (16, 128, 15), (16, 64, 64)):          # head extractions: box / class at P3, box at P4
        xq = torch.randn(16, C, H, H, 4, device=dev).to(bf).contiguous(memory_format=cl)
        wq, bq = torch.randn(N, 4 * C, 1, 1, device=dev), torch.randn(N, device=dev)
        buf = torch.empty(16, H, H, 80, device=dev, dtype=bf)
        dq = torch.randn(16, H, H, 80, device=dev).to(bf)[..., :N].permute(0, 3, 1, 2)
        by = 16 * H * H * (4 * C + N) * 2
        hbm_row("qer_fwd", f"16x{C}x{H}^2 -> {N}", lambda: ops.qer_fwd(xq, wq, bq, buf, 0, 80 if N % 8 else N), by)
        hbm_row("qer_dgrad", f"16x{C}x{H}^2 <- {N}", lambda: ops.qer_bwd(dq, xq, wq, True, False, False, 80 if N % 8 else 0), by)
        hbm_row("qer_wgrad", f"16x{C}x{H}^2 x {N}", lambda: ops.qer_bwd(dq, xq, wq, False, True, True, 80 if N % 8 else 0), by)
    a0 = torch.randn(16, 32, 64, 64, 4, device=dev).to(bf).contiguous(memory_format=cl)
    b0 = torch.randn(16, 16, 64, 64, 4, device=dev).to(bf).contiguous(memory_format=cl)
    h0, h1 = a0.chunk(2, 1)
    hbm_row("rows_cat (C2f: 2 halves + 1)", "16x(16+16+16)x64^2", lambda: ops.qcat([h0, h1, b0]), 2 * (a0.numel() + b0.numel()) * 2)
    emit({"workload": "QConv2D/IQBN layer sweep (BASELINE configs[1])", "peaks": peaks,
                      "note": "ops timed alone, L2 swept between launches; conv ms include the weight-packing / mix pre-pass / split-K fold "
                              "kernels of the call; tensor fractions against the measured bf16 burst peak (MEASURED_PEAKS.json) and, for tf32, "
                              "against cuBLAS TF32 8192^3 measured in this run; rows with a `shape` key: the HBM-bound ops around the convolutions at the "
                              "QUAN-YOLO11n shapes (host launch latency of the Python wrapper is inside the small ones' times)", "rows": rows})


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[0] / [3]: Q-WRN-16-2 (128 x 3 x 32 x 32) and Q-ResNet-34 (256 x 3 x 224 x 224 per GPU, DDP + synced IQBN) training steps
# ---------------------------------------------------------------------------------------------------------------
CLASSIFIERS = {"qresnet34": dict(size=224, classes=1000, title="Q-ResNet-34 (create_qrn34_imagenet)", lr=0.1, wd=1e-4, clip=1.0),
               "qwrn16_2": dict(size=32, classes=10, title="Q-WRN-16-2 (create_qwrn_16_2, Poincare mapping)", lr=0.1, wd=1e-4, clip=1.0)}


def classifier_config(a, world):
    c = CLASSIFIERS[a.workload]
    return {"workload": f"{c['title']} train step: {a.n} x 3 x {c['size']}^2 per GPU, cross-entropy, clip {c['clip']} + SGD(lr .1, momentum .9, wd 1e-4, nesterov)",
            "parallelism": f"dp{world}"}


def classifier_cpu_reference(a, n_images, steps, warmup):
    """The unmodified reference model (baseline/_ref/classification, PyTorch path) on the host cores: forward + CE + backward +
    clip_grad_norm_(1.0) + torch SGD (classification.py:202-203, utils/training.py:77-79), fp32."""
    from quan_ultralytics_b200 import workloads
    c = CLASSIFIERS[a.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = workloads.build_classifier(a.workload, c["classes"], "cpu", swapped=False).train()
    opt = torch.optim.SGD(model.parameters(), lr=c["lr"], momentum=0.9, weight_decay=c["wd"], nesterov=True)
    x, y = workloads.synthetic_classification_batch(n_images, c["size"], c["classes"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(model(x), y).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), c["clip"])
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_images / sec, sec, cores


def classifier_reference_arm(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n = 8 if a.workload == "qresnet34" else 128
    ips, sec, cores = classifier_cpu_reference(a, n, a.steps, a.warmup)
    emit({
        "impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": classifier_config(a, a.gpus),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "reference",
                         "sample": f"{n} images/step of the same training step through the unmodified reference, per-image normalised"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def run_classifier(a):
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import optim, workloads
    from quan_ultralytics_b200.graphs import BucketedGradSync, GraphedTrainStep
    import torch.distributed as dist
    lib = Q._lib.load()
    c = CLASSIFIERS[a.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ac = torch.bfloat16 if a.dtype == "bf16" else None
    B, S = a.n, c["size"]
    torch.manual_seed(0)
    model = workloads.build_classifier(a.workload, c["classes"], dev, swapped=True).train()
    sync_iqbn = world > 1 and (a.sync_iqbn or a.workload == "qresnet34")      # configs[3] asks for synced IQBN
    if world > 1:
        for t in list(model.parameters()) + [b for b in model.buffers() if b.is_floating_point()]:
            dist.broadcast(t.data, 0)
        if sync_iqbn:
            from quan_ultralytics_b200.distributed import convert_sync_iqbn
            convert_sync_iqbn(model)
    params = list(model.parameters())
    opt = optim.ClipSGD([{"params": params, "lr": c["lr"], "weight_decay": c["wd"]}], momentum=0.9, nesterov=True, max_norm=c["clip"])
    x, y = workloads.synthetic_classification_batch(B, S, c["classes"], dev, seed=1 + rank)
    sync = BucketedGradSync(params, nbuckets=a.buckets) if world > 1 else None
    gs = GraphedTrainStep(lambda xx, yy: model(xx), lambda out, xx, yy: (torch.nn.functional.cross_entropy(out.float(), yy), None), opt, [x, y],
                          params, autocast=ac, grad_sync=sync, capture_loss=True)
    run = lambda: gs(gs.static_inputs)[0]
    # e2e: uint8 images + labels from pinned host memory, normalised on the device (classification/utils/data_loading.py:90 mean / std)
    g = torch.Generator().manual_seed(100 + rank)
    img_host = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory()
    y_host = y.cpu().pin_memory()
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1) * 255
    inv_std = 1.0 / (torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1) * 255)
    stage = [(torch.empty_like(img_host, device=dev), torch.empty_like(y)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied, consumed, done = ([torch.cuda.Event(), torch.cuda.Event()] for _ in range(3))
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    st = {"i": 0, "primed": False}

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            stage[slot][0].copy_(img_host, non_blocking=True)
            stage[slot][1].copy_(y_host, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step():
        i = st["i"]
        cur = i & 1
        if not st["primed"]:
            consumed[0].record()
            consumed[1].record()
            prefetch(cur)
            st["primed"] = True
        prefetch(cur ^ 1)
        torch.cuda.current_stream().wait_event(copied[cur])
        gs.static_inputs[0].copy_(stage[cur][0])
        gs.static_inputs[0].sub_(mean).mul_(inv_std)
        gs.static_inputs[1].copy_(stage[cur][1])
        consumed[cur].record()
        loss = run()
        loss_host[cur].copy_(loss.detach().reshape(1).float(), non_blocking=True)
        done[cur].record()
        if i > 0:
            done[cur ^ 1].synchronize()
        st["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(a.warmup, 3)):
        run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms = timed(run, a.steps)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    st["primed"] = False
    ms_e2e = timed(e2e_step, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    imgs = B * world * a.steps
    if rank == 0:
        line = {"metric": "train_images_per_sec", "value": imgs / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": a.dtype, "data": "synthetic", "config": classifier_config(a, world),
                "implementation": {"step": "2 CUDA graphs (forward+loss | backward+all-reduce+optimizer)", "sync_iqbn": bool(sync_iqbn),
                                   "grad_buckets": a.buckets if world > 1 else None, "params": sum(p.numel() for p in params)},
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": img_host.numel() + 8 * B, "d2h_bytes_per_step": 4},
                "gpu_launches": int(gs.captured_launches * a.steps), "clocks": clocks, "final_loss": float(loss_host[(st["i"] - 1) & 1][0])}
        if world == 1 and not a.no_cpu_baseline:
            n = 8 if a.workload == "qresnet34" else 128
            ips, sec, cores = classifier_cpu_reference(a, n, min(a.cpu_steps, 2), 1)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "reference",
                                    "sample": f"{n} images/step through the unmodified reference model (baseline/_ref), fp32"}
        emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    a = parse_args()
    json_only_stdout()
    if a.workload.endswith("_trace") and a.impl == "ours":
        name = a.workload[:-len("_trace")]
        if a.n_default:                                         # per-GPU batch of the BASELINE config that names the model
            a.n = {"yolo11n": 16, "yolo11s": 8, "qresnet34": 256}[name]
        run_model_trace(a, name)
        return
    if a.workload.endswith("_obb"):
        (yolo_reference_arm if a.impl == "reference" else run_yolo_obb)(a)
        return
    if a.workload == "sweep":
        run_sweep(a)
        return
    if a.workload in ("qresnet34", "qwrn16_2"):
        (classifier_reference_arm if a.impl == "reference" else run_classifier)(a)
        return
    if a.impl == "reference":
        reference_arm(a)
        return

    import quan_ultralytics_b200 as Q
    lib = Q._lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    peaks = load_peaks()

    torch.manual_seed(1 + rank)
    conv_cls = Q.Conv
    blocks = []
    for _ in range(a.depth):
        b = conv_cls(a.cq * 4, a.cq * 4, 3, 1)
        b.conv.mix = a.mix
        blocks.append(b)
    net = torch.nn.Sequential(*blocks).to(dev).train()
    if a.sync_iqbn and world > 1:
        from quan_ultralytics_b200.distributed import convert_sync_iqbn
        convert_sync_iqbn(net)
    model = net
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=a.ddp_bucket_mb,
                                                          broadcast_buffers=a.broadcast_buffers)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, foreach=True)

    shape = (a.n, a.cq, a.hw, a.hw, 4)
    fmt = torch.channels_last_3d
    x_dev = torch.randn(shape, device=dev).to(dtype).contiguous(memory_format=fmt)
    dy_dev = torch.randn(shape, device=dev).to(dtype).contiguous(memory_format=fmt)
    # host copies for the e2e leg, already in the layout the device uses
    x_host = torch.empty(x_dev.shape, dtype=dtype).contiguous(memory_format=fmt).pin_memory()
    x_host.copy_(x_dev)
    x_stage = torch.empty_like(x_dev, memory_format=torch.preserve_format)
    g_host = [torch.empty_like(net[0].conv.weight_r, device="cpu").pin_memory() for _ in range(2)]
    step_done = [torch.cuda.Event(), torch.cuda.Event()]

    def step(xin):
        opt.zero_grad(set_to_none=True)
        y = model(xin)
        y.backward(dy_dev)
        opt.step()

    # e2e leg: every step copies ITS input batch from pinned host memory and reads its result back.  The copies run on
    # a side stream into two staging buffers, so the H2D of step i+1 overlaps the compute of step i (what any input
    # pipeline does); every byte still moves inside the timed region, including the first batch.
    copy_stream = torch.cuda.Stream(device=dev)
    x_stage2 = [x_stage, torch.empty_like(x_stage, memory_format=torch.preserve_format)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "primed": False}

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])            # the step that last read this buffer has finished
        with torch.cuda.stream(copy_stream):
            x_stage2[slot].copy_(x_host, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step():
        i = e2e_state["i"]
        cur = i & 1
        if not e2e_state["primed"]:
            consumed[0].record()
            consumed[1].record()
            prefetch(cur)                                  # H2D of this step's inputs (first step: not overlapped)
            e2e_state["primed"] = True
        prefetch(cur ^ 1)                                  # next step's inputs, overlapping this step's compute
        torch.cuda.current_stream().wait_event(copied[cur])
        step(x_stage2[cur])
        consumed[cur].record()
        g_host[cur].copy_(net[0].conv.weight_r.grad, non_blocking=True)   # D2H of this step's result
        step_done[cur].record()
        if i > 0:
            step_done[cur ^ 1].synchronize()               # the host consumes step i-1's result while step i runs
        e2e_state["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(a.warmup, 3)):
        step(x_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.quan_launch_count()
    ms = timed(lambda: step(x_dev), a.steps)
    launches = lib.quan_launch_count() - n0
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    e2e_state["primed"] = False                            # the timed region starts with nothing staged
    ms_e2e = timed(e2e_step, a.steps)
    clocks = sampler.stop() if rank == 0 else None

    # every library kernel timed inside the training step: ALL ranks run these extra steps (the step holds DDP's
    # gradient all-reduce), rank 0 reports
    ks = kernels_in_step(lib, lambda: step(x_dev), a.steps, a, dtype, peaks) if not a.no_kernel_table else None

    imgs = a.n * world * a.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    if rank == 0:
        line = {
            "metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "activations (%.0f MB/tensor) exceed the 126 MB L2" %
                       (x_dev.numel() * x_dev.element_size() / 1e6), "parallelism": f"dp{world}",
                       "sync_iqbn": bool(a.sync_iqbn and world > 1), "optimizer": "SGD(momentum=0.9)",
                       "ddp_bucket_mb": a.ddp_bucket_mb if world > 1 else None,
                       "ddp_broadcast_buffers": bool(a.broadcast_buffers) if world > 1 else None,
                       "train_gflop_per_image": flops_per_image(a) / 1e9},
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
                    "d2h_bytes_per_step": g_host[0].numel() * g_host[0].element_size()},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "step_tflops": flops_per_image(a) * a.n * world / (ms / a.steps) / 1e9,
        }
        if not a.no_kernel_table:
            # (1) the dominant in-step kernel carries the roofline
            step_ms = sum(v["ms_per_step"] for v in ks.values())
            dom = max((k for k in ks if "frac" in ks[k]), key=lambda k: ks[k]["ms_per_step"])
            r = {k: ks[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
            tr = ncu_traffic().get(dom)
            r.update({"kernel": dom, "peak_src": peaks["src"] + (" (sustained: kernel timed inside the step)" if r["bound"] == "tensor" else ""),
                      "traffic": tr, "ms_per_launch": ks[dom]["ms_per_launch"],
                      "share_of_step_kernel_time": ks[dom]["ms_per_step"] / step_ms})
            line["roofline"] = r
            line["kernels_in_step"] = ks
            # (2) each op of one block timed ALONE (L2 swept between launches), against the burst peak
            line["ops_isolated"] = kernel_table(a, dev, dtype, peaks)
        if world == 1:
            ext = reference_cuda_ext(a, dev)
            if ext is not None:
                line["reference_cuda_ext"] = ext
        if world == 1 and not a.no_cpu_baseline:
            ips, sec, cores = run_cpu_port(a, a.cpu_n, a.cpu_steps, 1)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{a.cpu_n} images/step of the same block stack, {a.cpu_steps} timed steps, fp32"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
