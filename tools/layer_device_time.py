"""Device time of every distinct QConv2D(+IQBN+SiLU) layer of a model trace, forward and backward, measured WITHOUT the host in the
loop: each op is captured `reps` times into a CUDA graph and the replay is timed with CUDA events (narrow layers take 10-40 us on the
device but ~100 us of host time per eager call, so eager per-layer timings only show the host).  Printed next to the layer's HBM
floor (algorithmic bytes / measured copy bandwidth) and tensor floor.
    python tools/layer_device_time.py [--model yolo11n] [--batch 16] [--reps 10]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import quan_ultralytics_b200 as Q  # noqa: E402
from quan_ultralytics_b200 import ops  # noqa: E402


def graph_time(fn, reps, iters=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    e1.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (iters * reps)          # us per op


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolo11n")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    t = json.loads((ROOT / "tests" / "golden" / "model_traces.json").read_text())[a.model]
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6551.7, "bf16_tflops": 1678.6}
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    esz = 2 if dtype == torch.bfloat16 else 4
    dev = "cuda"
    L = ops.LAYOUT_BHWQC
    mix = ops.MIX[t["mix"]]
    B = a.batch
    tot = {"fwd": 0.0, "bwd": 0.0, "ffl": 0.0, "bfl": 0.0}
    print(f"{'layer (Ci,Co,k,s,g,Ho)':28s} cnt   fwd us (floor)   bwd us (floor)   engine fwd/dgrad/wgrad")
    for ci, co, k, s, g, ho, has_bn, has_bias, cnt in t["rows"]:
        if not has_bn or has_bias or (ci == 1 and ho * s == t["image"]):
            continue                                            # the fused Conv block only (the bulk of the model)
        hin = ho * s
        x = torch.randn(B, ci, hin, hin, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
        w = [torch.randn(co, ci // g, k, k, device=dev) * 0.1 for _ in range(4)]
        gamma, beta = torch.ones(co, 4, device=dev), torch.zeros(co, 4, device=dev)
        rm, rv = torch.zeros(co, 4, device=dev), torch.ones(co, 4, device=dev)
        args = ((s, s), (k // 2, k // 2), (1, 1), g, mix, Q.ALGO_AUTO)
        y, out, stats = ops.conv_block_fwd(x, w, gamma, beta, rm, rv, *args, 1e-5, 0.1, Q.ACT_SILU, L, True)
        dout = torch.randn_like(out)
        f = lambda: ops.conv_block_fwd(x, w, gamma, beta, rm, rv, *args, 1e-5, 0.1, Q.ACT_SILU, L, True)
        b = lambda: ops.conv_block_bwd(dout, x, y, w, stats, gamma, beta, *args, Q.ACT_SILU, L, True, True)
        tf, tb = graph_time(f, a.reps), graph_time(b, a.reps)
        sx, sy = x.numel() * esz, y.numel() * esz
        flops = 8.0 * B * ho * ho * co * (ci // g) * k * k
        # algorithmic bytes: fwd = conv (x, y) + IQBN apply (y, out) [statistics ride in the epilogue]; bwd = IQBN reduce (dout, y) +
        # apply (dout, y, g) + dgrad (g, dx) + wgrad (g, x)
        fb, bb = sx + 3 * sy, 2 * sy + 3 * sy + (sy + sx) + (sy + sx)
        ffl = max(fb / (peaks["hbm_gbs"] * 1e3), flops / (peaks["bf16_tflops"] * 1e6))
        bfl = max(bb / (peaks["hbm_gbs"] * 1e3), 2 * flops / (peaks["bf16_tflops"] * 1e6))
        algo = [ops.qconv2d_pick_algo(x.shape, w[0].shape, (s, s), (k // 2, k // 2), (1, 1), g, dtype, L, ps) for ps in range(3)]
        print(f"({ci},{co},k{k},s{s},g{g},{ho}^2)".ljust(28) + f" {cnt:3d}  {tf:7.1f} ({ffl:5.1f})  {tb:7.1f} ({bfl:5.1f})   {algo}", flush=True)
        for key, v in (("fwd", tf), ("bwd", tb), ("ffl", ffl), ("bfl", bfl)):
            tot[key] += cnt * v
    print(f"sum over the model's Conv blocks: fwd {tot['fwd'] / 1e3:.2f} ms (floor {tot['ffl'] / 1e3:.2f}), bwd {tot['bwd'] / 1e3:.2f} ms "
          f"(floor {tot['bfl'] / 1e3:.2f})")


if __name__ == "__main__":
    main()
