"""Bring-up probe for the tcgen05 engine: one case per process (a trapped kernel poisons the CUDA context), compares
the tensor-core path with the golden-validated direct engine on the same device tensors.

    python tests/tc_probe.py <case> [pass]      # pass: fwd | dgrad | wgrad | all
Prints one line per check; exit code 0 iff all checks are within tolerance.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import ops  # noqa: E402

# name: (dtype, B, Ci, Co, H, W, k, s, p, d, bias, mix)
CASES = [
    ("bf16_c64", "bf16", 2, 64, 64, 16, 16, 3, 1, 1, 1, False, "A"),
    ("bf16_c128", "bf16", 2, 128, 128, 16, 16, 3, 1, 1, 1, False, "A"),
    ("tf32_c64", "f32", 2, 64, 64, 16, 16, 3, 1, 1, 1, False, "A"),
    ("bf16_c16_sw32", "bf16", 2, 16, 16, 16, 16, 3, 1, 1, 1, False, "A"),
    ("bf16_c32_k1_bn48", "bf16", 2, 32, 48, 16, 16, 1, 1, 0, 1, False, "A"),
    ("bf16_s2", "bf16", 2, 64, 64, 16, 16, 3, 2, 1, 1, False, "A"),
    ("bf16_7x7_bias_mixB", "bf16", 4, 64, 32, 7, 7, 3, 1, 1, 1, True, "B"),
    ("bf16_c256_32x32", "bf16", 4, 256, 256, 32, 32, 3, 1, 1, 1, False, "A"),
    ("tf32_c16", "f32", 2, 16, 16, 16, 16, 3, 1, 1, 1, False, "B"),
    ("bf16_dil2", "bf16", 2, 64, 64, 16, 16, 3, 1, 2, 2, False, "A"),
    ("bf16_c512_16x16", "bf16", 8, 512, 512, 16, 16, 3, 1, 1, 1, False, "A"),
    ("tf32_c128_s2", "f32", 2, 128, 128, 32, 32, 3, 2, 1, 1, True, "A"),
    ("bf16_k7s2", "bf16", 2, 16, 32, 32, 32, 7, 2, 3, 1, False, "A"),
    ("bf16_c96_64", "bf16", 2, 96, 64, 20, 12, 1, 1, 0, 1, False, "A"),
    # more work units than SMs: every CTA (pair) of the persistent kernel loops over several units (barrier phases,
    # accumulator hand-back, odd tile count -> masked spare tile of the last pair)
    ("bf16_c64_persist", "bf16", 75, 64, 64, 32, 32, 3, 1, 1, 1, True, "B"),
    ("tf32_c32_persist", "f32", 40, 32, 64, 40, 24, 3, 1, 1, 1, False, "A"),
    ("bf16_c128_k1_persist", "bf16", 33, 128, 256, 32, 32, 1, 1, 0, 1, False, "A"),
    # narrow layers (QUAN-YOLO11n channel counts): dense Hamilton form, 128/64/32-byte swizzle rows, partial M tiles
    ("bf16_c4_c8_dense", "bf16", 4, 4, 8, 32, 32, 3, 1, 1, 1, True, "A"),
    ("bf16_c8_k1_dense", "bf16", 4, 8, 8, 32, 32, 1, 1, 0, 1, False, "B"),
    ("bf16_c12_c16_k1_dense", "bf16", 4, 12, 16, 32, 32, 1, 1, 0, 1, False, "A"),
    ("bf16_c24_c32_dense", "bf16", 4, 24, 32, 32, 32, 3, 1, 1, 1, False, "A"),
    ("bf16_c32_s2_dense", "bf16", 4, 32, 32, 32, 32, 3, 2, 1, 1, False, "A"),
    ("bf16_c16_c4_dense", "bf16", 4, 16, 4, 32, 32, 3, 1, 1, 1, True, "B"),
    ("tf32_c8_dense", "f32", 4, 8, 8, 32, 32, 3, 1, 1, 1, True, "B"),
    ("bf16_c16_persist_dense", "bf16", 40, 16, 16, 64, 64, 3, 1, 1, 1, False, "A"),
    # dense form in halo mode (128-byte rows: C_q = 16 bf16 / 8 tf32; two k-blocks at C_q = 32), ragged 16 x 8 tiles, N != K
    ("bf16_c32_dense_halo_2kb", "bf16", 6, 32, 32, 40, 24, 3, 1, 1, 1, False, "A"),
    ("bf16_c16_c32_dense_halo_ragged", "bf16", 3, 16, 32, 30, 28, 3, 1, 1, 1, True, "B"),
    ("bf16_c32_c16_dense_halo_k5", "bf16", 5, 32, 16, 32, 32, 5, 1, 2, 1, False, "A"),
    ("tf32_c8_c16_dense_halo", "f32", 9, 8, 16, 48, 48, 3, 1, 1, 1, False, "A"),
    # strided dgrad as parity classes: odd image sizes (ragged class grids), dense and separable forms
    ("bf16_s2_odd", "bf16", 3, 32, 64, 15, 13, 3, 2, 1, 1, False, "A"),
    ("bf16_c64_s2_odd", "bf16", 5, 64, 128, 17, 19, 3, 2, 1, 1, True, "B"),
    ("bf16_c16_s2_k5", "bf16", 6, 16, 16, 32, 32, 5, 2, 2, 1, False, "A"),
    # one 128-pixel tile in total (the P5 map of a 256^2 image, batch 2): no CTA pair, so the whole 256-column weight tile sits in one CTA's
    # stage — with fp32 operands the TMA-store staging slots no longer fit beside a two-stage ring and the epilogue falls back to per-lane stores
    ("tf32_c64_one_tile", "f32", 2, 64, 64, 8, 8, 3, 1, 1, 1, False, "A"),
    ("tf32_c64_k1_one_tile", "f32", 2, 64, 64, 8, 8, 1, 1, 0, 1, True, "B"),
    ("bf16_c64_one_tile", "bf16", 2, 64, 64, 8, 8, 3, 1, 1, 1, False, "A"),
]


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    idx = int(sys.argv[1])
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    name, dt, B, Ci, Co, H, W, k, s, p, d, bias, mix = CASES[idx]
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dt == "bf16" else 1e-3
    dev = "cuda:0"
    torch.manual_seed(idx)
    L = ops.LAYOUT_BHWQC
    x = torch.randn(B, Ci, H, W, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5 for _ in range(4)]
    b = torch.randn(Co, device=dev) if bias else None
    M = ops.MIX[mix]
    args = ((s, s), (p, p), (d, d), 1, M)
    ok = True
    picks = [ops.qconv2d_pick_algo(x.shape, w[0].shape, (s, s), (p, p), (d, d), 1, dtype, L, ps) for ps in range(3)]
    print(f"[{name}] pick_algo fwd/dgrad/wgrad = {picks}", flush=True)
    y_ref = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DIRECT, L)
    if which in ("fwd", "all") and picks[0] == ops.ALGO_TCGEN05:
        y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_TCGEN05, L)
        torch.cuda.synchronize()
        e = rel(y, y_ref)
        perq = [rel(y[..., q], y_ref[..., q]) for q in range(4)]
        print(f"[{name}] fwd rel={e:.3e} per-q={['%.1e' % v for v in perq]} tol={tol}", flush=True)
        if not e <= tol:
            ok = False
            bad = ((y.double() - y_ref.double()).abs() > tol * y_ref.abs().max()).nonzero()
            print(f"[{name}] fwd mismatches: {bad.shape[0]} of {y.numel()}; first: {bad[:6].tolist()}", flush=True)
            print(f"[{name}] sample y={y.flatten()[:8].tolist()} ref={y_ref.flatten()[:8].tolist()}", flush=True)
    dy = torch.randn_like(y_ref)
    if which in ("dgrad", "wgrad", "all") and (picks[1] == ops.ALGO_TCGEN05 or picks[2] == ops.ALGO_TCGEN05):
        dx_ref, dw_ref, _ = ops.qconv2d_bwd(dy, x, w, *args, True, True, False, ops.ALGO_DIRECT)
        dx, dw, _ = ops.qconv2d_bwd(dy, x, w, *args, True, True, False, ops.ALGO_AUTO)
        torch.cuda.synchronize()
        e = rel(dx, dx_ref)
        print(f"[{name}] dgrad rel={e:.3e} tol={2 * tol}", flush=True)
        ok &= e <= 2 * tol
        ew = max(rel(a, r) for a, r in zip(dw, dw_ref))
        print(f"[{name}] wgrad rel={ew:.3e} tol={2 * tol}", flush=True)
        ok &= ew <= 2 * tol
    print(f"[{name}] {'OK' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
