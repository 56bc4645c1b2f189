"""Per-shape timing of the model's GEMMs (ViT-B/16@384, batch 16 → M = 9232) through the C ABI.
CUDA events around each launch, L2 flushed (256 MB write) between launches; prints TFLOP/s and
the fraction of the measured burst peak.  Usage: python tools/bench_gemm.py [variants...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = "cuda"
bf16 = torch.bfloat16
M, D, F = (int(os.environ.get(k, v)) for k, v in (("VITK_BG_M", 9232), ("VITK_BG_D", 768), ("VITK_BG_F", 3072)))   # ViT-L at batch 8: 4616 1024 4096
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1590.0

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(n, *shape, dt=bf16):
    return (torch.randn(*shape, device=dev) * 0.05).to(dt)


def case(name, Mm, N, K, epi, a_mn=False, b_mn=False, **kw):
    a = t("a", K, Mm) if a_mn else t("a", Mm, K)
    b = t("b", K, N) if b_mn else t("b", N, K)
    f32_out = epi in (ops.EPI_BIAS_RESID_F32, ops.EPI_ACCUM_F32, ops.EPI_STORE_F32)
    d = torch.zeros(Mm, N, device=dev, dtype=torch.float32 if f32_out else bf16)
    extra = {}
    if epi in (ops.EPI_BIAS_BF16, ops.EPI_BIAS_GELUG_BF16, ops.EPI_BIAS_RESID_F32):
        extra["bias"] = torch.randn(N, device=dev)
    if epi == ops.EPI_BIAS_GELUG_BF16:
        extra["d2"] = torch.empty(Mm, N, device=dev, dtype=bf16)
    if epi == ops.EPI_BIAS_RESID_F32:
        extra["aux"] = torch.randn(Mm, N, device=dev)
    if epi == ops.EPI_MUL_BF16:
        extra["aux"] = t("x", Mm, N)
    return dict(name=name, a=a, b=b, M=Mm, N=N, K=K, d=d, epi=epi, a_mn=a_mn, b_mn=b_mn, extra=extra)


CASES = [
    case("qkv fwd      ", M, 3 * D, D, ops.EPI_BIAS_BF16),
    case("out fwd      ", M, D, D, ops.EPI_BIAS_RESID_F32),
    case("fc1 fwd gelu ", M, F, D, ops.EPI_BIAS_GELUG_BF16),
    case("fc2 fwd      ", M, D, F, ops.EPI_BIAS_RESID_F32),
    case("fc2 dgrad mul", M, F, D, ops.EPI_MUL_BF16, b_mn=True),
    case("fc1 dgrad    ", M, D, F, ops.EPI_STORE_BF16, b_mn=True),
    case("out dgrad    ", M, D, D, ops.EPI_STORE_BF16, b_mn=True),
    case("qkv dgrad    ", M, D, 3 * D, ops.EPI_STORE_BF16, b_mn=True),
    case("fc2 wgrad    ", D, F, M, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True),
    case("fc1 wgrad    ", F, D, M, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True),
    case("out wgrad    ", D, D, M, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True),
    case("qkv wgrad    ", 3 * D, D, M, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True),
]


def run(c, reps=10, **kw):
    if kw.get("cublas"):        # torch.matmul (cuBLASLt), no epilogue: the library's number for the same shape, for reference
        a2 = c["a"].t() if c["a_mn"] else c["a"]
        b2 = c["b"] if c["b_mn"] else c["b"].t()
        out = torch.empty(c["M"], c["N"], device=dev, dtype=bf16)

        def go():
            torch.matmul(a2, b2, out=out)
    else:
        go = None
    def go_vitk():
        ops.gemm(c["a"], c["b"], c["M"], c["N"], c["K"], c["d"], epilogue=c["epi"], a_mn_major=c["a_mn"], b_mn_major=c["b_mn"],
                 **c["extra"], **kw)
    go = go or go_vitk
    for _ in range(2):
        go()
    # everything is enqueued behind a ~1 ms spin kernel so the host runs ahead of the GPU: the event pair then
    # brackets device time only (with a sync per repetition the interval also contains ≈15–25 µs of host launch latency)
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)
    evs = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        go()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    ts.sort()
    return ts[len(ts) // 2]


configs = [dict(variant=1), dict(variant=2), dict(variant=2, tile_n=128), dict(variant=2, tile_n=192), dict(variant=2, tile_n=256)]
if len(sys.argv) > 1 and sys.argv[1] == "--quick":
    configs = [dict(variant=2), dict(variant=2, tile_n=256)]
if len(sys.argv) > 1 and sys.argv[1] == "--splitk":        # weight-gradient launches only, split-K factor forced
    configs = [dict(variant=2)] + [dict(variant=2, split_k=k) for k in (1, 2, 3, 4, 5, 6, 8, 10)]
    CASES = [c for c in CASES if "wgrad" in c["name"]]
if len(sys.argv) > 1 and sys.argv[1] == "--auto":
    configs = [dict(variant=0), dict(variant=1)]
if len(sys.argv) > 1 and sys.argv[1] == "--ew":          # 16-warp heavy epilogues (default) vs the 8-warp kernel (variant 3)
    configs = [dict(variant=2), dict(variant=3)]
if len(sys.argv) > 1 and sys.argv[1] == "--cublas":
    configs = [dict(variant=2), dict(cublas=True)]
if len(sys.argv) > 1 and sys.argv[1] == "--none":         # imported as a module (tools/gemm_timeline.py)
    configs = []
if len(sys.argv) > 1 and sys.argv[1] == "--profile":      # one launch per case, default tiling (for ncu)
    for c in CASES:
        ops.gemm(c["a"], c["b"], c["M"], c["N"], c["K"], c["d"], epilogue=c["epi"], a_mn_major=c["a_mn"], b_mn_major=c["b_mn"],
                 **c["extra"])
    torch.cuda.synchronize()
    for c in CASES:
        flush.fill_(1)
        ops.gemm(c["a"], c["b"], c["M"], c["N"], c["K"], c["d"], epilogue=c["epi"], a_mn_major=c["a_mn"], b_mn_major=c["b_mn"],
                 **c["extra"])
    torch.cuda.synchronize()
    sys.exit(0)
if not configs:
    CASES_TO_RUN = []
else:
    CASES_TO_RUN = CASES
compact = len(configs) > 4
print(f"{'gemm':14s} {'M':>5s} {'N':>5s} {'K':>5s} | " + " | ".join(f"{str(k):>24s}" if not compact else f"{str(k.get('split_k', 'auto')):>7s}" for k in configs))
tot = [0.0] * len(configs)
for c in CASES_TO_RUN:
    fl = 2.0 * c["M"] * c["N"] * c["K"]
    cells = []
    for i, k in enumerate(configs):
        try:
            ms = run(c, **k)
            cells.append(f"{ms * 1e3:7.1f}us {fl / ms / 1e9:6.0f}TF {100 * fl / ms / 1e9 / PEAK:4.0f}%" if not compact else f"{ms * 1e3:7.1f}")
            tot[i] += ms
        except RuntimeError as e:
            cells.append(f"{'n/a':>24s}")
            tot[i] += float("nan")
    print(f"{c['name']} {c['M']:5d} {c['N']:5d} {c['K']:5d} | " + " | ".join(f"{x:>24s}" if not compact else f"{x:>7s}" for x in cells))
print("sum over one layer's 12 GEMMs (ms):", ["%.3f" % x for x in tot])
