"""Per-K-block clock64 timeline of the leader CTA of pair 0 of the CTA-pair GEMM (vitk_debug_timeline stamps):
when the producer got a free stage, when the MMA warp saw it full, when the MMAs were issued, and when the
epilogue started / finished each tile.  Usage: python tools/gemm_timeline.py [case-substring ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
# the stamps are compiled out of the production library: use the diagnostic build (tools/build_variants.sh stamps "-DVITK_GEMM_STAMPS=1")
_stamps = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "chest-x-ray-vit_b200", "csrc", "build", "variants", "libvitk_stamps.so")
if "VITK_LIB" not in os.environ:
    if not os.path.exists(_stamps):
        sys.exit("build the diagnostic library first: tools/build_variants.sh stamps \"-DVITK_GEMM_STAMPS=1\"")
    os.environ["VITK_LIB"] = os.path.abspath(_stamps)
sys.argv, argv = sys.argv[:1] + ["--none"], sys.argv[1:]
import bench_gemm as bg  # noqa: E402  (builds CASES; "--none" keeps it from running its table)

pkg, ops = bg.pkg, bg.ops
lib = pkg._lib.lib()


KW = {}
for a_ in argv:
    if a_.startswith("--tile_n="):
        KW["tile_n"] = int(a_.split("=")[1])


def timeline(c, flush):
    tl = torch.zeros(8192, dtype=torch.int64, device="cuda")

    def go():
        ops.gemm(c["a"], c["b"], c["M"], c["N"], c["K"], c["d"], epilogue=c["epi"], a_mn_major=c["a_mn"], b_mn_major=c["b_mn"],
                 **c["extra"], **KW)
    go()
    if flush:
        bg.flush.fill_(1)
    torch.cuda.synchronize()
    lib.vitk_debug_timeline(tl.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cs = torch.cuda.current_stream().cuda_stream
    torch.cuda._sleep(1_000_000)
    if flush:
        bg.flush.fill_(1)
    lib.vitk_debug_stamp(7000, cs)
    e0.record()
    go()
    e1.record()
    lib.vitk_debug_stamp(7001, cs)
    lib.vitk_debug_stamp(7002, cs)     # back-to-back stamps: the launch-to-launch gap of trivial kernels
    lib.vitk_debug_stamp(7003, cs)
    torch.cuda.synchronize()
    lib.vitk_debug_timeline(None)
    out = tl.cpu().tolist()
    out.append(e0.elapsed_time(e1) * 1e3)
    return out


def chain(c, n=10):
    """n identical launches back to back (PDL chain, like the training step): where the device time of a launch goes."""
    tl = torch.zeros(8192, dtype=torch.int64, device="cuda")
    big = torch.iinfo(torch.int64).max
    for i in range(n + 1):
        tl[7100 + 4 * i] = big
        tl[7101 + 4 * i] = big

    def go():
        ops.gemm(c["a"], c["b"], c["M"], c["N"], c["K"], c["d"], epilogue=c["epi"], a_mn_major=c["a_mn"], b_mn_major=c["b_mn"],
                 **c["extra"], **KW)
    go()
    torch.cuda.synchronize()
    lib.vitk_debug_timeline(tl.data_ptr())
    cs = torch.cuda.current_stream().cuda_stream
    torch.cuda._sleep(1_000_000)
    lib.vitk_debug_stamp(7000, cs)
    for _ in range(n):
        go()
    lib.vitk_debug_stamp(7001, cs)
    torch.cuda.synchronize()
    lib.vitk_debug_timeline(None)
    t = tl.cpu().tolist()
    L = [(t[7100 + 4 * i], t[7101 + 4 * i], t[7102 + 4 * i], t[7103 + 4 * i]) for i in range(n)]
    per = (t[7001] - t[7000]) / 1e3 / n
    mid = L[2:-1]
    gap = sum(L[i][0] - L[i - 1][3] for i in range(3, n - 1)) / max(len(mid) - 1, 1) / 1e3
    pro = sum(x[1] - x[0] for x in mid) / len(mid) / 1e3
    busy = sum(x[2] - x[1] for x in mid) / len(mid) / 1e3
    tear = sum(x[3] - x[2] for x in mid) / len(mid) / 1e3
    print(f"=== {c['name'].strip()}: {n} launches back to back: {per:.1f} us per launch = previous kernel's last exit → first CTA entry {gap:.1f}"
          f" + entry → work start (prologue, dependency wait) {pro:.1f} + work {busy:.1f} + work end → last exit {tear:.1f} us")


def report(c, flush):
    t = timeline(c, flush)
    P = [t[4 * i] for i in range(1024) if t[4 * i]]
    F = [t[4 * i + 1] for i in range(1024) if t[4 * i + 1]]
    I = [t[4 * i + 2] for i in range(1024) if t[4 * i + 2]]
    E0 = [t[4096 + 4 * i] for i in range(256) if t[4096 + 4 * i]]
    E1 = [t[4096 + 4 * i + 1] for i in range(256) if t[4096 + 4 * i + 1]]
    n = min(len(P), len(F), len(I))
    t0 = P[0]
    starts = [t[6000 + 2 * c] for c in range(148) if t[6000 + 2 * c]]
    ends = [t[6001 + 2 * c] for c in range(148) if t[6001 + 2 * c]]
    if starts and ends:
        s0 = min(starts)
        durs = sorted(e - s for s, e in zip(starts, ends))
        print(f"   CTAs {len(starts)}: start spread {(max(starts) - s0) / 1e3:.1f} us, first start → last end {(max(ends) - s0) / 1e3:.1f} us, "
              f"per-CTA busy min/median/max {durs[0] / 1e3:.1f}/{durs[len(durs) // 2] / 1e3:.1f}/{durs[-1] / 1e3:.1f} us, "
              f"CTA 0 busy {(t[6001] - t[6000]) / 1e3:.1f} us; event time {t[-1]:.1f} us")
        exits = [t[6600 + c] for c in range(148) if t[6600 + c]]
        if exits:
            print(f"   last CTA work end → last TMEM hand-back (just before exit) {(max(exits) - max(ends)) / 1e3:.2f} us; that → next kernel's stamp {(t[7001] - max(exits)) / 1e3:.2f} us")
        entries = [t[6400 + c] for c in range(148) if t[6400 + c]]
        print(f"   device clock: stamp before → first CTA entry {(min(entries) - t[7000]) / 1e3:.1f} us, entry → start (prologue) "
              f"{(s0 - min(entries)) / 1e3:.1f} us, last end → stamp after {(t[7001] - max(ends)) / 1e3:.1f} us, stamp to stamp {(t[7001] - t[7000]) / 1e3:.1f} us "
              f"(trivial kernel → trivial kernel: {(t[7002] - t[7001]) / 1e3:.1f}, {(t[7003] - t[7002]) / 1e3:.1f} us)")
    print(f"=== {c['name'].strip()}  M={c['M']} N={c['N']} K={c['K']}  L2 {'flushed' if flush else 'warm'}: {n} K blocks, {len(E0)} tiles, "
          f"span {max(E1) - t0 if E1 else 0} cycles")
    kb_per_tile = n // max(len(E0), 1)
    cad = [F[i + 1] - F[i] for i in range(n - 1)]
    lat = [F[i] - P[i] for i in range(n)]
    iss = [I[i] - F[i] for i in range(n)]
    print(f"   first stage: request→full {lat[0]}   cadence (full→full) median {sorted(cad)[len(cad) // 2]}  mean {sum(cad) / len(cad):.0f}  "
          f"max {max(cad)}   issue median {sorted(iss)[len(iss) // 2]}")
    print(f"   request→full by K block (first 14): {lat[:14]}")
    print(f"   request→full median {sorted(lat)[len(lat) // 2]}  producer stamp gap median {sorted(P[i + 1] - P[i] for i in range(n - 1))[(n - 1) // 2]}")
    for ti in range(len(E0)):
        kb0 = ti * kb_per_tile
        hi = min(kb0 + kb_per_tile, n)
        tc = sorted(F[i + 1] - F[i] for i in range(kb0, hi - 1))
        tis = sorted(I[i] - F[i] for i in range(kb0, hi))
        tl_ = sorted(F[i] - P[i] for i in range(kb0, hi))
        print(f"   tile {ti}: cadence median {tc[len(tc) // 2] if tc else 0}  issue median {tis[len(tis) // 2]}  request→full median {tl_[len(tl_) // 2]}")
        print(f"   tile {ti}: first K block full {F[kb0] - t0:7d}  last MMA issued {I[min(kb0 + kb_per_tile, n) - 1] - t0:7d}  "
              f"epilogue start {E0[ti] - t0:7d}  end {E1[ti] - t0:7d}  (epilogue {E1[ti] - E0[ti]})")
    if "-v" in argv:
        for i in range(n):
            print(f"      kb {i:3d}  P {P[i] - t0:7d}  F {F[i] - t0:7d}  I {I[i] - t0:7d}")


sel = [a for a in argv if not a.startswith("-")]
for c in bg.CASES:
    if sel and not any(s in c["name"] for s in sel):
        continue
    if "--chain" in argv:
        chain(c)
        continue
    for flush in (True, False):
        report(c, flush)
