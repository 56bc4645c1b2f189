"""Fine-grained host timing of DeviceFeeder's stages (why does the feeder alone cost milliseconds per batch?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = 16
xh = [torch.randn(B, 3, 384, 384).pin_memory() for _ in range(4)]
torch.cuda.init()
dev = torch.device("cuda", 0)
s = torch.cuda.Stream()
d = [torch.empty(B, 3, 384, 384, device=dev) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
freed = [torch.cuda.Event() for _ in range(2)]
def T(): return time.perf_counter()
acc = {}
def add(k, t0): acc[k] = acc.get(k, 0.0) + (T() - t0)
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize()
    N = 40
    for i in range(N):
        slot = i % 2
        t0 = T(); p = xh[i % 4].is_pinned(); add("is_pinned", t0)
        t0 = T()
        with torch.cuda.stream(s):
            add("stream ctx enter", t0)
            t0 = T(); s.wait_event(freed[slot]); add("wait_event", t0)
            t0 = T(); d[slot].copy_(xh[i % 4], non_blocking=True); add("copy_", t0)
            t0 = T(); ready[slot].record(s); add("record", t0)
            t0 = T()
        add("stream ctx exit", t0)
        t0 = T(); cur = torch.cuda.current_stream(); cur.wait_event(ready[slot]); freed[slot].record(cur); add("consumer events", t0)
    t0 = T(); torch.cuda.synchronize(); add("final sync", t0)
    print(f"rep {rep}: " + "  ".join(f"{k} {1e3 * v / N:.3f} ms" for k, v in acc.items()))
