"""Host<->device copy bandwidth on this box (pinned vs pageable), CUDA-event timed."""
import torch
n = 4096 * 3200
for name, t in (("pinned", torch.rand(n).pin_memory()), ("pageable", torch.rand(n))):
    d = torch.empty(n, device="cuda")
    for _ in range(2):
        d.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        d.copy_(t, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print("H2D %s: %.2f ms for %.1f MB -> %.1f GB/s (is_pinned=%s)" % (name, ms, n * 4 / 1e6, n * 4 / ms / 1e6, t.is_pinned()))
h = torch.empty(n).pin_memory()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    h.copy_(d, non_blocking=True)
e.record(); torch.cuda.synchronize()
print("D2H pinned: %.1f GB/s" % (n * 4 / (s.elapsed_time(e) / 5) / 1e6))
