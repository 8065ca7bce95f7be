import os, time, torch
print("OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads())
dev = torch.device("cuda", 0)
h = torch.rand(4096, 1, 32, 100).pin_memory()
d = torch.empty_like(h, device=dev)
big = torch.rand(8192, 8192, device=dev)
side = torch.cuda.Stream(dev)
def busy():
    for _ in range(20):
        big @ big
for trial in range(3):
    busy()                               # ~40+ ms of GPU work on the default stream
    t0 = time.perf_counter(); p = h.is_pinned(); t1 = time.perf_counter()
    with torch.cuda.stream(side):
        d.copy_(h, non_blocking=True)
    t2 = time.perf_counter()
    ev = torch.cuda.Event(); ev.record(side); t3 = time.perf_counter()
    torch.cuda.synchronize()
    print("is_pinned %.2f ms, copy_ enqueue %.2f ms, event %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
