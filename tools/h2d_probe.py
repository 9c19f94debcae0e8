import torch, time
n = 56_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(k, reps=30):
    ss = [torch.cuda.Stream() for _ in range(k)]
    per = n // k
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                d[i*per:(i+1)*per].copy_(h[i*per:(i+1)*per], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9
for k in (1, 2, 4, 8):
    run(k, 3)
    print("streams", k, "%.1f GB/s" % run(k))
# chunked on one stream (8 pieces)
s = torch.cuda.Stream()
for pieces in (1, 8, 32):
    per = n // pieces
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(30):
        with torch.cuda.stream(s):
            for i in range(pieces):
                d[i*per:(i+1)*per].copy_(h[i*per:(i+1)*per], non_blocking=True)
        torch.cuda.synchronize()
    print("one stream, %d pieces: %.1f GB/s" % (pieces, n / ((time.perf_counter() - t0) / 30) / 1e9))
