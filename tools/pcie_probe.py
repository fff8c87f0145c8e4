import torch, time
d = torch.empty(64<<20, dtype=torch.uint8, device='cuda')
h = torch.empty(64<<20, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
for size in (128<<10, 512<<10, 1<<20, 4<<20, 22<<20, 64<<20):
    for direction in ('d2h','h2d'):
        torch.cuda.synchronize()
        reps = max(4, (256<<20)//size)
        t0=time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                if direction=='d2h': h[:size].copy_(d[:size], non_blocking=True)
                else: d[:size].copy_(h[:size], non_blocking=True)
        s.synchronize()
        dt=time.perf_counter()-t0
        print(direction, size>>10, 'KiB', f'{size*reps/dt/1e9:.1f} GB/s')
# bidirectional
s2=torch.cuda.Stream(); h2=torch.empty(64<<20, dtype=torch.uint8).pin_memory(); d2=torch.empty(64<<20, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(8):
    with torch.cuda.stream(s): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2[:12<<20].copy_(h2[:12<<20], non_blocking=True)
s.synchronize(); s2.synchronize(); dt=time.perf_counter()-t0
print('bidir d2h', 64*8/1024/dt*1.0737, 'GB/s', 'h2d', 12*8/1024/dt*1.0737)
