import torch, time
torch.cuda.set_device(0)
for chunk in (1<<20, 4<<20, 64<<20):
    n = (4<<30)//chunk
    host = torch.empty(4<<30, dtype=torch.uint8).pin_memory()
    dev = torch.empty(4<<30, dtype=torch.uint8, device='cuda')
    ss = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(2):
        t0=time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(ss[i%2]):
                dev[i*chunk:(i+1)*chunk].copy_(host[i*chunk:(i+1)*chunk], non_blocking=True)
        torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(f"H2D pinned chunk={chunk>>20} MiB: {(4<<30)/dt/1e9:.1f} GB/s")
    del host, dev
