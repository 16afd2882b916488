"""Single-GPU giant-grid throughput of the tiled family (temporal blocking), free run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, carle_b200
sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384, 65536]
for size in sizes:
    env = carle_b200.CARLE(instances=1, height=size, width=size, obs_mode="packed")
    env.reset()
    env.packed_universe.random_(-2**31, 2**31 - 1)
    k = 64
    env.step_many(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 3
    for _ in range(reps):
        env.step_many(k)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"tiled T={os.environ.get('CARLE_TILE_T','16')} {size}x{size}: "
          f"{size*size*k/(ms*1e-3):.3e} cell-updates/s ({ms*1e3/k:.1f} us/gen), "
          f"algorithmic {size*size*k*0.25/(ms*1e-3)/1e9:.0f} GB/s")
    del env
    torch.cuda.empty_cache()
