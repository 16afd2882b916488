"""Free-run (zero-action) generations/s of the warp-resident kernels at several shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, carle_b200
for n, size, rule in ((4096, 128, "B3/S23"), (16384, 256, "B3/S23"), (16384, 256, "B368/S245"),
                      (131072, 64, "B3/S23"), (8192, 192, "B3/S23"), (16384, 256, "B37/S23")):
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=32, action_height=32,
                           obs_mode="packed")
    env.rules_from_string(rule)
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device="cuda") < 0.4).float()
    k = 64
    env.step_many(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        env.step_many(k)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"free-run {n} x {size}x{size} {rule}: {n*size*size*k/(ms*1e-3):.3e} cell-updates/s "
          f"({ms*1e3/k:.2f} us/gen)")
    del env
