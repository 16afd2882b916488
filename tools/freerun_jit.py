"""Free run (zero actions, 64 generations per launch) of an arbitrary rule: NVRTC-specialised vs
run-time-rule kernels (GPU box)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch, carle_b200
    n, size = 4096, 128
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=32, action_height=32,
                           obs_mode="packed")
    env.rules_from_string(sys.argv[1])
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device="cuda") < 0.5).float()
    env.step_many(64)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        env.step_many(64)
    b.record()
    torch.cuda.synchronize()
    print("RESULT free-run %s JIT=%s: %.3e cell-updates/s" % (
        sys.argv[1], os.environ.get("CARLE_JIT", "1"), n * size * size * 64 * 20 / (a.elapsed_time(b) * 1e-3)))
else:
    for rule in ("B3/S23", "B36/S125"):
        for jit in ("1", "0"):
            subprocess.run([sys.executable, __file__, rule], env=dict(os.environ, CARLE_JIT=jit))
