"""Random-agent rollout entirely on the GPU (carle_step_random), K steps in a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, carle_b200
from carle_b200 import _lib as _l
lib = _l.load()
for n, size, win in ((4096, 128, 32), (131072, 64, 32), (16384, 256, 64), (1, 64, 32)):
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                           obs_mode="packed")
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device="cuda") < 0.5).float()
    env._sync_rule()
    words = env._action_buf
    def steps(k0, k):
        for i in range(k):
            rc = lib.carle_step_random(env._handle, env._packed.data_ptr(), env._spare.data_ptr(), 7,
                                       k0 + i, 0.1, n, words.data_ptr(), env._counters.data_ptr(),
                                       None, env._stream())
            assert rc == 0, _l.last_error()
            env._packed, env._spare = env._spare, env._packed
    steps(0, 4)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        steps(4, 100)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("RESULT random-agent impl=%s %d x %dx%d: %.2f us/step, %.3e cell-updates/s" % (
        os.environ.get("CARLE_RANDOM_IMPL", "stream"), n, size, size, best * 10, n * size * size * 100 / (best * 1e-3)))
