"""Small workload touching every one-launch step kernel (strips with 2 / 4 rows per lane, stream
kernel at 64x64 / 128x128, 128x128 strips, device random agent, SpeedDetector tail) with built-in
and NVRTC-specialised rules; prints a checksum per run.  A quick end-to-end launch check on a GPU
box (and the workload to put under a memory checker where one is available)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, carle_b200
from carle_b200.env import RandomAction

def run(envs, size, win, n, rule, steps=3, sums=True, random_agent=False):
    for k in ("CARLE_FUSED_IMPL", "CARLE_STRIP_R", "CARLE_PDL"):
        os.environ.pop(k, None)
    os.environ.update(envs)
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                           obs_mode="packed", fused_reductions=sums)
    env.rules_from_string(rule)
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device="cuda") < 0.4).float()
    for t in range(steps):
        if random_agent:
            env.step(RandomAction(env, 3, t, 0.1, n))
        else:
            a = (torch.rand(n, 1, win, win, device="cuda") <= 0.1)
            env.step(a.float() if t % 2 == 0 else a.to(torch.uint8))
    torch.cuda.synchronize()
    return int(env.packed_universe.sum())

for rule in ("B3/S23", "B36/S125"):
    print(rule, run({"CARLE_STRIP_R": "4"}, 256, 64, 500, rule), run({"CARLE_STRIP_R": "2"}, 256, 64, 500, rule),
          run({}, 128, 32, 3000, rule), run({}, 64, 32, 5000, rule),
          run({"CARLE_FUSED_IMPL": "strip"}, 128, 32, 700, rule),
          run({}, 128, 32, 3000, rule, random_agent=True), run({}, 256, 64, 100, rule, random_agent=True))
env = carle_b200.SpeedDetector(carle_b200.CARLE(instances=900, height=64, width=64, action_width=32,
                                                action_height=32, obs_mode="packed"))
env.reset()
for t in range(3):
    env.step(1.0 * (torch.rand(900, 1, 32, 32, device="cuda") <= 0.1))
torch.cuda.synchronize()
print("all kernels ran")
