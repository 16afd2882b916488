#!/usr/bin/env python
"""HBM bandwidth of a write-only stream next to the copy figure MEASURED_PEAKS.json quotes (the
float32-observation mode writes 4 B per cell and reads 0.25): torch fill_ / zero_ / copy_ on 4 GiB."""
import json
import torch


def rate(fn, nbytes, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return nbytes / best / 1e6


def main():
    n = 1 << 30
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n // 2, dtype=torch.float32, device="cuda")
    z = torch.empty(n // 2, dtype=torch.float32, device="cuda")
    out = {"fill_f32_4GiB_GBs": rate(lambda: x.fill_(1.0), 4 * n),
           "zero_4GiB_GBs": rate(lambda: x.zero_(), 4 * n),
           "copy_2GiB_read_plus_write_GBs": rate(lambda: z.copy_(y), 4 * n),
           "read_sum_4GiB_GBs": rate(lambda: x.sum(), 4 * n)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
