"""Raw PCIe rates of the box next to bench.py's e2e figure: pinned H2D alone, D2H alone, both at once (two streams),
for the step's byte counts (336 MB in, 168 MB out)."""
import json
import torch

def main():
    n_in, n_out = 336_000_000, 168_000_056
    hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def h2d():
        with torch.cuda.stream(s1):
            s1.wait_stream(torch.cuda.current_stream())
            din.copy_(hin, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            s2.wait_stream(torch.cuda.current_stream())
            hout.copy_(dout, non_blocking=True)

    def both():
        h2d()
        d2h()

    ms = timed(h2d)
    res["h2d_alone"] = {"ms": ms, "GB/s": n_in / ms / 1e6}
    ms = timed(d2h)
    res["d2h_alone"] = {"ms": ms, "GB/s": n_out / ms / 1e6}
    ms = timed(both)
    res["both"] = {"ms": ms, "h2d GB/s": n_in / ms / 1e6, "d2h GB/s": n_out / ms / 1e6,
                   "Mpixel/s if the step were copies only": 14e6 / ms / 1e3}
    print(json.dumps(res))

if __name__ == "__main__":
    main()
