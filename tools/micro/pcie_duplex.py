"""PCIe micro-benchmark for the end-to-end path (bench.py `e2e`): pinned H2D alone, D2H alone, both at once on two
streams (is the host link full duplex on this box?), the same with chunked copies, and zero-copy kernel stores into mapped
pinned memory.  Prints one JSON object.  Run on a GPU box: `python tools/micro/pcie_duplex.py`."""
import json
import sys

import torch


def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for mb in (33, 66):
        n = mb * 1000 * 1000
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        d_in = torch.empty(n, dtype=torch.uint8, device=dev)
        d_out = torch.zeros(n, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream()

        def h2d():
            d_in.copy_(h_in, non_blocking=True)

        def d2h():
            h_out.copy_(d_out, non_blocking=True)

        def both():
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)

        def both_chunked(chunks=8):
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            step = n // chunks
            for c in range(chunks):
                with torch.cuda.stream(s1):
                    d_in[c * step:(c + 1) * step].copy_(h_in[c * step:(c + 1) * step], non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out[c * step:(c + 1) * step].copy_(d_out[c * step:(c + 1) * step], non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)

        t_h2d, t_d2h, t_both, t_chunk = timed(h2d), timed(d2h), timed(both), timed(both_chunked)
        out["%dMB" % mb] = {
            "h2d_gbs": n / t_h2d / 1e6, "d2h_gbs": n / t_d2h / 1e6,
            "duplex_total_gbs": 2 * n / t_both / 1e6, "duplex_ms": t_both, "serial_ms": t_h2d + t_d2h,
            "duplex_chunked_total_gbs": 2 * n / t_chunk / 1e6,
        }
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
