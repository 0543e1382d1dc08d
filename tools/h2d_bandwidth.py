"""H2D bandwidth from pinned memory: one copy vs the same bytes split over several streams."""
import torch


def main():
    n = 503326720
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    for parts in (1, 2, 4, 8):
        streams = [torch.cuda.Stream() for _ in range(parts)]
        chunk = n // parts
        best = 1e9
        for rep in range(4):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i, s in enumerate(streams):
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    dev[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"parts {parts}: {best:.2f} ms  {n / best / 1e6:.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
