"""PCIe copy rates of the box: H2D alone, D2H alone, both directions at once (pinned host memory)."""
import torch, time, json
n = 453 << 20
h_in = torch.empty(365 << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(365 << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        for c in range(chunks):
            if h2d:
                with torch.cuda.stream(s1):
                    a = h_in.numel() // chunks
                    d_in[c*a:(c+1)*a].copy_(h_in[c*a:(c+1)*a], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    b = h_out.numel() // chunks
                    h_out[c*b:(c+1)*b].copy_(d_out[c*b:(c+1)*b], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 3 * 1e3
for _ in range(2): run(True, True)
print(json.dumps(dict(h2d_ms=run(True, False), d2h_ms=run(False, True), both_ms=run(True, True), both_6chunks_ms=run(True, True, 6),
                      h2d_GBs=(365 << 20) / run(True, False) / 1e6, d2h_GBs=n / run(False, True) / 1e6)))
