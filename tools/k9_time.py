"""Times K9 (chunk gather + byte shuffle, with and without the byte counts) on device-resident data of
three kinds: raw counts, denoised step-1 counts, coarsely quantized counts.  Developer tool."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import torch
import b4d
from b4d import synth
dev = torch.device("cuda", 0)
dn = b4d.Denoiser(0)
clean = torch.from_numpy(synth.clean_tile(1000)).to(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
big = clean.repeat(4, 8, 8)  # 512 x 1024 x 1024
raw = torch.clamp(torch.round(big + torch.randn(big.shape, generator=g, device=dev) * 24.0), 0, 65535).to(torch.int32).to(torch.uint16)
den = torch.clamp(torch.round(big), 0, 65535).to(torch.int32).to(torch.uint16)
coarse = torch.clamp(torch.round(big / 12.6), 0, 65535).to(torch.int32).to(torch.uint16)
ext = torch.cuda.ExternalStream(dn.stream_ptr(), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {"unit": "GB/s (4 B per voxel)"}
for name, x in (("raw", raw), ("denoised", den), ("coarse", coarse)):
    for hist in (True, False):
        best = None
        for _ in range(5):
            e0.record(ext); r = dn.chunk_shuffle(x, (64, 64, 64), True, hist) if hist else dn.chunk_shuffle(x, (64, 64, 64), True, False); e1.record(ext)
            torch.cuda.synchronize(); ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        res["%s_%s" % (name, "hist" if hist else "bytes")] = round(4.0 * x.numel() / (best * 1e-3) / 1e9)
print(json.dumps(res))
