import sys
sys.path.insert(0,"aind-exaspim-image-compression_b200"); sys.path.insert(0,".")
import torch, b4d
dev=torch.device("cuda",0)
g=torch.Generator(device=dev); g.manual_seed(1)
vol=(100+torch.randint(0,60,(1024,1024,1024),device=dev,dtype=torch.int32,generator=g)).to(torch.uint16)
dn=b4d.Denoiser(0, b4d.BM4DProfile(stage_arg=None) if False else None)
y=dn.denoise(vol, 24.0); print({k:round(v[0],2) for k,v in dn.last_timings().items()}, dn.last_match_stats())
