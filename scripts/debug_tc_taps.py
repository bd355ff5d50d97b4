"""Debug aid: per-tap check of the tcgen05 conv (one-hot weights) against torch."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import ops

dev = torch.device("cuda:0")
cin, cout, stride, transposed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ext = tuple(int(v) for v in sys.argv[5].split(","))
g = torch.Generator().manual_seed(0)
x = torch.randn(1, cin, *ext, generator=g).bfloat16().float()
vol = dm.G8Volume.from_ncdhw(x.to(dev), torch.bfloat16)
for tap in range(27):
    for ci, co in ((0, 0), (cin - 1, cout - 1)):
        w = torch.zeros((cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3))
        kd, kh, kw = tap // 9, (tap // 3) % 3, tap % 3
        if transposed:
            w[ci, co, kd, kh, kw] = 1.0
            want = F.conv_transpose3d(x, w, None, stride=2, padding=1, output_padding=1)
        else:
            w[co, ci, kd, kh, kw] = 1.0
            want = F.conv3d(x, w, None, stride=stride, padding=1)
        packed = ops.conv3d_pack_weight(w.to(dev), cin, cout, bool(transposed), ops.CONV_TCGEN05, stride)
        got = ops.conv3d(vol, packed, None, None, cout, stride, bool(transposed), False, None, torch.bfloat16, False,
                         ops.CONV_TCGEN05).to_ncdhw().cpu()
        err = (got - want).abs()
        bad = (err > 1e-2)
        if bad.any():
            idx = bad.nonzero()
            print(f"tap {tap} (kd{kd} kh{kh} kw{kw}) ci{ci} co{co}: {bad.sum().item()} bad of {bad.numel()}, first {idx[0].tolist()}, "
                  f"z-set {sorted(set(idx[:,2].tolist()))[:8]} y-set {sorted(set(idx[:,3].tolist()))[:8]} x-set {sorted(set(idx[:,4].tolist()))[:8]}")
        else:
            print(f"tap {tap} ci{ci} co{co}: ok")
