"""Measured bf16-vs-reference error of the native path on the golden fixture (for DESIGN.md / test bounds)."""
import sys
import torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import synthetic
from tests import golden_io
dev = torch.device("cuda:0")
for mode in ("adaptive", "variance"):
    sd, stages = golden_io.load_depthnet(mode)
    for stage in range(3):
        st = stages[stage]
        cin = synthetic.STAGE_CHANNELS[stage]
        net = dm.DepthNet(mode, list(synthetic.STAGE_CHANNELS)).eval()
        cr = dm.CostRegNet(cin, 8).eval()
        if mode == "adaptive":
            net.load_state_dict({k[len("DepthNet."):]: v for k, v in sd.items() if k.startswith("DepthNet.")})
        pre = f"cost_regularization.{stage}."
        cr.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        net, cr = net.to(dev), cr.to(dev)
        for prec, impl in (("fp32", "auto"), ("bf16", "direct"), ("bf16", "tcgen05")):
            with dm.precision(prec, impl), torch.no_grad():
                out = net(stage, [f.to(dev) for f in st["features"]], st["proj"].to(dev), st["depth_values"].to(dev),
                          st["depth_values"].shape[1], cr)
            d = out["depth"].cpu()
            err = (d - st["depth"]).abs()
            rel = err / st["depth"].abs().clamp_min(1.0)
            span = (st["depth_values"].max(1).values - st["depth_values"].min(1).values).clamp_min(1e-3)
            nrm = err / span
            pe = (out["prob_volume"].cpu() - st["prob_volume"]).abs().max().item()
            ce = (out["photometric_confidence"].cpu() - st["photometric_confidence"]).abs()
            q = lambda t, p: torch.quantile(t.flatten(), p).item()
            print(f"{mode:9s} stage{stage + 1} {prec}/{impl:8s} depth rel med {rel.median():.2e} p99 {q(rel, .99):.2e} max {rel.max():.2e} | "
                  f"err/span med {nrm.median():.2e} p99 {q(nrm, .99):.2e} max {nrm.max():.2e} | prob max {pe:.2e} | conf med {ce.median():.2e} p99 {q(ce, .99):.2e}"
                  f" | span med {span.median():.1f}")
