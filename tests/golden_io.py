"""Loader for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights(npz):
    sd = {}
    for k in npz.files:
        if k.startswith("w/"):
            a = npz[k]
            sd[k[2:]] = torch.from_numpy(a.astype(np.float32)) if a.dtype == np.float16 else torch.from_numpy(a)
    return sd


def load_depthnet(mode: str):
    """Returns (state_dict, [stage dicts]) for mode 'adaptive' or 'variance'."""
    base = np.load(os.path.join(GOLDEN, "depthnet_adaptive.npz"))
    sd = _weights(base)
    npz = base
    if mode == "variance":
        npz = np.load(os.path.join(GOLDEN, "depthnet_variance.npz"))
        sd = {k: v for k, v in sd.items() if not k.startswith("DepthNet.")}
        sd.update(_weights(npz))
    stages = []
    for s in range(3):
        p = f"s{s}/"
        st = {k[len(p):]: torch.from_numpy(npz[k].astype(np.float32) if npz[k].dtype == np.float16 else npz[k])
              for k in npz.files if k.startswith(p)}
        st["features"] = list(st["features"].unbind(0))
        stages.append(st)
    return sd, stages


def load_homo():
    npz = np.load(os.path.join(GOLDEN, "homo_warping.npz"))
    return {k: torch.from_numpy(npz[k].astype(np.float32) if npz[k].dtype == np.float16 else npz[k]) for k in npz.files}


def load_train_grads(mode: str, bn_mode: str):
    """Training-path fixture (tests/golden/make_golden_train.py): inputs, weights, loss weights, the reference's
    outputs, gradients and post-step BatchNorm buffers for mode in {adaptive, variance} x bn_mode in {train, eval}."""
    npz = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    t = lambda a: torch.from_numpy(np.asarray(a))
    tag = f"{mode}/{bn_mode}/"
    out = {"features": list(t(npz["features"]).unbind(0)), "proj": t(npz["proj"]), "depth_values": t(npz["depth_values"]),
           "r_depth": t(npz["r_depth"]), "r_prob": t(npz["r_prob"]), "r_var": t(npz["r_var"]),
           "loss": float(npz[tag + "loss"]), "g_features": list(t(npz[tag + "g_features"]).unbind(0)),
           "sd": {}, "grads": {}, "buffers": {}, "out": {}}
    for k in npz.files:
        if k.startswith(f"{mode}/w/"):
            out["sd"][k[len(mode) + 3:]] = t(npz[k])
        elif k.startswith(tag + "g/"):
            out["grads"][k[len(tag) + 2:]] = t(npz[k])
        elif k.startswith(tag + "buf/"):
            out["buffers"][k[len(tag) + 4:]] = t(npz[k])
        elif k.startswith(tag + "out/"):
            out["out"][k[len(tag) + 4:]] = t(npz[k])
    return out


def load_hypotheses():
    """Hypothesis-sampling fixture (tests/golden/make_golden_hyp.py)."""
    npz = np.load(os.path.join(GOLDEN, "hypotheses.npz"))
    return {k: torch.from_numpy(np.asarray(npz[k])) for k in npz.files}


def load_fusion_filter():
    """Geometric-consistency filter fixture (tests/golden/make_golden_filter.py), numpy arrays keyed 'a/...', 'b/...'."""
    npz = np.load(os.path.join(GOLDEN, "fusion_filter.npz"))
    return {k: np.asarray(npz[k]) for k in npz.files}


def load_cross_view_loss():
    """Cross-view photometric loss fixture (tests/golden/make_golden_cvl.py)."""
    npz = np.load(os.path.join(GOLDEN, "cross_view_loss.npz"))
    out = {"imgs": torch.from_numpy(npz["imgs"]), "loss": float(npz["loss"]), "dlossw": [float(x) for x in npz["dlossw"]],
           "cams": {}, "depth_est": {}, "depth_gt": {}, "grad": {}}
    for s in (1, 2, 3):
        k = f"stage{s}"
        for f in ("cams", "depth_est", "depth_gt", "grad"):
            out[f][k] = torch.from_numpy(npz[f"{k}/{f}"])
    return out
