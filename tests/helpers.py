"""Test infrastructure: CPU (torch fp32/fp64) execution of a compiled layer plan, used to check the graph compiler
without a GPU and to localise a wrong kernel layer by layer on the GPU; plus seeded input generators."""
import numpy as np
import torch
import torch.nn.functional as F


def run_plan_cpu(model, x_nhwc: np.ndarray, dtype=torch.float32, keep=False):
    """model: HfrModel (host-only is fine).  x_nhwc: pre-processed float input [B,H,W,3].
    Returns (list of outputs in plan order, {layer index: NHWC ndarray} if keep)."""
    plan = model.plan()
    vals = {0: torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)}
    kept = {}
    for li, L in enumerate(plan["layers"]):
        w, b = model.layer_weights(li)
        w = torch.from_numpy(w).to(dtype)
        b = torch.from_numpy(b).to(dtype) if L["has_bias"] else None
        x = vals[L["in"]]
        kind = L["kind"]
        kh, kw = L["k"]
        pt, pb, pl, pr = L["pad"]
        if kind in ("stem", "pw", "conv", "dw"):
            xp = F.pad(x, (pl, pr, pt, pb))
            if kind == "stem":
                wt = w.view(kh, kw, L["cin"], L["cout"]).permute(3, 2, 0, 1)
                y = F.conv2d(xp, wt, b, stride=L["stride"])
            elif kind == "dw":
                wt = w.view(3, 3, L["cin"]).permute(2, 0, 1).unsqueeze(1)
                y = F.conv2d(xp, wt, b, stride=L["stride"], groups=L["cin"])
            else:
                wt = w.view(L["cout"], kh, kw, L["cin"]).permute(0, 3, 1, 2)
                y = F.conv2d(xp, wt, b, stride=L["stride"])
            if L["in2"] >= 0:
                y = y + vals[L["in2"]]
        elif kind == "subsample":
            y = x[:, :, ::L["stride"], ::L["stride"]]
        elif kind == "maxpool":
            fill = 0.0 if L["explicit_zero_pad"] else -float("inf")
            y = F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=fill), (kh, kw), L["stride"])
        elif kind == "gap":
            y = x.mean(dim=(2, 3))
        elif kind == "fc":
            y = x.reshape(x.shape[0], -1) @ w.view(L["cin"], L["cout"])
            if b is not None:
                y = y + b
        else:
            raise AssertionError(kind)
        act = L["act"]
        if act == "relu":
            y = torch.relu(y)
        elif act == "relu6":
            y = torch.clamp(y, 0.0, 6.0)
        elif act == "sigmoid":
            y = torch.sigmoid(y)
        elif act == "softmax":
            y = torch.softmax(y, dim=-1)
        vals[L["out"]] = y
        if keep:
            kept[li] = (y.permute(0, 2, 3, 1) if y.dim() == 4 else y).float().numpy()
    outs = []
    for o in plan["outputs"]:
        y = vals[o["value"]]
        outs.append(y.reshape(y.shape[0], -1).float().numpy())
    return outs, kept


def smooth_images(n, size, seed):
    """Seeded smooth colour fields (SURVEY.md 8d input class ii): low-res noise upsampled bilinearly + colour offset."""
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        g = rs.randint(8, 17)
        low = torch.from_numpy(rs.rand(1, 3, g, g).astype(np.float32))
        up = F.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)[0].permute(1, 2, 0).numpy()
        img = up * 200.0 + rs.uniform(-20, 60, size=(1, 1, 3))
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return np.stack(out)


def cosine(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)


def merge_pairs_reference(d_all: np.ndarray, i_all: np.ndarray):
    """Test oracle for hfr_knn_merge: per query the smallest distance over the shards, ties -> lowest global index."""
    P, nq = d_all.shape
    best_d = np.full(nq, np.inf, np.float64)
    best_i = np.full(nq, -1, np.int64)
    for p in range(P):
        for q in range(nq):
            if i_all[p, q] < 0:
                continue
            if best_i[q] < 0 or d_all[p, q] < best_d[q] or (d_all[p, q] == best_d[q] and i_all[p, q] < best_i[q]):
                best_d[q], best_i[q] = d_all[p, q], i_all[p, q]
    return best_d, best_i
