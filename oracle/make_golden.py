"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES on seeded inputs.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so these fixtures are
the pin: outputs of openset_imagenet/{losses,metrics,util}.py loaded unmodified by
``oracle/ref_loader.py``, of ``torch.nn.CrossEntropyLoss`` configured as train.py:343,347
configures it, and of the nested ``find_nearest`` helper extracted from
script/plot_all.py:346-350 by AST (it cannot be imported).  Inputs are stored next to outputs
so the fixtures are self-contained on the GPU box, where the reference is absent.
"""
from __future__ import annotations

import ast
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O          # noqa: E402
from oracle import ref_loader           # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _labels(rng, B, C, neg_frac, neg_labels=(-1,)):
    y = rng.integers(0, C, size=B).astype(np.int64)
    neg = rng.random(B) < neg_frac
    y[neg] = rng.choice(np.asarray(neg_labels), size=int(neg.sum()))
    return y


def golden_losses():
    L = ref_loader.load("losses")
    out = {}
    cases = [  # name, B, C, w, neg_frac, neg labels, scale
        ("cfg1", 64, 116, 1.0, 1 / 3, (-1,), 1.0),
        ("w05", 48, 30, 0.5, 0.4, (-1, -2), 3.0),
        ("allneg", 16, 116, 1.0, 1.1, (-1,), 1.0),
        ("allknown", 16, 10, 1.0, 0.0, (-1,), 5.0),
        ("single", 1, 7, 2.0, 0.0, (-1,), 1.0),
        ("p3", 96, 151, 1.0, 0.3, (-1,), 2.0),
    ]
    for i, (name, B, C, w, nf, nl, scale) in enumerate(cases):
        rng = np.random.default_rng(100 + i)
        z = (scale * rng.standard_normal((B, C))).astype(np.float32)
        y = _labels(rng, B, C, nf, nl)
        zt = torch.tensor(z, requires_grad=True)
        loss = L.EntropicOpensetLoss(C, w)(zt, torch.tensor(y))
        loss.backward()
        out[f"eos_{name}_logits"] = z
        out[f"eos_{name}_y"] = y
        out[f"eos_{name}_w"] = np.float64(w)
        out[f"eos_{name}_loss"] = np.float32(loss.item())
        out[f"eos_{name}_grad"] = zt.grad.numpy().copy()
    # float64 run of the reference for a tight closed-form check
    rng = np.random.default_rng(7)
    z = rng.standard_normal((32, 116))
    y = _labels(rng, 32, 116, 0.35, (-1, -2))
    # the reference builds float32 targets; run it in float64 by promoting its buffers
    lf = L.EntropicOpensetLoss(116, 0.7)
    lf.eye, lf.ones = lf.eye.double(), lf.ones.double()
    zt = torch.tensor(z, requires_grad=True)
    tgt = torch.zeros(z.shape, dtype=torch.float64)
    kn = torch.tensor(y >= 0)
    tgt[kn] = lf.eye[torch.tensor(y)[kn]]
    tgt[~kn] = lf.ones
    loss = lf.cross_entropy(zt, tgt)
    loss.backward()
    out["eos_f64_logits"], out["eos_f64_y"], out["eos_f64_w"] = z, y, np.float64(0.7)
    out["eos_f64_loss"], out["eos_f64_grad"] = np.float64(loss.item()), zt.grad.numpy().copy()

    # softmax variant, train.py:343 ; garbage variant, train.py:346-347 + dataset.py:77-86
    rng = np.random.default_rng(11)
    z = (2 * rng.standard_normal((80, 116))).astype(np.float32)
    y = _labels(rng, 80, 116, 0.3, (-1,))
    zt = torch.tensor(z, requires_grad=True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-1)(zt, torch.tensor(y))
    loss.backward()
    out["sm_logits"], out["sm_y"] = z, y
    out["sm_loss"], out["sm_grad"] = np.float32(loss.item()), zt.grad.numpy().copy()
    zt = torch.tensor(z[:8], requires_grad=True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-1)(zt, torch.full((8,), -1, dtype=torch.long))
    out["sm_allignored_loss"] = np.float32(loss.item())          # NaN

    rng = np.random.default_rng(12)
    C1 = 152
    train_labels = _labels(rng, 5000, C1 - 1, 0.3, (C1 - 1,))   # -1 already replaced by C (dataset.py:60-68)
    train_labels[:C1] = np.arange(C1)                            # every class present
    cw = O.class_weights_from_labels(train_labels, C1)
    z = (2 * rng.standard_normal((96, C1))).astype(np.float32)
    y = _labels(rng, 96, C1 - 1, 0.3, (C1 - 1,))
    zt = torch.tensor(z, requires_grad=True)
    loss = torch.nn.CrossEntropyLoss(weight=torch.tensor(cw))(zt, torch.tensor(y))
    loss.backward()
    out["bg_logits"], out["bg_y"], out["bg_cw"] = z, y, cw
    out["bg_loss"], out["bg_grad"] = np.float32(loss.item()), zt.grad.numpy().copy()
    out["bg_train_labels"] = train_labels

    # linear -> EOS -> backward in float64 (model.py:23-26,37-38 + losses.py)
    rng = np.random.default_rng(13)
    B, K, C = 40, 96, 116
    x = O.bf16_round(rng.standard_normal((B, K)).astype(np.float32)).astype(np.float64)
    W = O.bf16_round((rng.standard_normal((C, K)) / np.sqrt(K)).astype(np.float32)).astype(np.float64)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32).astype(np.float64)
    y = _labels(rng, B, C, 1 / 3, (-1,))
    lin = torch.nn.Linear(K, C, bias=True).double()
    with torch.no_grad():
        lin.weight.copy_(torch.tensor(W))
        lin.bias.copy_(torch.tensor(b))
    xt = torch.tensor(x, requires_grad=True)
    lf = L.EntropicOpensetLoss(C, 1.0)
    lf.eye, lf.ones = lf.eye.double(), lf.ones.double()
    zt = lin(xt)
    tgt = torch.zeros(zt.shape, dtype=torch.float64)
    kn = torch.tensor(y >= 0)
    tgt[kn] = lf.eye[torch.tensor(y)[kn]]
    tgt[~kn] = lf.ones
    loss = lf.cross_entropy(zt, tgt)
    loss.backward()
    out.update(head_x=x, head_W=W, head_b=b, head_y=y, head_loss=np.float64(loss.item()),
               head_dx=xt.grad.numpy().copy(), head_dw=lin.weight.grad.numpy().copy(),
               head_db=lin.bias.grad.numpy().copy())
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    return out


def golden_confidence():
    M = ref_loader.load("metrics")
    out = {}
    cases = [  # name, N, C, offset, unknown_class, last_valid_class, neg labels, neg frac
        ("eos", 64, 116, 1 / 116, -1, None, (-1,), 1 / 3),
        ("eos_mixed", 200, 30, 1 / 30, -1, None, (-1, -2), 0.5),
        ("table", 200, 30, 1 / 30, -2, None, (-1, -2), 0.5),
        ("bg", 150, 31, 0.0, 30, -1, (30,), 0.3),
        ("noneg", 8, 5, 0.2, -1, None, (-1,), 0.0),
        ("noknown", 9, 5, 0.2, -1, None, (-1,), 1.1),
    ]
    for i, (name, N, C, off, uc, lvc, nl, nf) in enumerate(cases):
        rng = np.random.default_rng(200 + i)
        s = torch.softmax(torch.tensor((2 * rng.standard_normal((N, C))).astype(np.float32)), dim=1).numpy()
        y = _labels(rng, N, C if uc < 0 else C - 1, nf, nl)
        r = M.confidence(torch.tensor(s), torch.tensor(y), offset=off, unknown_class=uc, last_valid_class=lvc)
        assert isinstance(r[1], int) and isinstance(r[3], int)
        out[f"{name}_scores"], out[f"{name}_y"] = s, y
        out[f"{name}_args"] = np.array([off, uc, np.nan if lvc is None else lvc], dtype=np.float64)
        out[f"{name}_out"] = np.array([r[0], r[1], r[2], r[3]], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "confidence.npz"), **out)
    return out


def _extract_find_nearest():
    """Compile the nested helper at script/plot_all.py:346-350 without importing the script."""
    path = os.path.join(ref_loader.REF_ROOT, "openset_imagenet", "script", "plot_all.py")
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "find_nearest":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"numpy": np}
            exec(compile(mod, path, "exec"), ns)
            return ns["find_nearest"]
    raise RuntimeError("find_nearest not found")


def golden_oscr():
    U = ref_loader.load("util")
    find_nearest = _extract_find_nearest()
    out = {}

    def run(name, gt, scores, unk_label):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ccr, fpr = U.calculate_oscr(gt, scores, unk_label=unk_label)
        out[f"{name}_gt"], out[f"{name}_scores"] = gt, scores
        out[f"{name}_unk"] = np.int64(unk_label)
        out[f"{name}_ccr"], out[f"{name}_fpr"] = np.asarray(ccr, dtype=np.float64), np.asarray(fpr, dtype=np.float64)
        return ccr, fpr

    gt, s = O.make_oscr_case(1500, 24, seed=3)
    for unk in (-1, -2):
        ccr, fpr = run(f"synth_m{-unk}", gt.astype(np.float32), s, unk)   # gt stored float32 as in the npz (train.py:212)
        # CCR@FPR lookup on the reference's own curve (plot_all.py:352,378-384)
        res = []
        for q in [1e-3, 1e-2, 0.1, 1.0]:
            idx, f = find_nearest(fpr, q)
            err = round(100 * abs(f - q) / q, 1)
            res.append([q, idx, f, np.nan if err >= 10.0 else ccr[idx]])
        out[f"synth_m{-unk}_ccr_at_fpr"] = np.array(res, dtype=np.float64)
    # heavy ties: quantised scores, C = 116 (P1 width)
    rng = np.random.default_rng(5)
    gt = O.make_labels(400, 116, 0.5, seed=6)
    gt[rng.random(400) < 0.2] = -2
    s = (rng.integers(0, 6, size=(400, 116)) / 8.0).astype(np.float32)
    run("ties", gt, s, -1)
    # BG-style: scores with the garbage column kept (plot_all.py:359-363 does not drop it)
    gt, s = O.make_oscr_case(600, 31, seed=9)
    run("bgcol", gt, s, -2)
    # known-answer vectors of SURVEY.md §9.6
    s6 = np.array([[.5, .5], [.2, .8], [.6, .4], [.7, .3], [.9, .1], [.8, .2]], dtype=np.float32)
    run("kat_ties", np.array([0, 1, -1, 1, -2, 0]), s6, -1)
    run("kat_nounk", np.array([0, 1, 1, 0]), s6[:4], -1)
    run("kat_oneknown", np.array([0, -1, -1]), s6[:3], -1)
    run("kat_noknown", np.array([-1, -1, -2]), s6[:3], -1)
    np.savez_compressed(os.path.join(OUT, "oscr.npz"), **out)
    return out


def golden_hist():
    """util.get_histogram (util.py:202-228) run from the reference on seeded arrays."""
    U = ref_loader.load("util")
    out = {}
    rng = np.random.default_rng(11)
    gt, s = O.make_oscr_case(1500, 24, seed=4)
    feats = (rng.standard_normal((1500, 16)) * rng.uniform(0.5, 6.0, size=(1500, 1))).astype(np.float32)
    arr = {"scores": s, "gt": gt.astype(np.float32), "features": feats}       # gt stored float32 as in the npz (train.py:212)
    out["gt"], out["scores"], out["features"] = arr["gt"], s, feats
    for name, kw in HIST_CASES.items():
        a = dict(arr)
        if kw.get("drop_bg"):
            a["gt"] = np.where(arr["gt"] >= 23, 0, arr["gt"]).astype(np.float32)     # labels must fit the 23 columns left
        kh, ke, uh, ue = U.get_histogram(a, **kw)
        out[f"{name}_kn_hist"], out[f"{name}_kn_edges"] = kh, ke
        out[f"{name}_unk_hist"], out[f"{name}_unk_edges"] = uh, ue
    np.savez_compressed(os.path.join(OUT, "hist.npz"), **out)
    return out


HIST_CASES = {
    "score_m1": dict(unk_label=-1, metric="score", bins=100),
    "score_m2_b37": dict(unk_label=-2, metric="score", bins=37),
    "score_dropbg": dict(unk_label=-1, metric="score", bins=50, drop_bg=True),
    "score_log": dict(unk_label=-1, metric="score", bins=30, log_space=True, geomspace_limits=(1e-4, 1.0)),
    "norm_m1": dict(unk_label=-1, metric="norm", bins=100),
    "norm_log": dict(unk_label=-2, metric="norm", bins=100, log_space=True, geomspace_limits=(1, 1e2)),
}


def main():
    if not ref_loader.available():
        raise SystemExit("reference not available; golden vectors can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    a, b, c, d = golden_losses(), golden_confidence(), golden_oscr(), golden_hist()
    print("wrote", len(a), len(b), len(c), len(d), "arrays to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
