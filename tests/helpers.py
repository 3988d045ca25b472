"""Shared test scaffolding: fake trainer / pool objects shaped like the reference's."""
import types

import numpy as np
import torch


class IdentityNet:
    """The acquisition tests hand logits over as 'images', so the network is the identity."""

    def eval(self):
        return self

    def __call__(self, x):
        return x


class PoolSet(torch.utils.data.Dataset):
    def __init__(self, logits, spx, im_idx, suppix):
        self.logits, self.spx = logits, spx
        self.im_idx = [list(k) for k in im_idx]
        self.suppix = {k: list(v) for k, v in suppix.items()}

    def __len__(self):
        return len(self.im_idx)

    def __getitem__(self, i):
        return {"images": self.logits[i], "spx": self.spx[i], "labels": self.spx[i]}


def selector_args(method, nseg, num_classes, predignore, temp, coeff, bs):
    return types.SimpleNamespace(val_batch_size=bs, val_num_workers=0, nseg=nseg, active_method=method,
                                 num_classes=num_classes, ce_temp=temp, cls_weight_coeff=coeff, save_scores=False,
                                 method="active_joint_multi_predignore_lossdecomp" if predignore
                                 else "active_joint_multi_lossdecomp")


def fake_trainer(device):
    return types.SimpleNamespace(net=IdentityNet(), device=torch.device(device), model_save_dir="/tmp", selection_iter=1)


def batches(logits, spx, bs):
    return [(logits[i:i + bs], spx[i:i + bs]) for i in range(0, logits.shape[0], bs)]


def assert_scores_close(got, ref, normalised, msg=""):
    """1e-5 relative (north_star) on the region means.  Normalised selectors then compute
    (u - min) / (max - min): a 1e-5 relative error on u becomes up to 1e-5 * |u| / (max - min) ABSOLUTE on the
    [0,1] scale (cancellation near the pool minimum), so those are held to 1e-5 absolute + 1e-5 relative."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 if normalised else 1e-12, err_msg=msg)


def class_weights_ref(prob_sum_all, pixels_per_image, ref_batch, coeff):
    """Test-side restatement of my_bvsb_predclsbal_pwr.py:36-47 on per-image probability sums (N,C) f64 in pool order:
    ``cumulated += mean(prob, dim=(0,2,3))`` per REFERENCE batch in fp32, ``/ len(loader)``, ``(coeff * p + 1) ** -2``."""
    prob = torch.as_tensor(prob_sum_all, dtype=torch.float64)
    cumulated = torch.zeros(prob.shape[1], dtype=torch.float32)
    n_batches = 0
    for b0 in range(0, prob.shape[0], ref_batch):
        part = prob[b0:b0 + ref_batch]
        cumulated += (part.sum(dim=0) / (part.shape[0] * pixels_per_image)).to(torch.float32)
        n_batches += 1
    return (float(coeff) * (cumulated / n_batches) + 1.0) ** (-2)


def tie_free(logits, temp, bump=0.01, dtype=None, drop_last_too=False):
    """Remove exact top-2 ties so that the oracle is well defined: the reference takes ``topk`` on the softmax
    PROBABILITIES (my_bvsb.py:20-21), whose order among equal values is arbitrary, while the kernels keep the first
    index.  Pixels whose two largest probabilities are equal floats get their arg-max logit bumped; with ``dtype`` the
    result stays exactly representable in it (bf16 inputs); ``drop_last_too``: also tie-free over the first C-1 planes
    (the ``preds[:, :-1]`` view plain my_bvsb reads on predignore nets)."""
    x = logits.float().clone()
    if dtype is not None:
        x = x.to(dtype).float()          # the ties that matter are those of the ROUNDED values
    for _ in range(8):
        dirty = False
        for view in ((x, x[:, :-1]) if drop_last_too else (x,)):
            prob = torch.softmax(view / temp, dim=1)
            top = prob.topk(2, dim=1)
            tie = top.values[:, 0] == top.values[:, 1]
            if bool(tie.any()):
                dirty = True
                first = view.argmax(dim=1, keepdim=True)
                view.scatter_add_(1, first, tie.unsqueeze(1).to(view.dtype) * bump)     # in place: views write through to x
        if dtype is not None:
            x = x.to(dtype).float()
        if not dirty:
            return x
    raise AssertionError("could not make the logits tie-free")
