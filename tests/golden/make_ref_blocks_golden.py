"""Generate tests/golden/ref_torch_blocks.npz and tests/golden/ref_sklearn_dbscan.npz.

* ref_torch_blocks.npz: the REAL ``PointTransformerLayer`` and ``TransitionDown`` of the reference
  (models/modules/cbl_point_transformer/blocks.py:14-79), imported from /root/reference and run on CPU (the CUDA extension
  replaced by the oracle's C restatement, which is pinned bitwise against the verbatim reference kernels), in train() and
  eval() BatchNorm mode, with the layer's full ``state_dict``.
* ref_sklearn_dbscan.npz: ``sklearn.cluster.DBSCAN(eps, min_samples).fit(points)`` -- the call of
  ops_utils.get_clustering_labels (ops_utils.py:98) -- on small clouds with noise, border points and duplicates.

    python tests/golden/make_ref_blocks_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_models  # noqa: E402
from toothgroupnetwork_b200 import clouds  # noqa: E402


def cpu_cuda_shims():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    mk = lambda dt: (lambda *a: torch.tensor(a[0], dtype=dt) if len(a) == 1 and isinstance(a[0], (list, tuple))
                     else torch.zeros(*[int(x) for x in a], dtype=dt))
    torch.cuda.IntTensor, torch.cuda.FloatTensor = mk(torch.int32), mk(torch.float32)


def randomize_bn(module, gen):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)


def blocks_fixture():
    warnings.filterwarnings("ignore")
    cpu_cuda_shims()
    w = ref_models.World("reference", cpu_dry_run=True)
    fix = {}
    gen = torch.Generator().manual_seed(5)
    p = torch.cat([clouds.dental_arch(420, 3)[0], clouds.dental_arch(280, 4)[0]]).contiguous()
    o = torch.tensor([420, 700], dtype=torch.int32)
    fix["p"], fix["o"] = p.numpy(), o.numpy()
    with w, torch.no_grad():
        B = w.mod("models.modules.cbl_point_transformer.blocks")
        for c, K in ((32, 16), (64, 8)):
            torch.manual_seed(c)
            layer = B.PointTransformerLayer(c, c, 8, K)
            randomize_bn(layer, gen)
            x = torch.randn(700, c, generator=gen)
            tag = f"ptl{c}"
            fix[tag + "_x"] = x.numpy()
            fix[tag + "_K"] = np.int64(K)
            for k, v in layer.state_dict().items():
                fix[f"{tag}_state_{k}"] = v.clone().numpy()
            state = {k: v.clone() for k, v in layer.state_dict().items()}
            fix[tag + "_out_train"] = layer.train()([p, x, o]).numpy()
            for k, v in layer.state_dict().items():
                if "running" in k:
                    fix[f"{tag}_after_{k}"] = v.clone().numpy()      # .numpy() alone would alias the live buffer
            layer.load_state_dict(state)
            fix[tag + "_out_eval"] = layer.eval()([p, x, o]).numpy()
        torch.manual_seed(7)
        td = B.TransitionDown(32, 64, 4, 16)
        randomize_bn(td, gen)
        x = torch.randn(700, 32, generator=gen)
        fix["td_x"] = x.numpy()
        for k, v in td.state_dict().items():
            fix[f"td_state_{k}"] = v.clone().numpy()
        state = {k: v.clone() for k, v in td.state_dict().items()}
        n_p, n_x, n_o = td.train()([p, x, o])
        fix["td_p"], fix["td_out_train"], fix["td_o"] = n_p.numpy(), n_x.numpy(), n_o.numpy()
        td.load_state_dict(state)
        fix["td_out_eval"] = td.eval()([p, x, o])[1].numpy()
    out = os.path.join(HERE, "ref_torch_blocks.npz")
    np.savez_compressed(out, **fix)
    print(out, os.path.getsize(out), "bytes")


def dbscan_fixture():
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(11)
    fix = {}
    blobs = [rng.normal(c, s, (m, 3)) for c, s, m in ((0.0, 0.02, 300), (0.25, 0.01, 150), (-0.3, 0.04, 400), (0.6, 0.004, 35))]
    pts = np.concatenate(blobs + [rng.uniform(-0.8, 0.8, (500, 3))]).astype(np.float32)
    pts = np.concatenate([pts, pts[:60]])
    pts = pts[rng.permutation(len(pts))]
    arch, _, label = clouds.dental_arch(3000, 9)
    arch = arch.numpy()[label.numpy() > 0]
    for name, x, eps, ms in (("blobs", pts, 0.03, 30), ("blobs_loose", pts, 0.05, 5), ("arch", arch, 0.06, 12)):
        c = DBSCAN(eps=eps, min_samples=ms).fit(x)
        fix[name + "_points"], fix[name + "_eps"], fix[name + "_min_samples"] = x, np.float64(eps), np.int64(ms)
        fix[name + "_labels"], fix[name + "_core"] = c.labels_, c.core_sample_indices_
        print(name, x.shape, "clusters", c.labels_.max() + 1, "noise", int((c.labels_ == -1).sum()), "core", len(c.core_sample_indices_),
              "border", int((c.labels_ >= 0).sum() - len(c.core_sample_indices_)))
    out = os.path.join(HERE, "ref_sklearn_dbscan.npz")
    np.savez_compressed(out, **fix)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    dbscan_fixture()
    blocks_fixture()
