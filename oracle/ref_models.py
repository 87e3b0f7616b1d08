"""oracle/ref_models.py -- the reference's OWN model code as a checker (TEST INFRASTRUCTURE ONLY;
never imported by toothgroupnetwork_b200/).

The model-level parity runs import the
reference's unmodified ``models/modules/*.py`` twice in one process:

* ``World("reference")`` -- served by the reference's own operators:
  ``external_libs/pointops/functions/pointops.py`` + ``external_libs/pointnet2_utils/pointnet2_utils.py``
  on top of a ``pointops_cuda`` module that forwards to ``oracle/_ref/libpointops_ref.so`` (the
  reference's six .cu files compiled verbatim, ``oracle/Makefile``) -- the same ten functions
  ``src/pointops_api.cpp:12-23`` binds;
* ``World("b200")`` -- the same model files after ``toothgroupnetwork_b200.dropin.install()``.

The reference checkout does not exist on the GPU box, so ``stage_archive()`` (run by
``__graft_entry__.build()`` in the container that has /root/reference) packs its Python/YAML files
into ``oracle/_ref/reference_py.tgz``: git-ignored like the rest of ``oracle/_ref`` (nothing enters
history), but it travels with the gpurun snapshot and is unpacked to a temp directory on the box.
"""
from __future__ import annotations

import importlib
import os
import sys
import tarfile
import tempfile
import types
from typing import Dict, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DEFAULT = "/root/reference"
ARCHIVE = os.path.join(_HERE, "_ref", "reference_py.tgz")
_unpacked: Optional[str] = None

# top-level module names the reference's model code defines / shadows; purged between worlds
_REF_TOP = ("models", "external_libs", "ops_utils", "gen_utils", "pointops_cuda", "loss_meter", "augmentator")
_MODEL_MODULES = (
    "external_libs.pointops.functions.pointops",
    "external_libs.pointnet2_utils.pointnet2_utils",
    "models.modules.pointnet_pp",
    "models.modules.tsg_centroid_module",
    "models.modules.tsg_seg_module",
    "models.modules.tsegnet",
    "models.modules.grouping_network_module",
    "models.modules.cbl_point_transformer.blocks",
    "models.modules.cbl_point_transformer.heads",
    "ops_utils",
    "gen_utils",
)


def stage_archive(ref: str = REF_DEFAULT) -> Optional[str]:
    """Pack the reference's .py/.yaml files (no weights, no .git) for the trip to the GPU box."""
    if not os.path.isdir(os.path.join(ref, "models")):
        return None
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    newest = 0.0
    members = []
    for base, dirs, files in os.walk(ref):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__")]
        for f in files:
            if f.endswith((".py", ".yaml")):
                p = os.path.join(base, f)
                members.append(p)
                newest = max(newest, os.path.getmtime(p))
    if os.path.exists(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        return ARCHIVE
    with tarfile.open(ARCHIVE + ".tmp", "w:gz") as tar:
        for p in sorted(members):
            tar.add(p, arcname=os.path.relpath(p, ref))
    os.replace(ARCHIVE + ".tmp", ARCHIVE)
    return ARCHIVE


def reference_root() -> Optional[str]:
    """Directory holding the reference checkout: /root/reference here, the unpacked archive on the box."""
    global _unpacked
    if os.path.isdir(os.path.join(REF_DEFAULT, "models")):
        return REF_DEFAULT
    if _unpacked is not None:
        return _unpacked
    if not os.path.exists(ARCHIVE):
        return None
    d = tempfile.mkdtemp(prefix="tgn_reference_")
    with tarfile.open(ARCHIVE, "r:gz") as tar:
        tar.extractall(d, filter="data")
    _unpacked = d
    return d


def available() -> bool:
    return reference_root() is not None


# ------------------------------------------------------------------------------------------ stubs
def _stub_missing_io_packages() -> None:
    """open3d / trimesh / matplotlib are imported at the top of gen_utils.py:1,5,9 and ops_utils.py:7
    but used only for mesh I/O and plots, which no parity run touches."""
    for name in ("open3d", "trimesh", "matplotlib", "matplotlib.pyplot", "wandb"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
            if "." in name:
                parent, _, child = name.rpartition(".")
                setattr(sys.modules[parent], child, mod)


def make_reference_pointops_cuda(cpu_dry_run: bool = False) -> types.ModuleType:
    """``pointops_cuda`` (src/pointops_api.cpp:12-23) on the verbatim reference kernels.
    The launchers run on the legacy default stream, which is torch's default current stream."""
    import ctypes
    import torch
    mod = types.ModuleType("pointops_cuda")
    if cpu_dry_run:                      # harness dry run in the GPU-less container: oracle C restatement
        import numpy as np
        from oracle import oracle as O

        def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
            idx.copy_(torch.from_numpy(O.furthestsampling(xyz.numpy(), offset.numpy(), new_offset.numpy())))

        def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
            i, _, d2 = O.knnquery(int(nsample), xyz.numpy(), new_xyz.numpy(), offset.numpy(), new_offset.numpy())
            idx.copy_(torch.from_numpy(np.ascontiguousarray(i)))
            dist2.copy_(torch.from_numpy(np.ascontiguousarray(d2)))

        mod.furthestsampling_cuda, mod.knnquery_cuda = furthestsampling_cuda, knnquery_cuda
        return mod
    from oracle import ref_cuda
    lib = ref_cuda.lib()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    I = lambda v: int(v)

    def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        lib.furthestsampling_cuda_launcher(I(b), I(n_max), p(xyz), p(offset), p(new_offset), p(tmp), p(idx))

    def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        lib.knnquery_cuda_launcher(I(m), I(nsample), p(xyz), p(new_xyz), p(offset), p(new_offset), p(idx), p(dist2))

    def grouping_forward_cuda(m, nsample, c, inp, idx, out):
        lib.grouping_forward_cuda_launcher(I(m), I(nsample), I(c), p(inp), p(idx), p(out))

    def grouping_backward_cuda(m, nsample, c, grad_out, idx, grad_in):
        lib.grouping_backward_cuda_launcher(I(m), I(nsample), I(c), p(grad_out), p(idx), p(grad_in))

    def interpolation_forward_cuda(n, c, k, inp, idx, weight, out):
        lib.interpolation_forward_cuda_launcher(I(n), I(c), I(k), p(inp), p(idx), p(weight), p(out))

    def interpolation_backward_cuda(n, c, k, grad_out, idx, weight, grad_in):
        lib.interpolation_backward_cuda_launcher(I(n), I(c), I(k), p(grad_out), p(idx), p(weight), p(grad_in))

    def subtraction_forward_cuda(n, nsample, c, in1, in2, idx, out):
        lib.subtraction_forward_cuda_launcher(I(n), I(nsample), I(c), p(in1), p(in2), p(idx), p(out))

    def subtraction_backward_cuda(n, nsample, c, idx, grad_out, g1, g2):
        lib.subtraction_backward_cuda_launcher(I(n), I(nsample), I(c), p(idx), p(grad_out), p(g1), p(g2))

    def aggregation_forward_cuda(n, nsample, c, w_c, inp, pos, weight, idx, out):
        lib.aggregation_forward_cuda_launcher(I(n), I(nsample), I(c), I(w_c), p(inp), p(pos), p(weight), p(idx), p(out))

    def aggregation_backward_cuda(n, nsample, c, w_c, inp, pos, weight, idx, grad_out, gi, gp, gw):
        lib.aggregation_backward_cuda_launcher(I(n), I(nsample), I(c), I(w_c), p(inp), p(pos), p(weight), p(idx), p(grad_out),
                                               p(gi), p(gp), p(gw))

    for f in (furthestsampling_cuda, knnquery_cuda, grouping_forward_cuda, grouping_backward_cuda, interpolation_forward_cuda,
              interpolation_backward_cuda, subtraction_forward_cuda, subtraction_backward_cuda, aggregation_forward_cuda,
              aggregation_backward_cuda):
        setattr(mod, f.__name__, f)
    return mod


def _purge() -> Dict[str, types.ModuleType]:
    gone = {}
    for name in list(sys.modules):
        if name.split(".")[0] in _REF_TOP:
            gone[name] = sys.modules.pop(name)
    return gone


class World:
    """One import of the reference's model code bound to one operator set.

        ref = World("reference"); new = World("b200")
        with ref:  m_ref = ref.mod("models.modules.pointnet_pp").get_model().cuda()
        with new:  m_new = new.mod("models.modules.pointnet_pp").get_model().cuda()

    Construction and forward calls happen inside ``with world:`` so that the reference's lazy
    imports (``cbl_point_transformer_module.get_model`` imports ``.util.config`` when called) resolve
    to that world's modules.
    """

    def __init__(self, ops: str, cpu_dry_run: bool = False, accelerate_crops: bool = True, fused_blocks: bool = True):
        assert ops in ("reference", "b200")
        self.ops = ops
        root = reference_root()
        if root is None:
            raise RuntimeError("reference checkout not staged: neither /root/reference nor oracle/_ref/reference_py.tgz exists")
        self.root = root
        saved = _purge()
        _stub_missing_io_packages()
        sys.path.insert(0, root)
        try:
            if ops == "b200":
                from toothgroupnetwork_b200 import dropin
                dropin.install()
            else:
                sys.modules["pointops_cuda"] = make_reference_pointops_cuda(cpu_dry_run)
            for name in _MODEL_MODULES:
                importlib.import_module(name)
            if ops == "b200" and accelerate_crops:
                # the crop search between the two stages (sklearn KDTree on the host in the reference) on the GPU too
                from toothgroupnetwork_b200 import clustering, crops
                crops.accelerate(sys.modules["ops_utils"])
                clustering.accelerate(sys.modules["ops_utils"])
                if "models.modules.tsegnet" in sys.modules:
                    clustering.accelerate_dbscan(sys.modules["models.modules.tsegnet"])
            if ops == "b200" and fused_blocks:
                # no-grad forwards of PointTransformerLayer / TransitionDown on the fused kernels (autograd keeps the reference's code)
                from toothgroupnetwork_b200 import blocks_fused
                blocks_fused.accelerate(sys.modules["models.modules.cbl_point_transformer.blocks"])
        finally:
            self.modules = _purge()
            sys.path.remove(root)
            sys.modules.update(saved)
        self._saved: Optional[Dict[str, types.ModuleType]] = None

    def __enter__(self) -> "World":
        self._saved = _purge()
        sys.modules.update(self.modules)
        sys.path.insert(0, self.root)
        return self

    def __exit__(self, *exc) -> None:
        self.modules.update(_purge())
        sys.path.remove(self.root)
        sys.modules.update(self._saved or {})
        self._saved = None

    def mod(self, name: str) -> types.ModuleType:
        return self.modules[name]
