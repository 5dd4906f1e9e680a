"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference implementation from /root/reference with stub
modules for the four third-party packages that are not installed in this image
(SURVEY.md Appendix C).  /root/reference exists only in the build container, so this
module is used solely by tools/make_golden.py (fixture generation) and by the CPU
tests that re-validate the oracle restatement when the reference tree is present.
It cannot travel to the GPU box; everything that runs there uses oracle/star_oracle.py
plus the committed fixtures under tests/golden/.

Stubs:
  torchtyping  TensorType[...] -> torch.Tensor, patch_typeguard -> no-op
  typeguard    typechecked -> identity
  lietorch     SO3 = SE3 = object               (imported at star__.py:6, unused)
  pypose       SE3(x).Act(p), SO3(q).Act(p), identity_SE3(n)   (star__.py:192,196)

pypose is not vendored and not version-pinned by the reference (environment.yaml has
no entry).  The stub restates the published pypose semantics:
  layout  [tx,ty,tz,qx,qy,qz,qw];   SO3 Act:  uv = 2 q_v x p ; out = p + q_w uv + q_v x uv
  SE3 Act: t + SO3Act(q, p);  quaternion NOT normalised by Act.
  backward (pypose custom Function): gradient w.r.t. a LEFT tangent perturbation,
  padded with one zero:  SE3 -> [sum g, sum out x g, 0],  SO3 -> [sum out x g, 0].
Parity for the quaternion-pose branch is therefore "unpinned" (no reference test or
golden vector exists); it is cross-checked against the 4x4-matrix branch, which is
reference code.
"""
import os
import sys
import types
import argparse

import torch

def _find_reference_root():
    """The mounted reference tree (build container), else oracle/_ref/ -- the reference's own hot-path files placed there
    unmodified by oracle/make_ref.py (git-ignored; travels to the GPU box with the snapshot)."""
    cands = [os.environ.get("STAR_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "models", "rendering__.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "rendering__.py"))


# --------------------------------------------------------------------------- pypose stub
def _quat_rotate(q, p):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.cross(qv.expand_as(p), p, dim=-1)
    return p + qw * uv + torch.cross(qv.expand_as(p), uv, dim=-1)


def _quat_to_matrix(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    # Matrix of the (un-normalised) map p -> p + w*(2 v x p) + v x (2 v x p)
    e = torch.eye(3, dtype=q.dtype, device=q.device)
    K = torch.zeros(q.shape[:-1] + (3, 3), dtype=q.dtype, device=q.device)
    K[..., 0, 1], K[..., 0, 2] = -z, y
    K[..., 1, 0], K[..., 1, 2] = z, -x
    K[..., 2, 0], K[..., 2, 1] = -y, x
    return e + 2.0 * w[..., None, None] * K + 2.0 * (K @ K)


class _SE3Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, p):
        out = X[..., :3] + _quat_rotate(X[..., 3:7], p)
        ctx.save_for_backward(X, out)
        return out

    @staticmethod
    def backward(ctx, g):
        X, out = ctx.saved_tensors
        rot = torch.cross(out, g, dim=-1)
        gx = torch.cat([g, rot, torch.zeros_like(g[..., :1])], -1)
        while gx.dim() > X.dim():
            gx = gx.sum(0)
        R = _quat_to_matrix(X[..., 3:7])
        gp = g @ R  # R^T g
        return gx, gp


class _SO3Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p):
        out = _quat_rotate(q, p)
        ctx.save_for_backward(q, out)
        return out

    @staticmethod
    def backward(ctx, g):
        q, out = ctx.saved_tensors
        rot = torch.cross(out, g, dim=-1)
        gq = torch.cat([rot, torch.zeros_like(g[..., :1])], -1)
        while gq.dim() > q.dim():
            gq = gq.sum(0)
        R = _quat_to_matrix(q)
        return gq, g @ R


class _LieSE3:
    def __init__(self, x):
        self.x = x

    def Act(self, p):
        return _SE3Act.apply(self.x, p)

    def tensor(self):
        return self.x


class _LieSO3:
    def __init__(self, x):
        self.x = x

    def Act(self, p):
        return _SO3Act.apply(self.x, p)

    def tensor(self):
        return self.x


def _identity_SE3(n):
    x = torch.zeros(n, 7)
    x[:, 6] = 1.0
    return _LieSE3(x)


def install_stubs():
    if "torchtyping" not in sys.modules:
        m = types.ModuleType("torchtyping")

        class TensorType:
            def __class_getitem__(cls, item):
                return torch.Tensor

        m.TensorType = TensorType
        m.patch_typeguard = lambda: None
        sys.modules["torchtyping"] = m
    m = types.ModuleType("typeguard")
    m.typechecked = lambda f=None, **k: f if f is not None else (lambda g: g)
    sys.modules["typeguard"] = m
    if "lietorch" not in sys.modules:
        m = types.ModuleType("lietorch")
        m.SO3 = object
        m.SE3 = object
        sys.modules["lietorch"] = m
    if "pypose" not in sys.modules:
        m = types.ModuleType("pypose")
        m.SE3 = _LieSE3
        m.SO3 = _LieSO3
        m.identity_SE3 = _identity_SE3
        sys.modules["pypose"] = m


_REF = None


def load_reference():
    """Returns a namespace with the reference's hot-path modules (rendering__, star__, nerf, embedder)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")
             or k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        rendering = importlib.import_module("models.rendering__")
        star = importlib.import_module("models.star__")
        nerf = importlib.import_module("models.nerf")
        embedder = importlib.import_module("models.embedder")
        ns = types.SimpleNamespace(rendering=rendering, star=star, nerf=nerf, embedder=embedder)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # keep the reference modules private to `ns`; restore whatever was there before
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")
                  or k == "utils" or k.startswith("utils.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    _REF = ns
    return ns


def make_args(num_vehicles=0, N_importance=128, chunk=8192, netchunk=16384, white_bkgd=False,
              end_barf=-1, far_dist=1e10, raw_noise_std=0.0, netdepth=8, netwidth=256,
              multires=10, multires_views=4):
    """The 16 fields read by STaR.__init__/NeRF.__init__ (star__.py:28-53, nerf.py:41-101)."""
    return argparse.Namespace(
        num_vehicles=num_vehicles, chunk=chunk, far_dist=far_dist, N_importance=N_importance,
        netdepth=netdepth, netwidth=netwidth, netdepth_fine=netdepth, netwidth_fine=netwidth,
        multires=multires, multires_views=multires_views, end_barf=end_barf, i_embed=0,
        use_viewdirs=True, netchunk=netchunk, raw_noise_std=raw_noise_std, white_bkgd=white_bkgd)
