"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import shim that lets the *unmodified* reference model code under
``/root/reference/model`` run in this container.

The reference imports a handful of third-party packages at module level that
are not installed here (yacs, dgl, rdkit, torch_geometric, lightning_utilities).
None of them does arithmetic on the hot path except DGL's
``update_all(copy_u, sum)`` (reference ``model/basic_model.py:591,612,617``) whose
published semantics -- an unweighted sum over in-edges, duplicate edges counted
once per duplicate -- is restated by :class:`FakeGraph` with ``index_add_``.

Only ``tests/`` (fixture generation, the live oracle checks and the drop-in test) and the
reference arm of ``bench.py`` use this file.  ``/root/reference`` does not exist on the GPU box:
there the shim falls back to ``oracle/_ref`` -- the same modules byte-compiled by
``oracle/build_ref.py`` in the build container (sourceless code objects, git-ignored, shipped by
gpurun) and imported through ``_RefbinFinder`` below.
``available()`` says whether either is present.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_SRC_ROOT = os.environ.get("DRUGLAMP_REFERENCE_ROOT", "/root/reference")
_PYC_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    if os.path.isdir(os.path.join(_SRC_ROOT, "model")):
        return _SRC_ROOT
    ver = os.path.join(_PYC_ROOT, "PYTHON_VERSION")
    if os.path.isdir(os.path.join(_PYC_ROOT, "model")) and os.path.exists(ver) and \
            open(ver).read().strip() == "%d.%d" % sys.version_info[:2]:
        return _PYC_ROOT
    return _SRC_ROOT


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def kind() -> str:
    """"source" (the tree under /root/reference) or "pyc" (oracle/_ref, byte-compiled from it)."""
    return "pyc" if REFERENCE_ROOT == _PYC_ROOT else "source"


class _AttrDict(dict):
    """Minimal stand-in for yacs.config.CfgNode (attribute access + clone)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        out = _AttrDict()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, _AttrDict) else v
        return out


class FakeGraph:
    """Duck-typed stand-in for a batched DGLGraph (only what GraphConv touches).

    Reference call sites: ``basic_model.py:148`` (ndata.pop, batch_size),
    ``:579-581`` (local_scope, in_degrees), ``:596`` (out_degrees),
    ``:611-618`` (srcdata/update_all/dstdata).
    """

    is_block = False

    def __init__(self, src, dst, num_nodes, batch_size, h=None):
        self.src = src.long()
        self.dst = dst.long()
        self.n = int(num_nodes)
        self.batch_size = int(batch_size)
        self.ndata = {} if h is None else {"h": h}
        self.srcdata = {}
        self.dstdata = {}

    @contextlib.contextmanager
    def local_scope(self):
        yield

    def num_nodes(self):
        return self.n

    def edges(self):
        return self.src, self.dst

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.n)

    def update_all(self, message_fn, reduce_fn):
        h = self.srcdata["h"]
        out = torch.zeros(self.n, h.shape[1], dtype=h.dtype, device=h.device)
        out.index_add_(0, self.dst, h[self.src])
        self.dstdata["h"] = out


class _RefbinFinder:
    """Meta-path finder / loader for the ``*.refbin`` tree of oracle/build_ref.py (py_compile output:
    16-byte header + marshalled code object)."""

    def __init__(self, root):
        self.root = root

    def _path(self, fullname):
        rel = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(os.path.join(rel, "__init__.refbin")):
            return os.path.join(rel, "__init__.refbin"), True
        if os.path.isfile(rel + ".refbin"):
            return rel + ".refbin", False
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        import importlib.util
        f, is_pkg = self._path(fullname)
        if f is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, self, origin=f, is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.dirname(f)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        with open(module.__spec__.origin, "rb") as f:
            code = marshal.loads(f.read()[16:])
        module.__file__ = module.__spec__.origin
        exec(code, module.__dict__)


_INSTALLED = False


def install() -> None:
    """Register stub modules and put the reference on sys.path (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    if "yacs" not in sys.modules:
        mod("yacs")
        mod("yacs.config", CfgNode=_AttrDict)
    if "dgl" not in sys.modules:
        fn = mod("dgl.function", copy_u=lambda *a, **k: ("copy_u", a, k),
                 sum=lambda *a, **k: ("sum", a, k))
        d = mod("dgl", function=fn, batch=None)
        d.function = fn
    if "rdkit" not in sys.modules:
        chem = mod("rdkit.Chem")
        r = mod("rdkit", Chem=chem)
        r.Chem = chem
    if "torch_geometric" not in sys.modules:
        u = mod("torch_geometric.utils", from_smiles=None)
        t = mod("torch_geometric", utils=u)
        t.utils = u
    if "lightning_utilities" not in sys.modules:
        rz = mod("lightning_utilities.core.rank_zero", rank_zero_only=lambda f: f)
        core = mod("lightning_utilities.core", rank_zero=rz)
        lu = mod("lightning_utilities", core=core)
        core.rank_zero = rz
        lu.core = core
    if REFERENCE_ROOT == _PYC_ROOT:
        sys.meta_path.insert(0, _RefbinFinder(_PYC_ROOT))
    elif REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


def reference_cfg():
    """The config keys the hot path reads (configs/default_config.py:4-61 + yaml)."""
    c = _AttrDict()
    c.DRUG = _AttrDict(NODE_IN_FEATS=75, MAX_NODES=512, PADDING=True)
    c.PROTEIN = _AttrDict(KERNEL_SIZE=[3, 6, 9], PADDING=True, SEQ_LEN=9 * 256, SITE_LEN=9)
    c.DECODER = _AttrDict(NAME="MLP", IN_DIM=256, HIDDEN_DIM=512, OUT_DIM=128, BINARY=1)
    c.RS = _AttrDict(MAX_MARGIN=0.5, RESET_EPOCH=100)
    return c


def build_reference_model(kind: str = "DrugLAMP", n_drug_feature: int = 384,
                          n_prot_feature: int = 640, n_hidden: int = 128):
    """Instantiate an unmodified reference model class (MInterface is unusable on py3.12)."""
    install()
    import importlib

    m = importlib.import_module(f"model.{kind}")
    cls = getattr(m, kind)
    return cls(n_drug_feature, n_prot_feature, n_hidden, **reference_cfg())
