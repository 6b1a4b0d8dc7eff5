"""Import shim that lets the UNMODIFIED reference (``/root/reference``) be imported
in a container without pyro-ppl / PyTables / anndata / matplotlib.

TEST INFRASTRUCTURE ONLY.  Used by ``tools/make_golden.py`` (in the build container,
where ``/root/reference`` exists) to generate the golden vectors committed under
``tests/golden/``.  Nothing in the product path (``cellbender_b200/``) imports it, and
nothing run on the GPU box reads ``/root/reference``.

What is stubbed: the ``pyro`` namespace is mapped onto
``torch.distributions`` + a dict-backed param store; ``tables``/``anndata``/``matplotlib``
are empty modules.  No reference source is copied; the reference files are imported
from where they lie.
"""
import os
import random
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("CELLBENDER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cellbender", "remove_background"))


_PARAM_STORE = {}


def _param(name, init=None, constraint=None, event_dim=None):
    if name not in _PARAM_STORE:
        if init is None:
            raise KeyError(f"pyro.param('{name}') requested before initialisation (shim store)")
        _PARAM_STORE[name] = init() if callable(init) else init
    return _PARAM_STORE[name]


class _Store(dict):
    def keys(self):
        return _PARAM_STORE.keys()

    def clear(self):
        _PARAM_STORE.clear()


def _set_rng_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def install():
    """Install the stub modules into ``sys.modules`` and put the reference on ``sys.path``."""
    if "pyro" in sys.modules and getattr(sys.modules["pyro"], "__cb_shim__", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    pyro = mod("pyro")
    pyro.__cb_shim__ = True
    pyro.param = _param
    pyro.get_param_store = lambda: _Store()
    pyro.clear_param_store = _PARAM_STORE.clear
    pyro.set_rng_seed = _set_rng_seed
    pyro.module = lambda name, module, update_module_params=False: module
    pyro.enable_validation = lambda *a, **k: None

    dist = mod("pyro.distributions")
    import torch.distributions as td

    for k in dir(td):
        if not k.startswith("_"):
            setattr(dist, k, getattr(td, k))
    dist.enable_validation = lambda *a, **k: None
    tdm = mod("pyro.distributions.torch_distribution")

    class TorchDistributionMixin:  # noqa: D401 - marker base class only
        pass

    class TorchDistribution(td.Distribution, TorchDistributionMixin):
        pass

    tdm.TorchDistributionMixin = TorchDistributionMixin
    tdm.TorchDistribution = TorchDistribution
    dist.torch_distribution = tdm
    pyro.distributions = dist

    pyro.poutine = mod("pyro.poutine")
    infer = mod("pyro.infer")

    def config_enumerate(default=None, **kw):
        def deco(f):
            return f

        return deco

    infer.config_enumerate = config_enumerate
    for cls in ("SVI", "Trace_ELBO", "TraceEnum_ELBO", "JitTrace_ELBO", "JitTraceEnum_ELBO"):
        setattr(infer, cls, type(cls, (), {}))
    pyro.infer = infer

    optim = mod("pyro.optim")
    optim.PyroOptim = type("PyroOptim", (), {})
    optim.ClippedAdam = type("ClippedAdam", (), {})
    lrs = mod("pyro.optim.lr_scheduler")
    lrs.PyroLRScheduler = type("PyroLRScheduler", (), {})
    optim.lr_scheduler = lrs
    ca = mod("pyro.optim.clipped_adam")
    ca.ClippedAdam = optim.ClippedAdam
    optim.clipped_adam = ca
    pyro.optim = optim

    util = mod("pyro.util")
    util.ignore_jit_warnings = lambda *a, **k: (lambda f: f)
    util.set_rng_seed = _set_rng_seed
    pyro.util = util

    for name in ("tables", "matplotlib", "matplotlib.pyplot", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "matplotlib.colors"):
        mod(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    ann = mod("anndata")
    ann.AnnData = object

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def set_param(name, value):
    """Seed the shim's param store (e.g. ``d_empty_loc`` before ``EncodeNonZLatents.forward``)."""
    _PARAM_STORE[name] = value


def clear_params():
    _PARAM_STORE.clear()
