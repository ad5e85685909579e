"""Import shims that let the UNMODIFIED reference (ypw0102/GDR, mounted read-only at
/root/reference) run in the dev container.  TEST INFRASTRUCTURE ONLY.

Nothing here is imported by the product package `gdr_b200`; it is used by
`oracle/make_golden.py` (which generates the committed fixtures under tests/golden/)
and by the `-m "not gpu"` tests that cross-check the restatement in `oracle/gdr_oracle.py`
against the live reference when /root/reference is mounted.  /root/reference does not
exist on the GPU box: there the same unmodified files are found under baseline/_ref/ (a git-ignored
copy made by `__graft_entry__.build()` in the dev container, shipped by gpurun like the built .so),
which is what bench.py's `--impl reference` / `cpu_baseline` legs and tests/test_gpu_integration.py use.

Two loaders, because the reference's two halves need different `transformers`:
  * load_ref_dense()        -> reference GDR_model/dense.py (+ encoder.py) against the
                               INSTALLED transformers (only AutoModel/PreTrainedModel names are used).
  * load_ref_main_models()  -> reference GDR_model/main_models.py against the VENDORED
                               transformers 3.4.0 (must run in a separate process from the former).
"""
import builtins
import importlib
import os
import sys
import types

def _find_ref_root() -> str:
    """The reference tree: $GDR_REFERENCE_ROOT, the read-only mount of the dev container, or the git-ignored copy that
    `__graft_entry__.build()` ships to the GPU box as baseline/_ref/ (unmodified files, never committed)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("GDR_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "GDR_model", "dense.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_ref_root()
REF_MODEL_DIR = os.path.join(REF_ROOT, "GDR_model")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_MODEL_DIR, "dense.py"))


def load_ref_dense():
    """Return the reference `dense` module (dense.py:1-71)."""
    # encoder.py:150-151 annotates with names whose import is commented out (encoder.py:13-14).
    builtins.ModelArguments = builtins.TrainingArguments = object
    sys.dont_write_bytecode = True
    if "gdr_ref" not in sys.modules:
        pkg = types.ModuleType("gdr_ref")
        pkg.__path__ = [REF_MODEL_DIR]
        sys.modules["gdr_ref"] = pkg
    return importlib.import_module("gdr_ref.dense")


def load_ref_main_models():
    """Return (main_models, generation_utils_previous) of the reference.

    Run in a process that has NOT imported the installed `transformers`.
    """
    import collections
    import collections.abc

    os.environ["PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION"] = "python"
    os.environ["WANDB_DISABLED"] = "true"
    sys.dont_write_bytecode = True
    for n in ("Sequence", "Mapping", "Iterable", "MutableMapping"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    sys.modules.setdefault("sacremoses", types.ModuleType("sacremoses"))
    import torch

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = torch.nn.Module
        pl.__version__ = "stub"
        sys.modules["pytorch_lightning"] = pl
    if REF_MODEL_DIR not in sys.path:
        sys.path.insert(0, REF_MODEL_DIR)  # vendored transformers 3.4.0 shadows the installed one
    mm = importlib.import_module("main_models")
    gp = importlib.import_module("transformers.generation_utils_previous")
    # CPU-only container: the fine stage calls .cuda() per document (main_models.py:1458-1462).
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    return mm, gp
