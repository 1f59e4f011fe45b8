"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference IQL modules.

Imports ``/root/reference/algorithms/{finetune,offline}/iql.py`` with empty
``sys.modules`` stubs for the packages that are absent from this image
(gym, gymnasium, d4rl, pyrallis, wandb is present).  The hot path only needs
torch + numpy.  ``/root/reference`` exists only in the
build container, never on the GPU box: nothing under ``-m gpu`` tests,
``smoke()`` or ``bench.py`` may call this module.  It is used by
``oracle/gen_golden.py`` to produce the fixtures under ``tests/golden/`` and by
CPU-only tests (skipped when the reference tree is missing).
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JSRL_REFERENCE_ROOT", "/root/reference")
# build-time copy of the two reference files (oracle/make_ref.py, called by __graft_entry__.build()): git-ignored,
# but it travels to the GPU box with the repo snapshot so that bench.py's reference arm can time the UNMODIFIED
# reference classes there.  Only bench.py's `--impl reference` / `cpu_baseline` legs and CPU tests use it.
LOCAL_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _ref_path(variant: str) -> str:
    live = os.path.join(REFERENCE_ROOT, "algorithms", variant, "iql.py")
    if os.path.isfile(live):
        return live
    return os.path.join(LOCAL_REF, f"{variant}_iql.py")


def reference_available() -> bool:
    return os.path.isfile(_ref_path("finetune")) and os.path.isfile(_ref_path("offline"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _install_stubs():
    gym = _stub("gym", Env=object)
    if not hasattr(gym, "Env"):
        gym.Env = object
    _stub("gymnasium", Env=object)
    _stub("d4rl")

    def _wrap(*a, **k):
        def deco(fn):
            return fn
        return deco

    _stub("pyrallis", wrap=_wrap, parse=lambda *a, **k: None, dump=lambda *a, **k: None)
    # wandb is importable in this image but slow / chatty; a stub is enough.
    if "wandb" not in sys.modules:
        sys.modules["wandb"] = types.ModuleType("wandb")


_CACHE = {}


def load_reference_iql(variant: str = "finetune"):
    """Return the reference module ``algorithms/<variant>/iql.py`` (variant in
    {"finetune", "offline"})."""
    if variant in _CACHE:
        return _CACHE[variant]
    if not reference_available():
        raise FileNotFoundError(f"reference iql.py not found under {REFERENCE_ROOT} or {LOCAL_REF}")
    _install_stubs()
    path = _ref_path(variant)
    spec = importlib.util.spec_from_file_location(f"_ref_{variant}_iql", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE[variant] = mod
    return mod
