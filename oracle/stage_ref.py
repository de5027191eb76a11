"""Recipe that stages the UNMODIFIED reference for the GPU box (test / baseline infrastructure only).

The reference (tqch/v-diffusion-torch) is a pure-Python package: there is nothing to compile, and
``/root/reference`` does not exist on the GPU box.  ``stage()`` copies its importable package and its JSON
configs, byte for byte, from ``/root/reference`` into the git-ignored directory ``oracle/_ref/`` so that they
travel with the repo snapshot (``oracle/_ref/`` is listed in ``.gitignore`` and NOT in ``.gpurunignore``): the
sources never enter the history, only this recipe does.  ``__graft_entry__.build()`` calls it in the build
container; on the GPU box (no ``/root/reference``) it is a no-op and the previously staged copy is used.

``load()`` imports the staged copy (matplotlib stubbed: ``v_diffusion/utils.py`` imports it at module top only for
plotting helpers that this path never calls) and returns the module.  Consumers: ``bench.py`` --
``--impl reference``, ``cpu_baseline`` and the ``reference_gpu_eager`` block -- and ``tests/`` (reference-on-GPU
parity at sizes the CPU cannot reach).  Nothing under ``v-diffusion-torch_b200/`` may import this.
"""
import filecmp
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
PARTS = ("v_diffusion", "configs")


def staged():
    return os.path.isfile(os.path.join(REF_DST, "v_diffusion", "diffusion.py"))


def _same(a, b):
    c = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if c.left_only or c.right_only or c.diff_files or c.funny_files:
        return False
    return all(_same(os.path.join(a, d), os.path.join(b, d)) for d in c.common_dirs)


def stage(verbose=False):
    """Copy /root/reference/{v_diffusion,configs} -> oracle/_ref/ (only where /root/reference exists)."""
    if not os.path.isdir(REF_SRC):
        return staged()
    for part in PARTS:
        src, dst = os.path.join(REF_SRC, part), os.path.join(REF_DST, part)
        if os.path.isdir(dst) and _same(src, dst):
            continue
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if verbose:
            print(f"staged {src} -> {dst}")
    return staged()


_mod = None


def load():
    """Import the staged reference package; raises if it has not been staged."""
    global _mod
    if _mod is not None:
        return _mod
    if not staged():
        raise RuntimeError("oracle/_ref is not staged: run `python -c 'import __graft_entry__ as g; g.build()'` in the "
                           "build container (needs /root/reference)")
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        m.rcParams = {}
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import v_diffusion
    if not os.path.abspath(v_diffusion.__file__).startswith(REF_DST):
        raise RuntimeError(f"imported v_diffusion from {v_diffusion.__file__}, not from the staged copy")
    _mod = v_diffusion
    return v_diffusion


def config(name):
    """configs/<name>.json merged with configs/defaults.json by the reference's own fill_with_defaults."""
    import json
    ref = load()
    with open(os.path.join(REF_DST, "configs", f"{name}.json")) as f:
        cfg = json.load(f)
    with open(os.path.join(REF_DST, "configs", "defaults.json")) as f:
        defaults = json.load(f)
    ref.fill_with_defaults(cfg, defaults)
    return cfg


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else "not staged (no /root/reference here)")
