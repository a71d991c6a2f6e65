"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stage the reference for the GPU box.

/root/reference does not exist on the GPU box; `gpurun` ships only /root/repo.  This recipe copies what an
UNMODIFIED `import nets` of the reference needs -- the Python files of its `nets/` package, nothing else --
into the git-ignored baseline/_ref/aanet/ (never into the tracked tree), and makes sure the reference's own
CUDA op is built into oracle/_ref/ (oracle/build_ref.py).  Run in the build container:

    python oracle/stage_ref.py            (also called by __graft_entry__.build())

Consumers: tests/test_gpu_full_model.py and profiles/full_model.py, which run the reference's
nets.AANet.forward (nets/aanet.py:212-229) twice -- stock (reference modules + reference CUDA op) and with
aanet_b200.dropin -- and compare final disparities / full-model pairs/s.
"""
import os
import shutil
import sys

REF = os.environ.get("AANET_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref", "aanet")


def staged_path():
    """Directory to put on sys.path so that `import nets` finds the staged reference, or None."""
    return DST if os.path.exists(os.path.join(DST, "nets", "aanet.py")) else None


def stage(verbose=False):
    src = os.path.join(REF, "nets")
    if not os.path.isdir(src):
        return staged_path()
    n = 0
    for d, _, files in os.walk(src):
        rel = os.path.relpath(d, REF)
        if os.sep + "src" in os.sep + rel or "__pycache__" in rel:      # the op's C++/CUDA sources are compiled in place
            continue
        for f in files:
            if not f.endswith(".py") or f == "setup.py":
                continue
            out = os.path.join(DST, rel, f)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(d, f)):
                shutil.copyfile(os.path.join(d, f), out)
                n += 1
    if verbose:
        print("staged %d file(s) under %s" % (n, DST))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import build_ref
    build_ref.build_if_possible()
    return staged_path()


if __name__ == "__main__":
    print(stage(verbose=True))
