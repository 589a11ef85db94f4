"""Recipe for ``oracle/_ref``: the reference's OWN model modules, so that ``bench.py --impl reference``
and the ``cpu_baseline`` leg can time the unmodified reference forward on the GPU box's host cores
(``cpu_baseline.kind == "reference"``) instead of the oracle port.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python: "building" it means placing the two files its forward needs where they
can travel with the ``gpurun`` snapshot -- ``oracle/_ref/`` is git-ignored (never part of the history)
but not gpurun-ignored.  Nothing is copied into tracked paths.

  /root/reference/vad/models/self_attention.py   -> oracle/_ref/vad/models/self_attention.py
  /root/reference/vad/modeling/transformer.py    -> oracle/_ref/vad/modeling/transformer.py

(they import only torch / numpy / math / typing; the reference ships no ``__init__.py`` files there, the
recipe writes empty ones).  Run by ``__graft_entry__.build()`` when ``/root/reference`` exists; on the
GPU box the pre-placed files are used, and when they are absent the bench falls back to the oracle port
and says ``kind: "port"``.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
FILES = ("vad/models/self_attention.py", "vad/modeling/transformer.py")


def build_ref(reference_root: str = os.environ.get("VAD_REFERENCE", "/root/reference")) -> bool:
    """Returns True when oracle/_ref holds the reference modules afterwards."""
    if os.path.isdir(reference_root):
        for rel in FILES:
            src, dst = os.path.join(reference_root, rel), os.path.join(REF_DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
        # regular packages (empty __init__.py written here, not copied): a namespace portion would lose
        # against the repository's own `vad` compatibility package further down sys.path
        for pkg in ("vad", "vad/models", "vad/modeling"):
            open(os.path.join(REF_DST, pkg, "__init__.py"), "a").close()
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


def import_reference_model():
    """-> the reference's ``SelfAttentiveVAD`` class from oracle/_ref, or None.  Must run in a process
    that has not imported the repository's own ``vad`` compatibility package."""
    if not all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES):
        return None
    for name in [m for m in sys.modules if m == "vad" or m.startswith("vad.")]:
        del sys.modules[name]
    sys.path.insert(0, REF_DST)
    try:
        from vad.models.self_attention import SelfAttentiveVAD   # the reference's module
        if os.path.realpath(sys.modules["vad.models.self_attention"].__file__).startswith(os.path.realpath(REF_DST)):
            return SelfAttentiveVAD
        return None
    except Exception:
        return None
    finally:
        sys.path.remove(REF_DST)


if __name__ == "__main__":
    print("oracle/_ref ready:", build_ref())
