"""TEST INFRASTRUCTURE ONLY.  Import the UNMODIFIED reference from /root/reference (build container
only -- the path does not exist on the GPU box) with the stand-ins for its three absent dependencies."""
import os
import sys
import types

REF_ROOT = "/root/reference"
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "nuwa_pytorch"))


def import_reference():
    """Returns the modules (nuwa_pytorch.nuwa_pytorch, nuwa_pytorch.vqgan_vae) of the reference."""
    if not reference_available():
        raise RuntimeError("reference not present (expected only in the build container)")
    if _STANDINS not in sys.path:
        sys.path.insert(0, _STANDINS)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "ftfy" not in sys.modules:  # tokenizer-only dependency (tokenizer.py:10), irrelevant to the hot path
        m = types.ModuleType("ftfy")
        m.fix_text = lambda s: s
        sys.modules["ftfy"] = m
    import nuwa_pytorch.nuwa_pytorch as NP
    import nuwa_pytorch.vqgan_vae as VQ
    return NP, VQ
