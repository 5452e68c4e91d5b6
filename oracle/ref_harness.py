"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference for fixture generation.

Runs only in the build container, where /root/reference exists (it does not
exist on the GPU box).  Nothing under grassmanntn_b200/, bench.py or the `-m gpu`
tests imports this file; it is used by tests/golden/make_golden.py (committed
next to the fixtures it produced) and by optional `not gpu` cross-checks that
skip when /root/reference is absent.

Recipe:
  * the reference package imports itself by name, so a scratch dir with a
    symlink  grassmanntn -> /root/reference  is put on sys.path;
  * `sparse` and `opt_einsum` are not installed -> oracle/ref_shims stand-ins;
  * HEAD fix 1: gauge2d.py:204 calls arith.exp which does not exist
    (arith.py:830 defines exp_gnum)  -> alias;
  * HEAD fix 2: power_block (__init__.py:6071-6082) discards its np.where
    result so gtn.sqrt(block) is a no-op -> replaced by a version that stores it
    (the commented-out line :6078 shows the intent).
"""
import os
import sys
import tempfile

REFERENCE_DIR = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_cached = None


def available():
    return os.path.isdir(REFERENCE_DIR)


def load_reference():
    """Return the imported reference module (`grassmanntn`), patched as documented."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present (expected on the build container only)")
    scratch = tempfile.mkdtemp(prefix="gtn_ref_")
    os.symlink(REFERENCE_DIR, os.path.join(scratch, "grassmanntn"))
    sys.path.insert(0, scratch)
    sys.path.insert(0, os.path.join(_HERE, "ref_shims"))
    sys.dont_write_bytecode = True
    import numpy as np
    import grassmanntn as gtn  # noqa: E402  (the real reference)
    from grassmanntn import arith

    if not hasattr(arith, "exp"):
        arith.exp = arith.exp_gnum
    for mod in (gtn.gauge2d, gtn.gauge2d_block):
        if hasattr(mod, "arith") and not hasattr(mod.arith, "exp"):
            mod.arith.exp = arith.exp_gnum

    def power_block_fixed(T, p, rcond=1e-10):    # the reference's own default (:6071); gtn.power drops the caller's
        # restatement of the intent of __init__.py:6071-6082 (result stored, not dropped)
        this_format = T.format
        T = T.force_format("matrix")
        it = np.nditer(T.data, flags=["multi_index", "refs_ok"])
        for _ in it:
            blk = it.multi_index
            d = T.data[blk]
            with np.errstate(all="ignore"):
                T.data[blk] = np.where(np.abs(d) > rcond, np.power(d + 0.0 * (np.abs(d) <= rcond) + (np.abs(d) <= rcond), p), 0)
        return T.force_format(this_format)

    gtn._power_block_head = gtn.power_block
    gtn.power_block = power_block_fixed
    _cached = gtn
    return gtn
