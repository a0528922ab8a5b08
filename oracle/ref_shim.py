"""Import shim for the UNMODIFIED upstream reference (test infrastructure only).

The reference (``/root/reference``, read-only) is pure Python + numba.  Two of its imports are
absent from this image and are not on the hot path: ``gym.spaces`` (only describes spaces,
stratego_multiagent_env.py:11, 362, 398-427) and ``h5py`` (curriculum inits only, util.py:3).
``import_reference()`` registers trivial stand-ins for both and returns the imported package.

Only ``oracle/gen_golden.py``, ``tools/gen_setup_tables.py`` and the reference-pinning tests use
this module, and only in the build container: ``/root/reference`` does not exist on the GPU box.
Nothing under ``stratego_env_b200/`` may import it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STRATEGO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stratego_env"))


def import_reference():
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_stratego_ref")
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low=None, high=None, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        class Dict:
            def __init__(self, spaces=None):
                self.spaces = spaces

        spaces.Discrete, spaces.Box, spaces.Dict = Discrete, Box, Dict
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import stratego_env  # noqa: F401
    return stratego_env
