"""Minimal space descriptors with the attribute names of ``gym.spaces`` (the reference only uses them to
describe shapes, maenv:362, 398-427).  If gym / gymnasium is installed its classes are used instead."""
try:  # pragma: no cover - neither is present in the build image
    from gym.spaces import Box, Dict, Discrete  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from gymnasium.spaces import Box, Dict, Discrete  # type: ignore
    except Exception:  # noqa: BLE001
        import numpy as np

        class Discrete:
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.int64

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high = low, high
                self.shape = tuple(int(s) for s in shape) if shape is not None else np.shape(low)
                self.dtype = dtype

            def __repr__(self):
                return "Box(%s, %s, %s)" % (self.low, self.high, self.shape)

        class Dict:
            def __init__(self, spaces=None):
                self.spaces = dict(spaces or {})

            def __getitem__(self, key):
                return self.spaces[key]

            def __repr__(self):
                return "Dict(%s)" % self.spaces
