"""Bake the reference's human setup strings into compact tables for on-device auto-reset.

Build-container only: imports the unmodified reference (oracle/ref_shim.py) and pushes EVERY
string of stratego_env/game/inits/{barrage,standard}_human_inits.py through the reference's own
transform ``create_initial_positions_from_human_data`` (util:241-275).  Row i of a table is the
player-1 piece map of string i over the four usable rows (40 piece codes, row-major); the same row
placed row-mirrored gives the player -1 map the reference builds for that string (checked below
for every string), which is what ``sx_reset`` does for ``p2_rot180 = 0``.

Output: stratego_env_b200/data/<name>_setups.npz  (uint8 [n, 40], compressed).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402


def main():
    import_reference()
    from stratego_env.game import util
    from stratego_env.game.config import STANDARD_STRATEGO_CONFIG, BARRAGE_STRATEGO_CONFIG
    from stratego_env.game.inits.standard_human_inits import STANDARD_INITS
    from stratego_env.game.inits.barrage_human_inits import BARRAGE_INITS
    out_dir = os.path.join(ROOT, "stratego_env_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    for name, inits, cfg in (("barrage", BARRAGE_INITS, BARRAGE_STRATEGO_CONFIG),
                             ("standard", STANDARD_INITS, STANDARD_STRATEGO_CONFIG)):
        table = np.zeros((len(inits), 40), dtype=np.uint8)
        for i, s in enumerate(inits):
            maps = util.create_initial_positions_from_human_data(s, s, cfg)
            p1, p2 = np.asarray(maps[0]), np.asarray(maps[1])
            assert not p1[4:].any() and not p2[4:].any()
            table[i] = p1[:4].reshape(-1)
            # reference places player 2's map rotated by 180 degrees (impl:221); net effect = row mirror of p1's map
            p2_abs = p2[::-1, ::-1]
            assert np.array_equal(p2_abs[6:], p1[:4][::-1]), i
        path = os.path.join(out_dir, "%s_setups.npz" % name)
        np.savez_compressed(path, setups=table)
        print(name, table.shape, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
