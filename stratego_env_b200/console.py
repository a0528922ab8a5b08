"""Text rendering of a dense ``int64[34, R, C]`` state for debugging -- the layout the reference's
``StrategoProceduralEnv.print_board_to_console`` (penv:183-216) prints: highest row and column first, player 1's
ranks positive, player 2's negative, lakes as ``R``.  Pure host code; nothing here touches the GPU."""
import numpy as np


def board_to_text(state, partially_observable=False, hide_still_piece_markers=True) -> str:
    state = np.asarray(state)
    _, rows, columns = state.shape
    theirs = state[4] if partially_observable else state[1]   # what player 1 knows of player 2's ranks, or the truth
    rule = "       " + "-" * (rows * 6 + 1) + "\n"             # (the reference sizes the rule by the ROW count)
    text = ["    COL", "".join("  %2d  " % c for c in range(columns - 1, -1, -1)), "\n", rule]
    for r in range(rows - 1, -1, -1):
        text.append("Row %2d |" % r)
        for c in range(columns - 1, -1, -1):
            if state[0, r, c] != 0:
                cell = str(state[0, r, c])
            elif theirs[r, c] != 0:
                cell = str(-theirs[r, c])
            else:
                cell = "R" if state[2, r, c] != 0 else ""
            if not hide_still_piece_markers:
                cell += ("a" if state[32, r, c] == 1 else "") + ("b" if state[33, r, c] == 1 else "")
            text.append(cell.rjust(4) + " |")
        text.append("\n" + rule)
    return "".join(text)
