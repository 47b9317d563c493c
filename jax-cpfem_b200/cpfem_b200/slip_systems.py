"""Slip-system tables, rows "normal(3) direction(3)" (un-normalised Miller indices), the same content and row
order as the reference's data/csv/input_slip_sys*.txt files (e.g. singlecrystal_copper/data/csv/input_slip_sys.txt,
polycrystal_DPsteel/data/csv/input_slip_sys_bcc24.txt).  `load(path)` reads such a file instead."""
import numpy as onp

FCC12 = onp.array([
    [1, 1, -1, 0, 1, 1], [1, 1, -1, 1, 0, 1], [1, 1, -1, 1, -1, 0],
    [1, -1, -1, 0, 1, -1], [1, -1, -1, 1, 0, 1], [1, -1, -1, 1, 1, 0],
    [1, -1, 1, 0, 1, 1], [1, -1, 1, 1, 0, -1], [1, -1, 1, 1, 1, 0],
    [1, 1, 1, 0, 1, -1], [1, 1, 1, 1, 0, -1], [1, 1, 1, 1, -1, 0]], dtype=onp.float64)

BCC12 = onp.array([
    [1, 1, 0, -1, 1, 1], [1, 1, 0, 1, -1, 1], [1, -1, 0, 1, 1, 1], [1, -1, 0, 1, 1, -1],
    [1, 0, 1, 1, 1, -1], [1, 0, 1, -1, 1, 1], [1, 0, -1, 1, 1, 1], [1, 0, -1, 1, -1, 1],
    [0, 1, 1, 1, 1, -1], [0, 1, 1, 1, -1, 1], [0, 1, -1, 1, 1, 1], [0, 1, -1, -1, 1, 1]], dtype=onp.float64)

BCC24 = onp.concatenate([BCC12, onp.array([
    [1, 1, 2, 1, 1, -1], [-1, 1, 2, 1, -1, 1], [1, -1, 2, -1, 1, 1], [1, 1, -2, 1, 1, 1],
    [1, 2, 1, 1, -1, 1], [-1, 2, 1, 1, 1, -1], [1, -2, 1, 1, 1, 1], [1, 2, -1, -1, 1, 1],
    [2, 1, 1, -1, 1, 1], [-2, 1, 1, 1, 1, 1], [2, -1, 1, 1, 1, -1], [2, 1, -1, 1, -1, 1]], dtype=onp.float64)])


def load(path):
    t = onp.loadtxt(path)
    assert t.ndim == 2 and t.shape[1] == 6
    return t
